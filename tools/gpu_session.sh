#!/bin/bash
# Current GPU session (overwritten per call; results land in gpurun_out/ and the kept ones are copied to profiles/).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
S=s28
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/${S}_pytest_gpu.log 2>&1
tail -n 3 gpurun_out/${S}_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${S}_smoke.log 2>&1; tail -n 2 gpurun_out/${S}_smoke.log
timeout 900 python bench.py > gpurun_out/${S}_bench_n1.json 2> gpurun_out/${S}_n1.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${S}_bench_reference.json 2> gpurun_out/${S}_ref.err
timeout 900 python bench.py --steps 8 --batch 1 --no-parity > gpurun_out/${S}_bench_b1.json 2> gpurun_out/${S}_b1.err
timeout 900 python bench.py --steps 8 --workload c2 --no-parity > gpurun_out/${S}_bench_c2.json 2> gpurun_out/${S}_c2.err
timeout 900 python bench.py --steps 3 --warmup 1 --workload clickloop > gpurun_out/${S}_bench_clickloop.json 2> gpurun_out/${S}_clickloop.err
timeout 900 python bench.py --steps 5 --workload train --batch 4 > gpurun_out/${S}_bench_train_b4.json 2> gpurun_out/${S}_train_b4.err
timeout 900 python bench.py --steps 4 --workload train --batch 4 --voxels 500000 > gpurun_out/${S}_bench_train_c4.json 2> gpurun_out/${S}_train_c4.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${S}_launches.csv python tools/profile_step.py --batch 8 > gpurun_out/${S}_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:"c2s_split_kernel|s2c_split_kernel|stem_conv" -c 4 -o /tmp/${S}_ncu_decoder python tools/profile_step.py --batch 8 > gpurun_out/${S}_ncu_decoder.log 2>&1
ncu -i /tmp/${S}_ncu_decoder.ncu-rep --page raw --csv > gpurun_out/${S}_ncu_decoder_raw.csv 2>/dev/null
for f in n1 reference b1 c2 clickloop train_b4 train_c4; do python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${S}_bench_$f.json"))
    print("$f", {k: d.get(k) for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"])
except Exception as e:
    print("$f no json:", e)
PY
done
tail -n 2 gpurun_out/${S}_*.err | tail -n 30
