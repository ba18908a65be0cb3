#!/bin/bash
mkdir -p gpurun_out
python -m agile3d_b200.build > gpurun_out/build.log 2>&1
bash tools/tc_probe.sh > gpurun_out/tc_probe5.txt 2>&1
cat gpurun_out/tc_probe5.txt
timeout 600 python -m pytest tests -m gpu -q -x --no-header -k "not train" 2>&1 | tail -5
for g in tma cpasync; do
AG3D_TC_GATHER=$g timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$g.json 2> gpurun_out/bench_$g.err
python - $g <<'PY'
import json,sys
try:
    d=json.loads(open(f"gpurun_out/bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(sys.argv[1],round(d["value"],1), "scenes/s e2e", round(d["e2e"]["value"],1), {k:v["ms_per_step"] for k,v in d["roofline"]["families"].items()})
except Exception as e:
    print(sys.argv[1], "failed", e); print(open(f"gpurun_out/bench_{sys.argv[1]}.err").read()[-1500:])
PY
done
