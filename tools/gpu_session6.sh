#!/bin/bash
# session 6: elect.sync MMA issuers + dual issuer: parity tests, forward bench with 1 and 2 issuers, per-layer times
mkdir -p gpurun_out
python -m agile3d_b200.build > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --no-header -k "not train" 2>&1 | tail -15 > gpurun_out/pytest_gpu_fwd.log
tail -3 gpurun_out/pytest_gpu_fwd.log
AG3D_TC_ISSUERS=1 timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_i1.json 2> gpurun_out/bench_i1.err
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_i2.json 2> gpurun_out/bench_i2.err
timeout 300 python tools/layer_times.py --batch 8 > gpurun_out/layer_times_b8_i2.txt 2>&1
for f in gpurun_out/bench_i1.json gpurun_out/bench_i2.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(d["value"],1), "scenes/s", {k:v["ms_per_step"] for k,v in d["roofline"]["families"].items()})
PY
done
tail -3 gpurun_out/bench_i2.err
timeout 600 python -m pytest tests/test_gpu_train.py -q --no-header 2>&1 | tail -8 > gpurun_out/pytest_gpu_train.log
tail -4 gpurun_out/pytest_gpu_train.log
