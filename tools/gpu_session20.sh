#!/bin/bash
mkdir -p gpurun_out
python -m agile3d_b200.build > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_train.py -q --no-header 2>&1 | grep -E "^E  |passed|failed|^FAILED" | head -12 | cut -c1-250
timeout 400 python bench.py --workload train --batch 8 --steps 3 --warmup 3 > gpurun_out/bench_train_b8_tc.json 2> gpurun_out/bench_train_b8_tc.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/bench_train_b8_tc.json").read().strip().splitlines()[-1])
    print("train", round(d["value"],2), "scenes/s", round(d["ms_per_step"],1), "ms/step", {k:v["ms_per_step"] for k,v in d["roofline"]["families"].items()})
except Exception as e:
    print("train bench failed", e); print(open("gpurun_out/bench_train_b8_tc.err").read()[-1500:])
PY
