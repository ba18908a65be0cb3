#!/bin/bash
mkdir -p gpurun_out
python -m agile3d_b200.build > gpurun_out/build.log 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py -q -x --no-header 2>&1 | grep -E "^E  |passed|failed|^FAILED|Error" | head -8 | cut -c1-250
timeout 200 python tools/phase_times.py 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_s26.json 2> gpurun_out/bench_s26.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/bench_s26.json").read().strip().splitlines()[-1])
    print(round(d["value"],1), "scenes/s e2e", round(d["e2e"]["value"],1), {k:v["ms_per_step"] for k,v in d["roofline"]["families"].items()})
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_s26.err").read()[-800:])
PY
