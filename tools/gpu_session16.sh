#!/bin/bash
mkdir -p gpurun_out
python -m agile3d_b200.build > gpurun_out/build.log 2>&1
for i in 1 2 3; do
CUDA_LAUNCH_BLOCKING=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -x --no-header -k "end_to_end" 2>&1 | grep -E "Ag3dError|passed|failed" | head -3 | cut -c1-300
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py -q -x --no-header 2>&1 | grep -E "Ag3dError|passed|failed|Accel" | head -3 | cut -c1-300
done
timeout 900 python -m pytest tests -m gpu -q --no-header 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_tma2.json 2> gpurun_out/bench_tma2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_tma2.json").read().strip().splitlines()[-1])
print(round(d["value"],1), "scenes/s e2e", round(d["e2e"]["value"],1), {k:v["ms_per_step"] for k,v in d["roofline"]["families"].items()})
PY
