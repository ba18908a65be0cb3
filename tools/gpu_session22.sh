#!/bin/bash
mkdir -p gpurun_out
python -m agile3d_b200.build > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --no-header -k "not train" 2>&1 | grep -E "^E  |passed|failed|^FAILED" | head -12 | cut -c1-250
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_s22.json 2> gpurun_out/bench_s22.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_s22.json").read().strip().splitlines()[-1])
print(round(d["value"],1), "scenes/s e2e", round(d["e2e"]["value"],1), {k:v["ms_per_step"] for k,v in d["roofline"]["families"].items()})
PY
