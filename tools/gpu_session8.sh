#!/bin/bash
mkdir -p gpurun_out
python -m agile3d_b200.build > gpurun_out/build.log 2>&1
bash tools/tc_probe.sh > gpurun_out/tc_probe2.txt 2>&1
cat gpurun_out/tc_probe2.txt
timeout 600 python -m pytest tests -m gpu -q -x --no-header -k "not train" 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_s8.json 2> gpurun_out/bench_s8.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_s8.json").read().strip().splitlines()[-1])
print(round(d["value"],1), "scenes/s e2e", round(d["e2e"]["value"],1), {k:v["ms_per_step"] for k,v in d["roofline"]["families"].items()})
PY
timeout 300 python tools/grad_diag.py > gpurun_out/grad_diag.txt 2>&1; cat gpurun_out/grad_diag.txt | cut -c1-200
