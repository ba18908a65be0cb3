#!/bin/bash
mkdir -p gpurun_out
python -m agile3d_b200.build > gpurun_out/build.log 2>&1
bash tools/tc_probe.sh > gpurun_out/tc_probe6.txt 2>&1
cat gpurun_out/tc_probe6.txt
timeout 900 python -m pytest tests -m gpu -q --no-header -k "not train" 2>&1 | tail -2
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_tma3.json 2> gpurun_out/bench_tma3.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_tma3.json").read().strip().splitlines()[-1])
print(round(d["value"],1), "scenes/s e2e", round(d["e2e"]["value"],1), {k:v["ms_per_step"] for k,v in d["roofline"]["families"].items()})
PY
timeout 300 python tools/layer_times.py --batch 8 > gpurun_out/layer_times_b8_tma.txt 2>&1
