#!/bin/bash
# Current GPU session (overwritten per call; results land in gpurun_out/ and the kept ones are copied to profiles/).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/s13_pytest_gpu.log 2>&1
timeout 900 python bench.py --steps 8 --no-parity > gpurun_out/s13_bench_forward_b8.json 2> gpurun_out/s13_b8.err
timeout 600 python tools/layer_times.py > gpurun_out/s13_layer_times.txt 2>&1
tail -n 8 gpurun_out/s13_pytest_gpu.log
python - <<PY
import json
d = json.load(open("gpurun_out/s13_bench_forward_b8.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"])
for k, v in d["roofline"]["families"].items():
    print("   ", k, v.get("ms_per_step"), v.get("frac_of_hbm_peak"), v.get("binding"), v.get("frac_of_binding_bound"))
PY
tail -n 4 gpurun_out/s13_b8.err; grep -E "spconv|^\{" gpurun_out/s13_layer_times.txt | tail -70
