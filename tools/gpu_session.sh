#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
S=s37
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/${S}_pytest_gpu.log 2>&1
tail -n 3 gpurun_out/${S}_pytest_gpu.log
timeout 300 python tools/tc_level0_probe.py 96 96 2>&1 | tail -1
AG3D_TC_TMA_OUT=0 timeout 300 python tools/tc_level0_probe.py 96 96 2>&1 | tail -1
timeout 300 python tools/tc_level0_probe.py 128 96 2>&1 | tail -1
AG3D_TC_TMA_OUT=0 timeout 300 python tools/tc_level0_probe.py 128 96 2>&1 | tail -1
timeout 900 python bench.py --steps 12 > gpurun_out/${S}_bench_n1.json 2> gpurun_out/${S}_n1.err
AG3D_TC_TMA_OUT=0 timeout 900 python bench.py --steps 12 --no-parity > gpurun_out/${S}_bench_n1_off.json 2> gpurun_out/${S}_n1_off.err
for f in n1 n1_off; do python - <<PY
import json
d = json.load(open("gpurun_out/${S}_bench_$f.json"))
print("$f", {k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"], d.get("parity") and d["parity"]["mask_logits_rel_err_per_layer"])
print("   ", {k: v["ms_per_step"] for k, v in d["roofline"]["families"].items()})
PY
done
