#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
S=s35
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/${S}_pytest_gpu.log 2>&1
tail -n 3 gpurun_out/${S}_pytest_gpu.log
timeout 300 python tools/dec_split_time.py 2>&1 | tail -5
timeout 900 python bench.py --steps 12 > gpurun_out/${S}_bench_n1.json 2> gpurun_out/${S}_n1.err
python - <<PY
import json
d = json.load(open("gpurun_out/${S}_bench_n1.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"], d["parity"]["mask_logits_rel_err_per_layer"])
print({k: v["ms_per_step"] for k, v in d["roofline"]["families"].items()})
PY
