#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
S=s27
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "stem or golden or maps" 2>&1 | tail -n 4
timeout 900 python bench.py --steps 8 --no-parity > gpurun_out/${S}_bench_n1.json 2> gpurun_out/${S}_n1.err
python - <<PY
import json
d = json.load(open("gpurun_out/${S}_bench_n1.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"])
print({k: v["ms_per_step"] for k, v in d["roofline"]["families"].items()})
PY
