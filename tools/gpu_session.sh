#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
S=s48
timeout 600 python -m pytest tests/test_gpu_interactive.py -x -q -m gpu 2>&1 | tail -n 4
timeout 900 python bench.py --steps 3 --warmup 1 --workload clickloop > gpurun_out/${S}_bench_clickloop.json 2> gpurun_out/${S}_clickloop.err
python - <<PY
import json
d = json.load(open("gpurun_out/${S}_bench_clickloop.json"))
print({k: d.get(k) for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"], d["config"].get("ms_per_round"))
PY
tail -n 3 gpurun_out/${S}_clickloop.err
