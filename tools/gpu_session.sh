#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
S=s24
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/${S}_pytest_gpu.log 2>&1
tail -n 3 gpurun_out/${S}_pytest_gpu.log
timeout 900 python bench.py --steps 8 --no-parity > gpurun_out/${S}_bench_n1.json 2> gpurun_out/${S}_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${S}_launches.csv python tools/profile_step.py --batch 8 > gpurun_out/${S}_launches.log 2>&1
python - <<PY
import json
d = json.load(open("gpurun_out/${S}_bench_n1.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"])
PY
