#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
S=s46
timeout 1800 python -m pytest tests/test_gpu_train.py -x -q -m gpu > gpurun_out/${S}_pytest_train.log 2>&1
tail -n 3 gpurun_out/${S}_pytest_train.log
timeout 900 python bench.py --steps 6 --workload train --batch 4 > gpurun_out/${S}_train_on.json 2> gpurun_out/${S}_train_on.err
AG3D_WGRAD_STREAM=0 timeout 900 python bench.py --steps 6 --workload train --batch 4 > gpurun_out/${S}_train_off.json 2> gpurun_out/${S}_train_off.err
for f in train_on train_off; do python - <<PY
import json
d = json.load(open("gpurun_out/${S}_$f.json"))
print("$f", {k: d.get(k) for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"])
PY
done
tail -n 2 gpurun_out/${S}_train_on.err
