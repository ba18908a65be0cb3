#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
S=s50
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/${S}_pytest_gpu.log 2>&1
tail -n 3 gpurun_out/${S}_pytest_gpu.log
timeout 900 python bench.py --steps 12 --no-parity > gpurun_out/${S}_bench_n1.json 2> gpurun_out/${S}_n1.err
timeout 900 python bench.py --steps 12 --no-parity --batch 1 > gpurun_out/${S}_bench_b1.json 2> gpurun_out/${S}_b1.err
timeout 900 python bench.py --steps 6 --workload train --batch 4 > gpurun_out/${S}_bench_train_b4.json 2> gpurun_out/${S}_train.err
timeout 900 python bench.py --steps 3 --warmup 1 --workload clickloop > gpurun_out/${S}_bench_clickloop.json 2> gpurun_out/${S}_clickloop.err
for f in n1 b1 train_b4 clickloop; do python - <<PY
import json
d = json.load(open("gpurun_out/${S}_bench_$f.json"))
print("$f", {k: d.get(k) for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"], d["config"].get("ms_per_round"))
PY
done
