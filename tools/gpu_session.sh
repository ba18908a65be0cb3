#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
S=s32
timeout 900 python -m pytest tests/test_gpu_train.py -x -q -m gpu -k "loss_kernels" 2>&1 | tail -n 3
timeout 900 python bench.py --no-parity > gpurun_out/${S}_bench_n1.json 2> gpurun_out/${S}_n1.err
timeout 900 python bench.py --no-parity --workload c2 > gpurun_out/${S}_bench_c2.json 2> gpurun_out/${S}_c2.err
for f in n1 c2; do python - <<PY
import json
d = json.load(open("gpurun_out/${S}_bench_$f.json"))
print("$f", {k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
PY
done
tail -n 3 gpurun_out/${S}_n1.err
