#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
S=s26
nproc; free -g | head -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 --no-parity > gpurun_out/${S}_bench_n8.json 2> gpurun_out/${S}_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 4 --steps 10 --warmup 3 --no-parity > gpurun_out/${S}_bench_n4.json 2> gpurun_out/${S}_n4.err
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-parity > gpurun_out/${S}_bench_n1.json 2> gpurun_out/${S}_n1.err
for f in n8 n4 n1; do python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${S}_bench_$f.json"))
    print("$f", {k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, "e2e", d["e2e"]["value"], d["e2e"].get("ms_per_step"))
except Exception as e:
    print("$f no json:", e)
PY
done
tail -n 4 gpurun_out/${S}_n8.err
