#!/bin/bash
# Current GPU session (overwritten per call; results land in gpurun_out/ and the kept ones are copied to profiles/).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/s14_pytest_gpu.log 2>&1
tail -n 4 gpurun_out/s14_pytest_gpu.log
timeout 900 python bench.py --steps 8 --no-parity > gpurun_out/s14_bench_n1.json 2> gpurun_out/s14_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 8 --no-parity > gpurun_out/s14_bench_n2.json 2> gpurun_out/s14_n2.err
timeout 900 python bench.py --steps 5 --workload train --batch 4 > gpurun_out/s14_train_n1.json 2> gpurun_out/s14_train_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 5 --workload train --batch 4 > gpurun_out/s14_train_n2_overlap.json 2> gpurun_out/s14_train_n2a.err
AG3D_NO_OVERLAP=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 5 --workload train --batch 4 > gpurun_out/s14_train_n2_after.json 2> gpurun_out/s14_train_n2b.err
for f in bench_n1 bench_n2 train_n1 train_n2_overlap train_n2_after; do python - <<PY
import json
try:
    d = json.load(open("gpurun_out/s14_$f.json"))
    print("$f", {k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, "e2e", d["e2e"]["value"], d["config"].get("gradient_exchange"))
except Exception as e:
    print("$f no json:", e)
PY
done
tail -n 3 gpurun_out/s14_*.err | tail -n 30
