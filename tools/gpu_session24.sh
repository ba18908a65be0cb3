#!/bin/bash
# 2 GPUs, final build: torchrun forward bench and training bench (gradient all-reduce) at N=2
mkdir -p gpurun_out
python -m agile3d_b200.build > gpurun_out/build.log 2>&1
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/final_bench_n2.json 2> gpurun_out/final_bench_n2.err
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 \
    bench.py --gpus 2 --workload train --batch 4 --steps 3 --warmup 3 > gpurun_out/final_bench_train_n2.json 2> gpurun_out/final_bench_train_n2.err
python - <<'PY'
import json
for f in ("final_bench_n2","final_bench_train_n2"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "n_gpus", d["n_gpus"], round(d["value"],1), "scenes/s", round(d["ms_per_step"],1), "ms/step e2e", round(d["e2e"]["value"],1))
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/{f}.err").read()[-1200:])
PY
