#!/bin/bash
# round-1 session 5 (2 GPUs): torchrun bench at N=2 (forward, train with NCCL gradient all-reduce), reference arm, per-layer times
mkdir -p gpurun_out
python -m agile3d_b200.build > gpurun_out/build.log 2>&1
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --workload train --batch 4 --steps 3 --warmup 3 > gpurun_out/bench_train_n2.json 2> gpurun_out/bench_train_n2.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 300 python tools/layer_times.py --batch 8 > gpurun_out/layer_times_b8.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_train.py -q -k backbone_backward --no-header 2>&1 | tail -15 > gpurun_out/pytest_train_fix.log
cat gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err; cat gpurun_out/bench_train_n2.json; tail -3 gpurun_out/bench_train_n2.err
cat gpurun_out/bench_reference.json; tail -5 gpurun_out/pytest_train_fix.log
