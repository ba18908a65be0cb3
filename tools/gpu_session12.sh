#!/bin/bash
mkdir -p gpurun_out
python -m agile3d_b200.build > gpurun_out/build.log 2>&1
timeout 300 python tools/grad_diag2.py > gpurun_out/grad_diag2.txt 2>&1; tail -34 gpurun_out/grad_diag2.txt | cut -c1-300
nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_gather_probe tools/probes/tma_gather_probe.cu && timeout 60 /tmp/tma_gather_probe > gpurun_out/tma_gather_probe.txt 2>&1
tail -24 gpurun_out/tma_gather_probe.txt
