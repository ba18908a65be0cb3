#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_gather_bw tools/probes/tma_gather_bw.cu && timeout 120 /tmp/tma_gather_bw > gpurun_out/tma_gather_bw.txt 2>&1
cat gpurun_out/tma_gather_bw.txt
