#!/bin/bash
# Current GPU session (overwritten per call; results land in gpurun_out/ and the kept ones are copied to profiles/).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tools/ddp_overlap_trace.py > gpurun_out/s54_ddp_overlap_trace.txt 2> gpurun_out/s54_trace.err
cat gpurun_out/s54_ddp_overlap_trace.txt | cut -c1-260; tail -n 5 gpurun_out/s54_trace.err
