#!/bin/bash
# Current GPU session (overwritten per call; results land in gpurun_out/ and the kept ones are copied to profiles/).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
tools/probes/bin/mma_probe > gpurun_out/s10_mma_probe.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_interactive.py -q -m gpu -x > gpurun_out/s10_pytest_interactive.log 2>&1
cat gpurun_out/s10_mma_probe.txt; tail -n 25 gpurun_out/s10_pytest_interactive.log
