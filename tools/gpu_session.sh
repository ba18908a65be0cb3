#!/bin/bash
# Current GPU session (overwritten per call; results land in gpurun_out/ and the kept ones are copied to profiles/).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python tools/parity_diag.py headline > gpurun_out/s2_parity_diag.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "baseline_shape" -s > gpurun_out/s2_pytest_baseline.log 2>&1
tail -n 30 gpurun_out/s2_parity_diag.txt; tail -n 8 gpurun_out/s2_pytest_baseline.log
