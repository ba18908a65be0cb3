#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
S=s47
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "split_rows_vs_oracle and (1-11 or 63-20 or 127-20)" 2>&1 | tail -n 25
