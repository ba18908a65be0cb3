#!/bin/bash
# Current GPU session (overwritten per call; results land in gpurun_out/ and the kept ones are copied to profiles/).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out/s8_pk_prof.txt
: > $O
for d in 0 31 255; do AG3D_PK_PROF=1 AG3D_PK_DEBUG=$d timeout 120 python tools/pk_probe.py 96 96 >> $O 2>&1; done
timeout 900 python -m pytest tests/test_gpu_train.py -q -m gpu -k "bwd" > gpurun_out/s8_pytest_bwd.log 2>&1
grep -v Warn $O | cut -c1-700; tail -n 15 gpurun_out/s8_pytest_bwd.log
