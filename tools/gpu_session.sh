#!/bin/bash
# Current GPU session (overwritten per call; results land in gpurun_out/ and the kept ones are copied to profiles/).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
tools/probes/bin/commit_probe > gpurun_out/s6_commit_probe.txt 2>&1
O=gpurun_out/s6_pk_probe.txt
: > $O
for d in 31 63 95 127 159 255 223; do AG3D_PK_NA=6 AG3D_PK_NB=6 AG3D_PK_DEBUG=$d timeout 120 python tools/pk_probe.py 96 96 >> $O 2>&1; done
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "query or fused or golden or many" > gpurun_out/s6_pytest_query.log 2>&1
cat gpurun_out/s6_commit_probe.txt; grep -v Warn $O; tail -n 15 gpurun_out/s6_pytest_query.log
