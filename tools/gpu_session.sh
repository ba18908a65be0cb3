#!/bin/bash
# Current GPU session (overwritten per call; results land in gpurun_out/ and the kept ones are copied to profiles/).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/pk_shapes.py 700 20000 > gpurun_out/s4_pk_small.txt 2>&1
if grep -q FAILED gpurun_out/s4_pk_small.txt; then
  AG3D_PK_REPS=1 timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/pk_shapes.py 700 > gpurun_out/s4_pk_sanitizer.txt 2>&1
else
  timeout 600 python tools/pk_shapes.py 150000 > gpurun_out/s4_pk_150k.txt 2>&1
fi
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "many" -s > gpurun_out/s4_pytest_many.log 2>&1
AG3D_PK_MIN_ROWS=100000000 timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "baseline_shape" -s > gpurun_out/s4_pytest_baseline.log 2>&1
tail -n 12 gpurun_out/s4_pk_small.txt gpurun_out/s4_pk_150k.txt; tail -n 30 gpurun_out/s4_pk_sanitizer.txt; tail -n 8 gpurun_out/s4_pytest_many.log gpurun_out/s4_pytest_baseline.log
