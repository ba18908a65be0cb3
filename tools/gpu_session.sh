#!/bin/bash
# Current GPU session (overwritten per call; results land in gpurun_out/ and the kept ones are copied to profiles/).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/s1_gpu.txt 2>&1
timeout 120 tools/probes/bin/epi_probe > gpurun_out/s1_epi_probe.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "baseline_shape" -s > gpurun_out/s1_pytest_baseline.log 2>&1
timeout 900 python -m pytest tests/test_gpu_train.py -q -m gpu > gpurun_out/s1_pytest_train.log 2>&1
timeout 600 python tools/train_err.py > gpurun_out/s1_train_err.txt 2>&1
timeout 600 python bench.py --steps 5 > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err
tail -3 gpurun_out/s1_epi_probe.txt gpurun_out/s1_pytest_baseline.log gpurun_out/s1_pytest_train.log gpurun_out/s1_train_err.txt gpurun_out/s1_bench.json
