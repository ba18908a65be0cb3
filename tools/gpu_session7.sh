#!/bin/bash
mkdir -p gpurun_out
python -m agile3d_b200.build > gpurun_out/build.log 2>&1
bash tools/tc_probe.sh > gpurun_out/tc_probe.txt 2>&1
cat gpurun_out/tc_probe.txt
timeout 300 python -m pytest tests/test_gpu_train.py -q --no-header -k "backbone_backward" 2>&1 | grep -E "^E|assert|passed|failed" | head -20 > gpurun_out/pytest_gpu_train2.log
cat gpurun_out/pytest_gpu_train2.log
timeout 600 python -m pytest tests -m gpu -q -x --no-header -k "stem or golden or end_to_end" 2>&1 | tail -4
timeout 300 python tools/layer_times.py --batch 8 2>&1 | grep -E "stem|^\{" 
