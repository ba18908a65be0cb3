#!/bin/bash
mkdir -p gpurun_out
python -m agile3d_b200.build > gpurun_out/build.log 2>&1
bash tools/tc_probe.sh > gpurun_out/tc_probe4.txt 2>&1
cat gpurun_out/tc_probe4.txt
timeout 300 python tools/grad_diag.py > gpurun_out/grad_diag.txt 2>&1; grep -A200 "all parameters" gpurun_out/grad_diag.txt | cut -c1-100
