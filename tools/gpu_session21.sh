#!/bin/bash
mkdir -p gpurun_out
python -m agile3d_b200.build > gpurun_out/build.log 2>&1
timeout 200 python tools/dec_wgrad_diag.py 2>&1 | tail -16 | cut -c1-200
