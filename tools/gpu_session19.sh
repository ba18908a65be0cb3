#!/bin/bash
mkdir -p gpurun_out
python -m agile3d_b200.build > gpurun_out/build.log 2>&1
timeout 120 python tools/wgrad_shapes.py 50 1000 > gpurun_out/wgrad_shapes.txt 2>&1; tail -30 gpurun_out/wgrad_shapes.txt | cut -c1-200
