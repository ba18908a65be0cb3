#!/bin/bash
mkdir -p gpurun_out
python -m agile3d_b200.build > gpurun_out/build.log 2>&1
timeout 100 python tools/wgrad_shapes.py 50 1000 2>&1 | grep -E "WRONG|FAILED|done" | head -5
timeout 200 python tools/wgrad_shapes.py 150000 > gpurun_out/wgrad_shapes_150k.txt 2>&1; tail -16 gpurun_out/wgrad_shapes_150k.txt | cut -c1-200
