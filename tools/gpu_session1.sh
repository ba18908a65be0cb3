#!/bin/bash
# first GPU session: parity tests, smoke, bench, ncu launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
python -m agile3d_b200.build > gpurun_out/build.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 -x --no-header -rA 2>&1 | tail -150 > gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
