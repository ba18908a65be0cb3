#!/bin/bash
# round-1 session 4: full GPU test-suite, forward + train bench, launch list, full ncu captures (spconv / c2s / s2c)
mkdir -p gpurun_out
python -m agile3d_b200.build > gpurun_out/build.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt
nproc >> gpurun_out/smi.txt
timeout 1100 python -m pytest tests -m gpu -q --maxfail=15 --no-header 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
timeout 500 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 500 python bench.py --workload train --batch 8 --steps 3 --warmup 3 > gpurun_out/bench_train_b8.json 2> gpurun_out/bench_train_b8.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:spconv_tc_kernel \
    --launch-skip 2 -c 8 -o gpurun_out/prof_spconv_tc -f python tools/profile_step.py > gpurun_out/ncu_full_spconv.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'c2s|s2c' \
    -c 6 -o gpurun_out/prof_decoder -f python tools/profile_step.py > gpurun_out/ncu_full_dec.log 2>&1
tail -8 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
cat gpurun_out/bench_train_b8.json; tail -3 gpurun_out/bench_train_b8.err; tail -3 gpurun_out/ncu_full_spconv.log
