#!/bin/bash
# round-1 session 3: full GPU test-suite, forward + train bench, launch list, full ncu capture of the split-row TC conv
mkdir -p gpurun_out
python -m agile3d_b200.build > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --no-header 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python bench.py --workload train --batch 8 --steps 3 --warmup 3 > gpurun_out/bench_train_b8.json 2> gpurun_out/bench_train_b8.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:spconv_tc_kernel \
    --launch-skip 50 -c 6 -o gpurun_out/prof_spconv_tc_split -f python tools/profile_step.py > gpurun_out/ncu_full.log 2>&1
tail -8 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
cat gpurun_out/bench_train_b8.json; tail -3 gpurun_out/bench_train_b8.err; tail -3 gpurun_out/ncu_full.log
