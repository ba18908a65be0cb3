#!/bin/bash
mkdir -p gpurun_out
python -m agile3d_b200.build > gpurun_out/build.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_train_b2.csv python tools/profile_train_step.py > gpurun_out/ncu_launches_train.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:wgrad_tc_kernel \
    --launch-skip 4 -c 4 -o gpurun_out/prof_wgrad_tc -f python tools/profile_train_step.py > gpurun_out/ncu_full_wgrad.log 2>&1
tail -2 gpurun_out/ncu_launches_train.log; tail -2 gpurun_out/ncu_full_wgrad.log
