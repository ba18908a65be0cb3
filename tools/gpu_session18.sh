#!/bin/bash
# profiles for the TMA-gather build: launch list at the bench batch size, full ncu captures of the level-0 convs and decoder kernels
mkdir -p gpurun_out
python -m agile3d_b200.build > gpurun_out/build.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_b8.csv python tools/profile_step.py --batch 8 > gpurun_out/ncu_launches.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:spconv_tc_kernel \
    --launch-skip 56 -c 5 -o gpurun_out/prof_spconv_tma_L0 -f python tools/profile_step.py --batch 8 > gpurun_out/ncu_full_spconv.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'c2s_tc|s2c_tc' \
    -c 2 -o gpurun_out/prof_decoder_b8 -f python tools/profile_step.py --batch 8 > gpurun_out/ncu_full_dec.log 2>&1
tail -2 gpurun_out/ncu_full_spconv.log gpurun_out/ncu_full_dec.log
