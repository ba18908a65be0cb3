#!/bin/bash
# Current GPU session (overwritten per call; results land in gpurun_out/ and the kept ones are copied to profiles/).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
S=s53
export AG3D_DEC_STREAMS=0
timeout 1500 ncu --set full --clock-control none --profile-from-start off -k regex:spconv_tc_kernel -c 70 -o /tmp/${S}_ncu_spconv python tools/profile_step.py --batch 8 > gpurun_out/${S}_ncu_spconv.log 2>&1
ncu -i /tmp/${S}_ncu_spconv.ncu-rep --page raw --csv > gpurun_out/${S}_ncu_spconv_raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:"s2c_split_kernel|c2s_split_kernel" -c 4 -o /tmp/${S}_ncu_decoder python tools/profile_step.py --batch 8 > gpurun_out/${S}_ncu_decoder.log 2>&1
ncu -i /tmp/${S}_ncu_decoder.ncu-rep --page raw --csv > gpurun_out/${S}_ncu_decoder_raw.csv 2>/dev/null
ls -la gpurun_out/${S}_*; tail -n 2 gpurun_out/${S}_ncu_spconv.log
