#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python tools/host_profile_train.py > gpurun_out/s55_host_profile_train.txt 2>&1
head -n 48 gpurun_out/s55_host_profile_train.txt | cut -c1-170
