#!/bin/bash
set -u
mkdir -p gpurun_out
TC_ITERS=2 timeout 120 python scripts/tc_profile.py > gpurun_out/plain_tc.log 2>&1 && \
TC_ITERS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:corr_tc_kernel -s 3 -c 1 -o gpurun_out/tc_r01d python scripts/tc_profile.py > gpurun_out/ncu_tc.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_tc.log
