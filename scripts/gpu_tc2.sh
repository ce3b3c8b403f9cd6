#!/bin/bash
set -u
mkdir -p gpurun_out
for d in 0 1 4 5 8 9 13; do COMET_TC_DEBUG=$d timeout 120 python scripts/tc_profile.py 2>&1 | tail -2; done
TC_Q=4 timeout 120 python scripts/tc_profile.py 2>&1 | tail -2
TC_ITERS=2 timeout 120 python scripts/tc_profile.py > gpurun_out/plain_tc.log 2>&1 && \
TC_ITERS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:corr_tc_kernel -s 3 -c 1 -o gpurun_out/tc_r01a python scripts/tc_profile.py > gpurun_out/ncu_tc.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_tc.log
