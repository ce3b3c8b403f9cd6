#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python scripts/fine_profile.py > gpurun_out/plain_fine.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:corr_lookup_c32 -s 2 -c 1 -o gpurun_out/fine_r01b python scripts/fine_profile.py > gpurun_out/ncu_fine.log 2>&1
echo "ncu rc=$?"; cat gpurun_out/plain_fine.log
