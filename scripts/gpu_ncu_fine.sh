#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python scripts/fine_profile.py > gpurun_out/plain_fine.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pyramid_cl -s 1 -c 1 -f -o gpurun_out/r01c_fine_pyramid python scripts/fine_profile.py > gpurun_out/ncu_fine_pyr.log 2>&1
echo "ncu pyr rc=$?"
cat gpurun_out/plain_fine.log
