#!/bin/bash
# ncu --set full capture of fc1 (GELU, planes out) and fc2 of time block 0 in the 4th eager forward of the coarse update
# transformer, autocast (np = 1) and float32-grade (np = 3).  usage: TAG=r02c bash scripts/gpu_gemm_cap.sh
set -u
TAG=${TAG:-r02c}
mkdir -p gpurun_out
for np in 1 3; do
  timeout 600 python scripts/former_profile.py coarse $np > gpurun_out/plain_gemm_np$np.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 333 -c 2 -f -o gpurun_out/${TAG}_gemm_np$np \
    python scripts/former_profile.py coarse $np > gpurun_out/ncu_gemm_np$np.log 2>&1
  echo "np=$np rc=$?"; cat gpurun_out/plain_gemm_np$np.log | tail -1
done
