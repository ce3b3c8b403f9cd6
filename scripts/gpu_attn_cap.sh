#!/bin/bash
# ncu --set full capture of the tensor-core attention kernel in the 4th eager autocast forward of the coarse update transformer
set -u
mkdir -p gpurun_out
timeout 600 python scripts/former_profile.py coarse 1 > gpurun_out/plain_attn.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_rows_mma_kernel -s 54 -c 3 -f -o gpurun_out/r02e_attention_mma \
  python scripts/former_profile.py coarse 1 > gpurun_out/ncu_attn.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/plain_attn.log
