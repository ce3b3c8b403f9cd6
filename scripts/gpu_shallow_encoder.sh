#!/bin/bash
# plain timing, then one ncu --set full capture of the fused patch encoder
set -x
mkdir -p gpurun_out
python scripts/shallow_encoder_profile.py 50 > gpurun_out/senc_plain.log 2>&1 || exit 1
cat gpurun_out/senc_plain.log
ncu --set full --clock-control none --import-source on -k regex:shallow_encoder_kernel -s 1 -c 1 -f -o gpurun_out/${1:-r02c}_shallow_encoder \
  python scripts/shallow_encoder_profile.py 0 > gpurun_out/senc_ncu.log 2>&1
tail -3 gpurun_out/senc_ncu.log
