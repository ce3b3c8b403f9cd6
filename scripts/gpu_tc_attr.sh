#!/bin/bash
# Attribution of the coarse tcgen05 kernel: time it with parts of the pipeline switched off (COMET_TC_DEBUG bits:
# 1 no epilogue/stager window work, 4 no token part, 8 no MMA, 64 no TMA, 256 no A staging, 512 unsorted/all tiles).
set -u
mkdir -p gpurun_out
for q in ${QS:-4}; do
for d in ${DBGS:-0 512 4 1 5 13 77 333}; do
  COMET_TC_DEBUG=$d TC_Q=$q timeout 120 python scripts/tc_profile.py 2>&1 | grep tokens
done; done
for s in ${SPLITS:-}; do
  COMET_TC_NSPLIT=$s TC_Q=4 timeout 120 python scripts/tc_profile.py 2>&1 | grep tokens | sed "s/^/nsplit=$s /"
done
COMET_TC_DEBUG=0 timeout 120 python scripts/tc_trace.py > gpurun_out/tc_trace.log 2>&1; tail -45 gpurun_out/tc_trace.log
