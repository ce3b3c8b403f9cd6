#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python scripts/tc_bringup.py > gpurun_out/tc_bringup_v2.log 2>&1; echo "bringup rc=$?"
grep -E "rel err|status|us/iter|Error|error|Traceback" gpurun_out/tc_bringup_v2.log | head -30
for d in 0 1 4 5 8 13 64 77; do COMET_TC_DEBUG=$d timeout 120 python scripts/tc_profile.py 2>&1 | tail -2; done
TC_Q=4 timeout 120 python scripts/tc_profile.py 2>&1 | tail -2
