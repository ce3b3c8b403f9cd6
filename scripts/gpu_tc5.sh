#!/bin/bash
set -u
timeout 300 python scripts/tc_bringup.py 2>&1 | grep -E "rel err|status|us/iter|Error|error|Traceback" | tail -5
for d in 0 1 4; do COMET_TC_DEBUG=$d timeout 120 python scripts/tc_profile.py 2>&1 | tail -2 ; done
TC_Q=4 timeout 120 python scripts/tc_profile.py 2>&1 | tail -2
COMET_TC_DEBUG=0 timeout 120 python scripts/tc_trace.py 2>&1 | tail -36 | head -9
