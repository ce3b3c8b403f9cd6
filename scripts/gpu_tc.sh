#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python scripts/tc_bringup.py > gpurun_out/tc_bringup.log 2>&1; echo "bringup rc=$?"
tail -60 gpurun_out/tc_bringup.log
