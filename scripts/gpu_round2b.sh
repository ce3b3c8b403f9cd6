#!/bin/bash
# Round-2 second profiling visit (after the coarse-kernel rework): ncu launch list of the bench step + one
# `ncu --set full` capture of the coarse tensor kernel and of its pre-kernel (each only after the same command has
# exited 0 without ncu).  usage: TAG=r02b bash scripts/gpu_round2b.sh
set -u
TAG=${TAG:-r02b}
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-variants"
timeout 600 $B > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
cap() {  # name regex skip count cmd...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 600 "$@" > gpurun_out/plain_$name.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -f -o gpurun_out/${TAG}_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "ncu $name rc=$?"
}
cap coarse_tc corr_tc_kernel 6 1 $B
cap coarse_pre tc_pre_kernel 6 1 $B
ls -la gpurun_out/${TAG}_*
