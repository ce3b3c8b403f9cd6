#!/bin/bash
# Full GPU visit: tests, smoke, bench (both arms), ncu launch list + full captures of the hot kernels.
# usage: TAG=r01b bash scripts/gpu_round.sh
set -u
TAG=${TAG:-r01b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench ref rc=$?"
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-variants"
timeout 600 $B > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
cap() {  # name regex skip [extra bench args]
  timeout 600 $B ${4:-} > gpurun_out/plain_$1.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o gpurun_out/${TAG}_$1 $B ${4:-} > gpurun_out/ncu_$1.log 2>&1
  echo "ncu $1 rc=$?"
}
cap fine_tokens_cl corr_lookup_c32_tma 8
cap fine_pyramid_cl pyramid_cl_in_fine 2
cap coarse_tc corr_tc_kernel 6
cap coarse_pre tc_pre_kernel 6
if [ "${NCHW:-0}" = "1" ]; then
cap fine_tokens_nchw "corr_lookup_c32_kernel" 8 "--fine-layout nchw"
cap fine_pyramid_nchw "pyramid_cl_fine_kernel" 2 "--fine-layout nchw"
fi
