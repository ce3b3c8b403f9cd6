#!/bin/bash
# Full end-of-milestone GPU visit: tests, smoke, bench (both arms), ncu launch list + full captures.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench ref rc=$?"
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -c 2500 gpurun_out/bench.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 600 $B > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 600 $B > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:corr_lookup_c32 -s 8 -c 1 -o gpurun_out/r01_fine_tokens $B > gpurun_out/ncu_fine.log 2>&1
echo "ncu fine rc=$?"
timeout 600 $B > gpurun_out/plain3.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:corr_tc_kernel -s 6 -c 1 -o gpurun_out/r01_coarse_tc $B > gpurun_out/ncu_tc.log 2>&1
echo "ncu tc rc=$?"
