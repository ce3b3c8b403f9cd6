#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list.  Usage: gpurun -- bash scripts/gpu_check.sh
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" | tee -a gpurun_out/bench.log
tail -c 3000 gpurun_out/bench.log
BENCH_SMALL="python bench.py --steps 2 --warmup 3 --batch 2 --no-cpu-baseline --no-e2e"
timeout 600 $BENCH_SMALL > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv $BENCH_SMALL > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"
tail -5 gpurun_out/pytest_gpu.log gpurun_out/smoke.log
