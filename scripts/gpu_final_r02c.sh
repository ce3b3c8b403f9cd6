#!/bin/bash
# Final visit of the third pass: GPU tests, both bench arms, ncu launch list of the bench step, GEMM captures.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench11.json 2> gpurun_out/bench11.err; echo "bench rc=$?"
python bench.py --impl reference > gpurun_out/bench11_ref.json 2> gpurun_out/bench11_ref.err; echo "ref rc=$?"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-variants"
timeout 600 $B > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02c_launches.csv $B > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
TAG=r02e bash scripts/gpu_gemm_cap.sh 2>&1 | tail -4
python scripts/former_check.py 2>&1 | grep "^former"
python scripts/gemm_time.py 2>&1 | tail -2
