#!/bin/bash
set -u
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.log 2>&1; echo "bench N=$N rc=$?"
grep '^{' gpurun_out/bench_n$N.log | tail -1 | cut -c1-700
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.log 2>&1; echo "ref N=$N rc=$?"
grep '^{' gpurun_out/bench_ref_n$N.log | tail -1 | cut -c1-200
tail -3 gpurun_out/bench_n$N.log | cut -c1-300
