#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): concurrent pinned H2D ceiling at 1/2/../N ranks, then both bench arms at N ranks.
# usage: N=4 bash scripts/gpu_multi2.sh
set -u
N=${N:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
: > gpurun_out/h2d_scaling.jsonl
for n in 1 2 4 8; do
  [ $n -gt $N ] && break
  for extra in "" "--streams 2 --chunk-mb 128" "--numa"; do
    timeout 300 $TR --nproc-per-node $n --master-port $((29540 + n)) scripts/h2d_scaling.py --mb 1024 --reps 8 $extra 2>/dev/null | grep '^{' >> gpurun_out/h2d_scaling.jsonl
  done
done
cat gpurun_out/h2d_scaling.jsonl
for n in 2 $N; do
  [ $n -gt $N ] && continue
  timeout 900 $TR --nproc-per-node $n --master-port $((29560 + n)) bench.py --gpus $n --steps 10 --warmup 3 --no-variants > gpurun_out/bench_n$n.log 2>&1; echo "bench n=$n rc=$?"
  grep '^{' gpurun_out/bench_n$n.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']
print('N', d['n_gpus'], 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(e['value'],1), 'ms', round(e['ms_per_step'],2), 'h2d GB/s per gpu', round(e['h2d_GBps_per_gpu'],1), 'numa', d.get('numa_bound'))"
done
