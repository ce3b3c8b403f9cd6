TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 2 4; do
  timeout 600 $TR --nproc-per-node $n --master-port $((29560 + n)) bench.py --gpus $n --steps 10 --warmup 3 --no-variants > gpurun_out/bench_n$n.log 2>&1; echo "bench n=$n rc=$?"
  grep '^{' gpurun_out/bench_n$n.log | tail -1 > gpurun_out/bench_final_n$n.json
  python -c "
import json
d=json.loads(open('gpurun_out/bench_final_n$n.json').read()); e=d['e2e']
print('N', d['n_gpus'], 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(e['value'],1), 'h2d GB/s per gpu', round(e['h2d_GBps_per_gpu'],1))"
done
