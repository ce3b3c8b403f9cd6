#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log | cut -c1-200
timeout 900 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench2.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench2.log') if x.startswith('{')][-1]
d=json.loads(l)
print('value',d['value'],'ms/step',d['ms_per_step'])
for k,v in d['kernels'].items(): print(k, {a:round(b,3) if isinstance(b,float) else b for a,b in v.items()})
PY
