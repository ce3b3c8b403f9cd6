#!/bin/bash
# A/B of coarse-kernel builds: each comet_pose_estimation_b200/build/variants/lib_<name>.so is swapped in and the short bench run.
set -u
cp comet_pose_estimation_b200/libcomet_b200.so /tmp/lib_orig.so
for f in comet_pose_estimation_b200/build/variants/lib_*.so; do
  n=$(basename $f .so)
  cp $f comet_pose_estimation_b200/libcomet_b200.so
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-variants > gpurun_out/vb_$n.json 2> gpurun_out/vb_$n.err
  python - "$n" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/vb_{n}.json").read().strip().splitlines()[-1])
    print(n, "coarse_tokens ms", round(d["kernels"]["coarse_tokens"]["ms_avg"], 4), "coarse_pyramid ms", round(d["kernels"]["coarse_pyramid"]["ms_avg"], 4), "step", round(d["ms_per_step"], 3))
except Exception as e:
    print(n, "failed", e, open(f"gpurun_out/vb_{n}.err").read()[-500:])
PY
done
cp /tmp/lib_orig.so comet_pose_estimation_b200/libcomet_b200.so
