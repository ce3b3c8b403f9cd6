#!/bin/bash
# Quick GPU visit: parity tests + a short bench (no CPU baseline), optional extra command in $1.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/bench_quick.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_quick.log").read().strip().splitlines()[-1])
    print("value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", d["e2e"] and round(d["e2e"]["value"], 1), "launches", d["gpu_launches"])
    for k, v in d["kernels"].items():
        print(" ", k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items()})
    print(" roofline", {k: d["roofline"][k] for k in ("achieved", "frac")}, "tensor", {k: d["roofline_tensor"][k] for k in ("achieved", "frac", "ms_per_launch")}, "variants", {k: round(v["value"], 1) for k, v in d.get("variants", {}).items()})
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench_quick.log").read()[-3000:])
PY
