"""CUDA-event timing of single GEMMs of the update transformer (shapes of one coarse time block, M = 9216 rows):
python scripts/gemm_time.py  -> us per launch for np in (1, 3) x {qkv, out_proj + resid, fc1 + GELU -> planes, fc2 + resid}."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from comet_pose_estimation_b200 import update_former_tc as tc

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)


def timed(fn, n=200):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


from comet_pose_estimation_b200 import _lib
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 3
_lib.check(_lib.lib.comet_set_option(_lib.OPT_GEMM_TMA_STORE, mode))   # bit 0: float32 result by TMA store, bit 1: planes
ew16 = int(sys.argv[2]) if len(sys.argv) > 2 else 1
_lib.check(_lib.lib.comet_set_option(_lib.OPT_GEMM_EW16, ew16))
bn96 = int(sys.argv[3]) if len(sys.argv) > 3 else 1
_lib.check(_lib.lib.comet_set_option(_lib.OPT_GEMM_BN96, bn96))
pair = int(sys.argv[4]) if len(sys.argv) > 4 else int(_lib.lib.comet_get_option(_lib.OPT_GEMM_PAIR))
_lib.check(_lib.lib.comet_set_option(_lib.OPT_GEMM_PAIR, pair))
print("tma store mode", mode, "ew16", ew16, "bn96", bn96, "pair", pair)
for np_ in (1, 3):
    run = tc._Run(tc._Weights(), np_, dev)
    M = 9216
    res = {}
    for name, K, N, kw in (("qkv", 384, 1152, dict()), ("out_proj+resid", 384, 384, dict(resid=True)),
                           ("fc1+gelu->planes", 384, 1536, dict(gelu=True, want_f32=False, want_planes=True)),
                           ("fc1 no gelu->planes", 384, 1536, dict(gelu=False, want_f32=False, want_planes=True)),
                           ("fc1 no gelu->f32", 384, 1536, dict(gelu=False)),
                           ("fc2+resid", 1536, 384, dict(resid=True))):
        x = torch.randn(M, K, device=dev, generator=g)
        w = torch.randn(N, K, device=dev, generator=g) / K ** 0.5
        b = torch.randn(N, device=dev, generator=g)
        r = torch.randn(M, N, device=dev, generator=g) if kw.pop("resid", False) else None
        xp = run.split(x)
        out = torch.empty(M, N, device=dev) if kw.get("want_f32", True) else None
        # the workspace allocator of _Run hands out fresh tensors per call: reuse explicit outputs where the API allows
        res[name] = round(timed(lambda: run.linear(xp, w, b, resid=r, out=out, **kw)), 1)
        flops = 2.0 * M * K * N * (1 if np_ == 1 else 6)
        res[name + " TF/s-equiv"] = round(flops / res[name] * 1e-6, 0)
    print("np", np_, res)
