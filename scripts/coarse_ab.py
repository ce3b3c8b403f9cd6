"""A/B of the coarse token path options on the bench shape (batch 4): ms per iteration, CUDA-event timed."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import comet_pose_estimation_b200 as cb
from comet_pose_estimation_b200 import _lib
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
Q, S, N = 4, 16, 512
fm = torch.randn(Q, S, 128, 64, 64, device=dev, generator=g)
co = torch.rand(Q, S, N, 2, device=dev, generator=g) * 63
ft = torch.randn(Q, S, N, 128, device=dev, generator=g)
tdim = cb.transformer_dim(5, 4, 128, False)
blk = cb.CorrBlock(fm, num_levels=5, radius=4)
tk = cb.TrackTokenizer(blk, co[:, 0], tdim)
out = torch.empty(Q, N, S, tdim, device=dev)
ref = None
OPT = getattr(_lib, sys.argv[1]) if len(sys.argv) > 1 else _lib.OPT_TC_REDUCE_STORE
for name, val in (("option off", False), ("option on", True), ("option off", False), ("option on", True)):
    _lib.set_option(OPT, val)
    for _ in range(5): tk.tokens(co, ft, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): tk.tokens(co, ft, out=out)
    e1.record(); torch.cuda.synchronize()
    if ref is None: ref = out.clone()
    print(f"{name:28s} {e0.elapsed_time(e1) / 50:.4f} ms/iteration   max|diff| vs first {float((out - ref).abs().max()):.1e}")
