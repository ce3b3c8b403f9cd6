"""Three iterations of the coarse token path per output mode (stores / bulk reductions), for an ncu launch list:
ncu --metrics gpu__time_duration.sum --clock-control none --csv python scripts/coarse_modes.py"""
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
for val in (False, True):
    _lib.set_option(_lib.OPT_TC_REDUCE_STORE, val)
    for _ in range(3): tk.tokens(co, ft, out=out)
    torch.cuda.synchronize()
print("done")
# the lookup layout (no token rows): its pre-kernel launch is the plan alone
blk.corr(ft)
for _ in range(3): blk.sample(co)
torch.cuda.synchronize()
print("lookup done")
