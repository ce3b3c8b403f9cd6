"""Times the fine-tracker token kernel (Q sequences: 512*Q patches x 16 frames, C=32, 31x31, L=3, r=3)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import comet_pose_estimation_b200 as cb
Q = int(os.environ.get("FQ", 4)); iters = int(os.environ.get("ITERS", 5))
dev = torch.device("cuda:0")
P = 512 * Q
fm = torch.randn(P, 16, 32, 31, 31, device=dev)
tdim = cb.transformer_dim(3, 3, 32, True)
cos = [torch.rand(P, 16, 1, 2, device=dev) * 30 for _ in range(iters)]
fts = [torch.randn(P, 16, 1, 32, device=dev) for _ in range(iters)]
blk = cb.CorrBlock(fm, num_levels=3, radius=3)
tk = cb.TrackTokenizer(blk, cos[0][:, 0], tdim)
out = torch.empty(P, 1, 16, tdim, device=dev)
for i in range(2): tk.tokens(cos[i], fts[i], out=out)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
ev[0].record()
for i in range(iters):
    tk.tokens(cos[i], fts[i], out=out); ev[i + 1].record()
torch.cuda.synchronize()
ts = [ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(iters)]
print(f"fine tokens Q={Q}: min {min(ts):.1f} us  median {sorted(ts)[len(ts)//2]:.1f} us")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); blk2 = cb.CorrBlock(fm, num_levels=3, radius=3); e1.record(); torch.cuda.synchronize()
print(f"fine pyramid: {e0.elapsed_time(e1)*1e3:.1f} us")
