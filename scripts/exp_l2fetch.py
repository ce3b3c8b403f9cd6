"""Experiment: does cudaLimitMaxL2FetchGranularity change the fine-tracker token kernel (32-byte window rows)?"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import comet_pose_estimation_b200 as cb
lib = cb._lib.lib
Q = 4; iters = 6
dev = torch.device("cuda:0")
P = 512 * Q
fm = torch.randn(P, 16, 32, 31, 31, device=dev)
tdim = cb.transformer_dim(3, 3, 32, True)
cos = [torch.rand(P, 16, 1, 2, device=dev) * 30 for _ in range(iters)]
fts = [torch.randn(P, 16, 1, 32, device=dev) for _ in range(iters)]
out = torch.empty(P, 1, 16, tdim, device=dev)
print("default granularity:", lib.comet_set_l2_fetch_granularity(0))
for gran in (0, 32, 64, 128, 32):
    got = lib.comet_set_l2_fetch_granularity(gran)
    blk = cb.CorrBlock(fm, num_levels=3, radius=3)
    tk = cb.TrackTokenizer(blk, cos[0][:, 0], tdim)
    for i in range(2): tk.tokens(cos[i], fts[i], out=out)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        tk.tokens(cos[i], fts[i], out=out); ev[i + 1].record()
    torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(iters)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); blk2 = cb.CorrBlock(fm, num_levels=3, radius=3); e1.record(); torch.cuda.synchronize()
    print(f"set {gran} -> in effect {got}: fine tokens min {min(ts):.1f} us median {sorted(ts)[len(ts)//2]:.1f} us; "
          f"pyramid {e0.elapsed_time(e1)*1e3:.1f} us", flush=True)
