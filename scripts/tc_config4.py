"""Times the coarse tcgen05 token kernel at BASELINE config 4 (S=64, N=4096 dense grid, B=1) and at the bench shape."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import comet_pose_estimation_b200 as cb
dev = torch.device("cuda:0")
tdim = cb.transformer_dim(5, 4, 128, False)
for (Q, S, N, dense) in ((1, 64, 4096, True), (4, 16, 512, False), (1, 16, 512, False)):
    fm = torch.randn(Q, S, 128, 64, 64, device=dev); ft = torch.randn(Q, S, N, 128, device=dev)
    if dense:
        ii, jj = torch.meshgrid(torch.arange(64, device=dev), torch.arange(64, device=dev), indexing="ij")
        q = torch.stack([jj.flatten() + 0.5, ii.flatten() + 0.5], -1).float()
        co = q[None, None] + torch.randn(Q, S, N, 2, device=dev) * 1.5
    else:
        co = torch.rand(Q, 1, N, 2, device=dev) * 63 + torch.randn(Q, S, N, 2, device=dev) * 1.5
    blk = cb.CorrBlock(fm, num_levels=5, radius=4); tk = cb.TrackTokenizer(blk, co[:, 0], tdim)
    out = torch.empty(Q, N, S, tdim, device=dev)
    for _ in range(3): tk.tokens(co, ft, out=out)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    ev[0].record()
    for i in range(5):
        tk.tokens(co, ft, out=out); ev[i + 1].record()
    torch.cuda.synchronize()
    ms = min(ev[i].elapsed_time(ev[i + 1]) for i in range(5))
    flop = 2.0 * Q * S * N * 128 * 5456
    byts = Q * S * (128 * 4096 * 4 + N * 128 * 4 + N * 8) + Q * N * S * tdim * 4
    print(f"Q={Q} S={S} N={N}: {ms*1e3:.1f} us/iteration  dense-GEMM-equivalent {flop/ms/1e9:.0f} TFLOP/s  "
          f"compulsory bytes {byts/1e6:.0f} MB -> {byts/ms/1e6:.0f} GB/s", flush=True)
    del fm, ft, co, out, blk, tk
