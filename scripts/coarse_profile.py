"""Kernel-level profile of the drop-in coarse BaseTrackerPredictor (B=1, S=16, N=512, 64x64 maps, 4 iterations)."""
import os, sys
from types import SimpleNamespace as NS
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import comet_pose_estimation_b200 as cb
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0"); torch.manual_seed(0)
cfg = NS(track_conf=False, MODEL=NS(TRACK=NS(efficient_corr=False)))
m = cb.BaseTrackerPredictor(cfg=cfg).eval().to(dev)          # defaults: stride 4, 5 levels, r=4, latent 128, hidden 384, depth 6
fmaps = torch.randn(1, 16, 128, 64, 64, device=dev)
q = torch.rand(1, 512, 2, device=dev) * 480 + 16
with torch.no_grad():
    for _ in range(2): m(query_points=q, fmaps=fmaps, iters=4, down_ratio=2, return_feat=True, TRACKorPOSE=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); m(query_points=q, fmaps=fmaps, iters=4, down_ratio=2, return_feat=True, TRACKorPOSE=False); e1.record()
    torch.cuda.synchronize(); print(f"coarse predictor, 4 iterations: {e0.elapsed_time(e1):.2f} ms")
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        m(query_points=q, fmaps=fmaps, iters=4, down_ratio=2, return_feat=True, TRACKorPOSE=False); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=64))
