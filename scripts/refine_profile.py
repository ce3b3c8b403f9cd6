"""Times the drop-in refine_track at the full configuration (B=1, S=16, N=512, 512x512 images): patch gather,
ShallowEncoder (torch / cuDNN, channels-last), fine tracker hot path (our kernels) with a tiny update transformer."""
import os, sys, time
from types import SimpleNamespace as NS
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import comet_pose_estimation_b200 as cb
from comet_pose_estimation_b200.refine_track import ShallowEncoder, extract_patches, refine_track

dev = torch.device("cuda:0")
torch.manual_seed(0)
cfg = NS(track_conf=False, MODEL=NS(TRACK=NS(efficient_corr=False)))
fnet = ShallowEncoder(3).eval().to(dev).to(memory_format=torch.channels_last)
ftr = cb.BaseTrackerPredictor(stride=1, corr_levels=3, corr_radius=3, latent_dim=32, hidden_size=384, depth=4,
                              use_spaceatt=False, fine=True, cfg=cfg).eval().to(dev)
B, S, N, HW = 1, 16, 512, 512
images = torch.rand(B, S, 3, HW, HW, device=dev)
coarse = torch.rand(B, 1, N, 2, device=dev) * (HW - 40) + 20 + torch.randn(B, S, N, 2, device=dev) * 2

def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e

with torch.no_grad():
    for tf32 in (True, False):
        torch.backends.cudnn.allow_tf32 = tf32
        for _ in range(2):
            refine_track(images, fnet, ftr, coarse, compute_score=True)
        torch.cuda.synchronize()
        e0 = ev()
        tl = (coarse.floor().int() - 15).clamp(0, HW - 31)
        p = extract_patches(images, tl, 31); e1 = ev()
        f = fnet(p); e2 = ev()
        torch.cuda.synchronize()
        t0 = ev(); r, sc = refine_track(images, fnet, ftr, coarse, compute_score=True); t1 = ev()
        torch.cuda.synchronize()
        print(f"cudnn tf32={tf32}: patch gather {e0.elapsed_time(e1):.2f} ms, ShallowEncoder {e1.elapsed_time(e2):.2f} ms, "
              f"whole refine_track (6 it, depth-4 time-attention transformer) {t0.elapsed_time(t1):.2f} ms", flush=True)

from torch.profiler import profile, ProfilerActivity
with torch.no_grad(), profile(activities=[ProfilerActivity.CUDA]) as prof:
    refine_track(images, fnet, ftr, coarse, compute_score=True); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))
