"""Two launches of the fused patch encoder at the shipped shape (8192 patches from 16 frames of 512x512) -- the target of
the ncu capture (scripts/gpu_shallow_encoder.sh) -- plus a CUDA-event timing of 50 launches when run plainly."""
import importlib
import sys

import torch

sys.path.insert(0, ".")
rt = importlib.import_module("comet_pose_estimation_b200.refine_track")

dev = torch.device("cuda:0")
torch.manual_seed(0)
B, S, N, H, W = 1, 16, 512, 512, 512
images = torch.rand(B, S, 3, H, W, device=dev)
tl = torch.randint(0, H - 31 + 1, (B, S, N, 2), device=dev, dtype=torch.int32)
fnet = rt.ShallowEncoder(3).eval().to(dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50
with torch.no_grad():
    for _ in range(2):
        fnet.encode_patches_of(images, tl)
    torch.cuda.synchronize()
    if n > 0:
        for _ in range(100):
            fnet.encode_patches_of(images, tl)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fnet.encode_patches_of(images, tl)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        macs = 8192 * 2.0398e6
        print({"ms_per_sequence": ms, "fp32_fma_per_s": macs / ms * 1e3, "frac_of_fp32_peak": macs / ms * 1e3 / (148 * 128 * 1.965e9)})
