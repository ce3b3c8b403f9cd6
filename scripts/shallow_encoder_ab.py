"""A/B of the patch encoder for one sequence (512 tracks x 16 frames = 8192 patches of 31x31 from 512x512 frames):
the fused one-kernel encoder (csrc/shallow_encoder.cu, gather included) against the per-operator path (gather kernel +
cuDNN float32 / TF32 convolutions + the library's norm / resize kernels).  Prints ms per sequence and the difference."""
import sys

import torch

sys.path.insert(0, ".")
import importlib  # noqa: E402

rt = importlib.import_module("comet_pose_estimation_b200.refine_track")

dev = torch.device("cuda:0")
torch.manual_seed(0)
B, S, N, H, W = 1, 16, 512, 512, 512
images = torch.rand(B, S, 3, H, W, device=dev)
tl = torch.randint(0, H - 31 + 1, (B, S, N, 2), device=dev, dtype=torch.int32)
fnet = rt.ShallowEncoder(3).eval().to(dev).to(memory_format=torch.channels_last)


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


with torch.no_grad():
    spin = torch.rand(8192, 8192, device=dev)
    for _ in range(200):           # bring the clocks up before the first measurement
        spin @ spin
    torch.cuda.synchronize()
    t_fused, a = timed(lambda: fnet.encode_patches_of(images, tl)[0])
    t_fused_p, a2 = timed(lambda: fnet(rt.extract_patches(images, tl, 31), defer_upsample=True)[0])
    res = {"fused_from_images_ms": t_fused, "fused_gather_then_encode_ms": t_fused_p, "bit_equal": bool(torch.equal(a, a2))}
    rt.USE_FUSED_ENCODER = False
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        t, b = timed(lambda: fnet(rt.extract_patches(images, tl, 31), defer_upsample=True)[0])
        res[f"per_operator_tf32_{tf32}_ms"] = t
        res[f"max_abs_diff_tf32_{tf32}"] = float((a - b).abs().max())
    rt.USE_FUSED_ENCODER = True
    res["fused_from_images_again_ms"] = timed(lambda: fnet.encode_patches_of(images, tl)[0])[0]
    res["gather_ms"] = timed(lambda: rt.extract_patches(images, tl, 31))[0]
    res["out_absmax"] = float(a.abs().max())
print(res)
