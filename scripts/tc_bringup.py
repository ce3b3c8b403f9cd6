"""Bring-up diagnostics for the tcgen05 kernel: prepare -> volume (pure GEMM) -> lookup -> tokens, each compared
with the SIMT kernels / torch on the same inputs.  Run under `timeout` on the GPU box."""
import os
import sys
import ctypes

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import comet_pose_estimation_b200 as cb
from comet_pose_estimation_b200 import _lib
from comet_pose_estimation_b200.blocks import _Pyramid

lib = _lib.lib
dev = torch.device("cuda:0")
print("tensor path:", lib.comet_has_tensor_path(), flush=True)


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def stage(name):
    print(f"--- {name}", flush=True)


S_, N_ = int(os.environ.get("TC_S", 2)), int(os.environ.get("TC_N", 200))
g = torch.Generator(device=dev).manual_seed(0)
fmaps = torch.randn(1, S_, 128, 64, 64, device=dev, generator=g)
feats = torch.randn(1, S_, N_, 128, device=dev, generator=g)
coords = torch.rand(1, S_, N_, 2, device=dev, generator=g) * 70 - 3
coords[0, 0, 0] = torch.tensor([0.0, 0.0])
coords[0, 0, 1] = torch.tensor([63.0, 63.0])
coords[0, 0, 2] = torch.tensor([-20.0, 30.0])
coords[0, 0, 3] = torch.tensor([31.5, 90.0])

stage("prepare")
pyr = _Pyramid(fmaps, 5)
torch.cuda.synchronize()
assert pyr.split is not None
sp = pyr.split.view(2, S_, 128, 5504).float()
rec = sp[0] + sp[1]
want0 = fmaps[0].reshape(S_, 128, 4096)
print("level0 hi+lo vs f32:", rel(rec[..., :4096], want0))
off = [0, 4096, 5120, 5376, 5440]
for l in range(1, 5):
    w = pyr.levels[l][0].reshape(S_, 128, -1)
    print(f"level{l} split vs f32 pyramid:", rel(rec[..., off[l]:off[l] + w.shape[-1]], w),
          " pyr vs avg_pool:", rel(w, torch.nn.functional.avg_pool2d(pyr.levels[l - 1][0], 2).reshape(S_, 128, -1)))
print("pad zeros:", float(sp[..., 5456:].abs().max()))

stage("volume (pure tcgen05 GEMM)")
blk = cb.CorrBlock(fmaps, num_levels=5, radius=4)
blk.corr(feats)
vols = blk.corrs_pyramid
torch.cuda.synchronize()
print("tc status:", lib.comet_tc_status())
for l, v in enumerate(vols):
    f = blk.fmaps_pyramid[l]
    want = torch.matmul(feats.double(), f.reshape(1, S_, 128, -1).double()) / np.sqrt(128.0)
    want = want.reshape(v.shape).float()
    e = rel(v, want)
    print(f"level {l}: shape {tuple(v.shape)} rel err {e:.3e}")
    if e > 1e-4:
        d = (v - want).abs()[0, 0]
        bad = (d > 1e-3 * want.abs().max()).float()
        print("   bad fraction:", float(bad.mean()), " bad by query (first 8):", bad.mean(dim=(1, 2))[:8].tolist())
        print("   bad by row (first 8):", bad.mean(dim=(0, 2))[:8].tolist(), " by col (first 16):", bad.mean(dim=(0, 1))[:16].tolist())
        print("   v[0,:4,:4] got:", v[0, 0, 0, :2, :6].tolist(), " want:", want[0, 0, 0, :2, :6].tolist())

stage("lookup vs SIMT")


def simt_lookup():
    out = torch.empty(1, S_, N_, 405, device=dev)
    _lib.check(lib.comet_corr_lookup_f32(pyr.fmaps0.data_ptr(), pyr.pyr.data_ptr(), feats.data_ptr(), *feats.stride()[:3], 0,
                                         coords.data_ptr(), *coords.stride()[:3], out.data_ptr(), *out.stride()[:3],
                                         1, S_, N_, 128, 64, 64, 5, 4, 0, 0, pyr.layout, torch.cuda.current_stream().cuda_stream))
    return out


ref = simt_lookup()
got = blk.sample(coords)
torch.cuda.synchronize()
print("tc status:", lib.comet_tc_status())
e = rel(got, ref)
print("lookup rel err:", e)
if e > 1e-4:
    d = (got - ref).abs()
    for l in range(5):
        dl = d[..., l * 81:(l + 1) * 81]
        print(f"  level {l}: max {float(dl.max()):.3e}; worst query {int(dl.amax(-1).flatten().argmax())}",
              " per-j max:", dl.reshape(-1, 9, 9).amax(dim=(0, 1)).tolist())
    q = int(d.amax(-1).flatten().argmax())
    print("  worst query coords:", coords.reshape(-1, 2)[q].tolist())

stage("tokens vs SIMT")
tdim = cb.transformer_dim(5, 4, 128, False)
tok = cb.TrackTokenizer(blk, coords[:, 0], tdim)
x = tok.tokens(coords, feats)
os.environ["COMET_B200_DISABLE_TC"] = "1"
blk2 = cb.CorrBlock(fmaps, num_levels=5, radius=4)
tok2 = cb.TrackTokenizer(blk2, coords[:, 0], tdim)
x2 = tok2.tokens(coords, feats)
os.environ["COMET_B200_DISABLE_TC"] = "0"
torch.cuda.synchronize()
print("tc status:", lib.comet_tc_status())
print("tokens rel err:", rel(x, x2), " segments:", [rel(x[..., a:b], x2[..., a:b]) for a, b in ((0, 130), (130, 535), (535, 664))])

stage("bf16 autocast mode")
with torch.autocast("cuda", dtype=torch.bfloat16):
    gb = blk.sample(coords)
    os.environ["COMET_B200_DISABLE_TC"] = "1"
    rb = cb.CorrBlock(fmaps, num_levels=5, radius=4)
    rb.corr(feats)
    rb = rb.sample(coords)
    os.environ["COMET_B200_DISABLE_TC"] = "0"
print("bf16 lookup rel err vs SIMT bf16:", rel(gb, rb), " vs fp32:", rel(gb, ref))

stage("timing (S=16, N=512)")
fm = torch.randn(1, 16, 128, 64, 64, device=dev)
ft = torch.randn(1, 16, 512, 128, device=dev)
co = torch.rand(1, 16, 512, 2, device=dev) * 63
b3 = cb.CorrBlock(fm, num_levels=5, radius=4)
t3 = cb.TrackTokenizer(b3, co[:, 0], tdim)
out = torch.empty(1, 512, 16, tdim, device=dev)
for _ in range(3):
    t3.tokens(co, ft, out=out)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    t3.tokens(co, ft, out=out)
e1.record()
torch.cuda.synchronize()
print("tc tokens kernel: %.1f us/iter" % (e0.elapsed_time(e1) / 20 * 1e3))
print("DONE")
