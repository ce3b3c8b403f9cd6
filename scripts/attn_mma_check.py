"""Autocast-mode attention with at most 64 keys: the mma.sync bf16 kernel against float64 and against the float32
lane-per-query kernel, then the forward time of the coarse update transformer with and without it."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from comet_pose_estimation_b200 import update_former_tc as tc, update_former as uf, _lib

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


run = tc._Run(tc._Weights(), 1, dev)
for (Bq, H, Lq, Lk, dh) in ((16, 8, 512, 64, 48), (16, 8, 64, 64, 48), (2, 8, 100, 37, 32), (3, 4, 200, 5, 64), (1, 8, 72, 64, 48), (16, 8, 64, 512, 48), (2, 8, 130, 150, 32), (576, 8, 16, 16, 48), (40, 8, 16, 16, 32), (5, 8, 7, 11, 48)):
    D = H * dh
    q = torch.randn(Bq, Lq, D, device=dev, generator=g); k = torch.randn(Bq, Lk, D, device=dev, generator=g); v = torch.randn(Bq, Lk, D, device=dev, generator=g)
    qq = q.double().view(Bq, Lq, H, dh).transpose(1, 2); kk = k.double().view(Bq, Lk, H, dh).transpose(1, 2); vv = v.double().view(Bq, Lk, H, dh).transpose(1, 2)
    ref = (torch.softmax(qq @ kk.transpose(-1, -2) / dh ** 0.5, -1) @ vv).transpose(1, 2).reshape(Bq * Lq, D)
    errs = []
    for on in (0, 3):
        _lib.check(_lib.lib.comet_set_option(_lib.OPT_ATTN_MMA, on))
        op = run.attention(q.view(-1, D), k.view(-1, D), v.view(-1, D), Bq, H, Lq, Lk, dh, Lq * D, D, Lk * D, D, Bq * Lq, D, Lq * D, D)
        torch.cuda.synchronize()
        errs.append(rel(op.float().sum(0)[:, :D], ref))
    print(f"attention np=1 B={Bq} H={H} Lq={Lq} Lk={Lk} dh={dh}: float32 kernel {errs[0]:.2e}  mma kernel {errs[1]:.2e}", flush=True)

torch.manual_seed(1)
m = uf.EfficientUpdateFormer(space_depth=6, time_depth=6, input_dim=664, hidden_size=384, output_dim=130).to(dev).eval()
x = torch.randn(1, 512, 16, 664, device=dev)
with torch.no_grad():
    ref = m.double()._forward_torch(x.double()); m.float()
    for on in (0, 1, 3):
        _lib.check(_lib.lib.comet_set_option(_lib.OPT_ATTN_MMA, on))
        m._tc_graphs = {}
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = m(x); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                m(x)
            e1.record(); torch.cuda.synchronize()
        print("mma attention", on, "autocast coarse forward ms", round(e0.elapsed_time(e1) / 20, 3), "err vs f64", f"{rel(y, ref):.2e}")
