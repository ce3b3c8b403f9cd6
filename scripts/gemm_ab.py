import os, sys, time, torch
sys.path.insert(0, "/root/repo")
import comet_pose_estimation_b200 as cb
from comet_pose_estimation_b200 import update_former_tc as tc, update_former as uf, _lib
dev = torch.device("cuda:0")
torch.manual_seed(1)
kw = dict(space_depth=6, time_depth=6, input_dim=664, hidden_size=384, output_dim=130)
m = uf.EfficientUpdateFormer(**kw).to(dev).eval()
x = torch.randn(1, 512, 16, 664, device=dev)
def rel(a, b): return float((a.double() - b.double()).abs().max() / b.double().abs().max())
with torch.no_grad():
    ref = m.double()._forward_torch(x.double()); m.float()
    for bk32 in (True, False):
        _lib.set_option(_lib.OPT_GEMM_BK32, bk32)
        m._tc_graphs = {}
        y = m(x); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10): m(x)
        torch.cuda.synchronize()
        print("BK32" if bk32 else "BK64", "coarse fp32 forward ms", (time.perf_counter() - t0) / 10 * 1e3, "err vs f64", rel(y, ref))
