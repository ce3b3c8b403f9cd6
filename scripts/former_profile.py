"""One eager forward of the coarse / fine update transformer on the sm_100a kernels (for an ncu launch list):
python scripts/former_profile.py [coarse|fine] [np]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from comet_pose_estimation_b200 import update_former_tc as tc, update_former as uf

which = sys.argv[1] if len(sys.argv) > 1 else "coarse"
np_ = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
kw, B, N, T = {"coarse": (dict(space_depth=6, time_depth=6, input_dim=664, hidden_size=384, output_dim=130), 1, 512, 16),
               "fine": (dict(space_depth=0, time_depth=4, input_dim=216, hidden_size=256, output_dim=34, add_space_attn=False), 512, 1, 16)}[which]
torch.manual_seed(1)
m = uf.EfficientUpdateFormer(**kw).to(dev).eval()
x = torch.randn(B, N, T, kw["input_dim"], device=dev)
tc.USE_CUDA_GRAPH = False
with torch.no_grad():
    for _ in range(3):
        tc.forward(m, x, np_)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tc.forward(m, x, np_)
    e1.record()
    torch.cuda.synchronize()
print(which, "np", np_, "eager forward ms", e0.elapsed_time(e1))
