import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import comet_pose_estimation_b200 as cb
lib = cb._lib.lib   # needs a trace build: python -m comet_pose_estimation_b200.build --trace
import ctypes
lib.comet_tc_debug_stamps.argtypes = [ctypes.c_void_p]
dev = torch.device("cuda:0")
tdim = cb.transformer_dim(5, 4, 128, False)
fm = torch.randn(1, 16, 128, 64, 64, device=dev); ft = torch.randn(1, 16, 512, 128, device=dev)
co = torch.rand(1, 16, 512, 2, device=dev) * 63
b3 = cb.CorrBlock(fm, num_levels=5, radius=4); t3 = cb.TrackTokenizer(b3, co[:, 0], tdim)
out = torch.empty(1, 512, 16, tdim, device=dev)
import contextlib
ctx = torch.autocast('cuda', dtype=torch.bfloat16) if os.environ.get('TC_BF16') else contextlib.nullcontext()
ctx.__enter__()
for _ in range(3): t3.tokens(co, ft, out=out)
buf = torch.zeros(4 * 64 * 2, dtype=torch.int64, device=dev)
lib.comet_tc_debug_stamps(buf.data_ptr())
t3.tokens(co, ft, out=out); torch.cuda.synchronize()
lib.comet_tc_debug_stamps(None)
s = buf.cpu().view(4, 64, 2)
t0 = int(s[s > 0].min())
print("debug", os.environ.get("COMET_TC_DEBUG", "0"))
print("tile  prod_issue | mma_start mma_end | epi_full epi_rel   (clk since first event)")
print('stager per job: [A start, A done] [rest done] [last WU got, last WU done]')
for j in range(8):
    print("stager job", j, [int(x) - t0 if x else -1 for x in s[3, j*4:(j+1)*4].flatten().tolist()])
for i in list(range(0, 30)):
    r = [int(s[0, i, 0]), int(s[1, i, 0]), int(s[1, i, 1]), int(s[2, i, 0]), int(s[2, i, 1])]
    print(f"{i:3d}  " + "  ".join(f"{(x - t0) if x else -1:8d}" for x in r))

