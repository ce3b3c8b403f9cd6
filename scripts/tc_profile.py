"""Times the tensor-core token kernel at the coarse config (Q sequences x S=16 x N=512) with CUDA events."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import comet_pose_estimation_b200 as cb

Q = int(os.environ.get("TC_Q", 1)); iters = int(os.environ.get("TC_ITERS", 10))
dev = torch.device("cuda:0")
tdim = cb.transformer_dim(5, 4, 128, False)
fm = torch.randn(Q, 16, 128, 64, 64, device=dev)
ft = torch.randn(Q, 16, 512, 128, device=dev)
co = torch.rand(Q, 16, 512, 2, device=dev) * 63
b3 = cb.CorrBlock(fm, num_levels=5, radius=4)
t3 = cb.TrackTokenizer(b3, co[:, 0], tdim)
out = torch.empty(Q, 512, 16, tdim, device=dev)
import contextlib
ctx = torch.autocast('cuda', dtype=torch.bfloat16) if os.environ.get('TC_BF16') else contextlib.nullcontext()
ctx.__enter__()
for _ in range(3):
    t3.tokens(co, ft, out=out)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
ev[0].record()
for i in range(iters):
    t3.tokens(co, ft, out=out)
    ev[i + 1].record()
torch.cuda.synchronize()
ts = [ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(iters)]
print(f"Q={Q} debug={os.environ.get('COMET_TC_DEBUG','0')} tokens: min {min(ts):.1f} us, median {sorted(ts)[len(ts)//2]:.1f} us", end="")
# GPU-only time: the same call captured in a CUDA graph (no Python / launch overhead between kernels)
g = torch.cuda.CUDAGraph()
st = torch.cuda.Stream()
st.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(st):
    t3.tokens(co, ft, out=out)
    with torch.cuda.graph(g, stream=st):
        for _ in range(iters):
            t3.tokens(co, ft, out=out)
torch.cuda.synchronize()
g.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
print(f"  | graph: {e0.elapsed_time(e1)*1e3/iters:.1f} us/iter")
b3.corr(ft)
for _ in range(2): b3.sample(co)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); b3.sample(co); e1.record(); torch.cuda.synchronize()
print(f"   lookup-only: {e0.elapsed_time(e1)*1e3:.1f} us")
