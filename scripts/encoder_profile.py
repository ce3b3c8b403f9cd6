import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from comet_pose_estimation_b200.refine_track import ShallowEncoder
dev = torch.device("cuda:0"); torch.manual_seed(0)
x = torch.rand(8192, 3, 31, 31, device=dev)
def t(f, inp, n=3):
    with torch.no_grad():
        for _ in range(2): f(inp)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): y = f(inp)
        e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, y
net = ShallowEncoder(3).eval().to(dev)
ms, y = t(net, x); print(f"NCHW: {ms:.2f} ms  out contiguous={y.is_contiguous()}")
net_cl = ShallowEncoder(3).eval().to(dev).to(memory_format=torch.channels_last)
xcl = x.contiguous(memory_format=torch.channels_last)
ms, y = t(net_cl, xcl); print(f"channels_last: {ms:.2f} ms  out CL={y.is_contiguous(memory_format=torch.channels_last)}")
ms, y = t(net, xcl); print(f"NCHW weights, CL input: {ms:.2f} ms  out CL={y.is_contiguous(memory_format=torch.channels_last)}")
from torch.profiler import profile, ProfilerActivity
for name, f, inp in (("NCHW", net, x), ("CL", net_cl, xcl)):
    with torch.no_grad(), profile(activities=[ProfilerActivity.CUDA]) as prof:
        f(inp); torch.cuda.synchronize()
    print(name); print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=8, max_name_column_width=60))
