"""BasicEncoder (cuDNN convolutions + the library's norm / resize kernels) on 16 frames of 512x512: ms per sequence in
float32 (TF32 off / on), under autocast, NCHW and channels-last, and a per-kernel-class breakdown from the profiler."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
tp = importlib.import_module("comet_pose_estimation_b200.track_predictor")

dev = torch.device("cuda:0")
torch.manual_seed(0)
net = tp.BasicEncoder().eval().to(dev)
x = torch.rand(16, 3, 512, 512, device=dev)


def timed(fn, n=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / n, 2)


res = {}
with torch.no_grad():
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        res[f"fp32_nchw_tf32_{tf32}"] = timed(lambda: net(x))
    with torch.autocast("cuda", dtype=torch.bfloat16):
        res["autocast_nchw"] = timed(lambda: net(x))
    net_cl = tp.BasicEncoder().eval().to(dev).to(memory_format=torch.channels_last)
    x_cl = x.contiguous(memory_format=torch.channels_last)
    res["fp32_cl_tf32_True"] = timed(lambda: net_cl(x_cl))
    with torch.autocast("cuda", dtype=torch.bfloat16):
        res["autocast_cl"] = timed(lambda: net_cl(x_cl))
    print(res)
    from torch.profiler import profile, ProfilerActivity
    for tag, fn in (("autocast_nchw", lambda: net(x)), ("autocast_cl", lambda: net_cl(x_cl))):
        with torch.autocast("cuda", dtype=torch.bfloat16), profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn()
            torch.cuda.synchronize()
        rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)[:8]
        print(tag, [(r.key[:48], r.count, round(r.device_time_total / 1e3, 2)) for r in rows])
