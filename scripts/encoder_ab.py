"""One forward of the fine tracker's patch encoder (8192 patches, channels-last) with the library kernels and one with
the ATen ops (F.interpolate / InstanceNorm2d) -- for an ncu launch list (profiles/r01d_encoder_launches.csv)."""
import importlib, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rt = importlib.import_module("comet_pose_estimation_b200.refine_track")
dev = torch.device("cuda:0"); torch.manual_seed(0)
x = torch.rand(8192, 3, 31, 31, device=dev).contiguous(memory_format=torch.channels_last)
net = rt.ShallowEncoder(3).eval().to(dev).to(memory_format=torch.channels_last)
with torch.no_grad():
    for flag in (True, False):
        rt.USE_LIBRARY_KERNELS = flag
        net(x); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); net(x); e1.record(); torch.cuda.synchronize()
        print(f"library kernels={flag}: {e0.elapsed_time(e1):.2f} ms")
