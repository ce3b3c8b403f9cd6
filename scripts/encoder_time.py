import sys, time, torch, importlib
sys.path.insert(0, "/root/repo")
import comet_pose_estimation_b200 as cb
rt = importlib.import_module("comet_pose_estimation_b200.refine_track")
dev = torch.device("cuda:0")
torch.manual_seed(0)
enc = cb.BasicEncoder().eval().to(dev)
x = torch.rand(16, 3, 256, 256, device=dev)
def timeit(fn, n=5):
    fn(); fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
for cl in (False, True):
    e = enc.to(memory_format=torch.channels_last) if cl else enc
    xi = x.contiguous(memory_format=torch.channels_last) if cl else x
    for lib in (True, False):
        rt.USE_LIBRARY_KERNELS = lib
        with torch.no_grad():
            print("channels_last", cl, "library kernels", lib, "fp32 ms", round(timeit(lambda: e(xi)), 2))
            with torch.autocast("cuda", dtype=torch.bfloat16):
                print("   autocast ms", round(timeit(lambda: e(xi)), 2))
rt.USE_LIBRARY_KERNELS = True
torch.backends.cudnn.benchmark = True
with torch.no_grad():
    e = enc.to(memory_format=torch.channels_last); xi = x.contiguous(memory_format=torch.channels_last)
    print("cudnn.benchmark channels_last library fp32 ms", round(timeit(lambda: e(xi)), 2))
from torch.profiler import profile, ProfilerActivity
torch.backends.cudnn.benchmark = False
with torch.no_grad(), profile(activities=[ProfilerActivity.CUDA]) as prof:
    e(xi); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=60))
