"""clock64 timeline of CTA 0 of the coarse tcgen05 kernel (needs a trace build:
python -m comet_pose_estimation_b200.build --trace; rebuild without the flag afterwards).
python scripts/tc_trace.py [batch]   -- per job: tiles, MMA / epilogue spans, stager target staging and window units."""
import contextlib, ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import comet_pose_estimation_b200 as cb
lib = cb._lib.lib
lib.comet_tc_debug_stamps.argtypes = [ctypes.c_void_p]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda:0")
tdim = cb.transformer_dim(5, 4, 128, False)
torch.manual_seed(0)
fm = torch.randn(B, 16, 128, 64, 64, device=dev); ft = torch.randn(B, 16, 512, 128, device=dev)
co = torch.rand(B, 16, 512, 2, device=dev) * 63
b3 = cb.CorrBlock(fm, num_levels=5, radius=4); t3 = cb.TrackTokenizer(b3, co[:, 0], tdim)
out = torch.empty(B, 512, 16, tdim, device=dev)
ctx = torch.autocast('cuda', dtype=torch.bfloat16) if os.environ.get('TC_BF16') else contextlib.nullcontext()
with ctx:
    for _ in range(3): t3.tokens(co, ft, out=out)
    buf = torch.zeros(6 * 64 * 2, dtype=torch.int64, device=dev)
    lib.comet_tc_debug_stamps(buf.data_ptr())
    t3.tokens(co, ft, out=out); torch.cuda.synchronize()
    lib.comet_tc_debug_stamps(None)
s = buf.cpu().view(6, 64, 2)
times = torch.cat([s[:3].flatten(), s[3:5].reshape(8, 16, 2)[:, :14].flatten()])
t0 = int(times[times > 0].min())
rel = lambda x: int(x) - t0 if int(x) else -1
print("batch", B, "debug", os.environ.get("COMET_TC_DEBUG", "0"), "bf16" if os.environ.get('TC_BF16') else "fp32")
st = s[3:5].reshape(8, 16, 2)
tile = 0
for j in range(8):
    nt, nseg = int(st[j, 15, 0]), int(st[j, 15, 1])
    if nt == 0: break
    lo, hi = tile, min(tile + nt, 64) - 1
    print(f"job {j}: tiles {lo}..{tile + nt - 1} ({nseg} unit(s))")
    if lo <= hi:
        print(f"   producer first issue {rel(s[0, lo, 0])}   mma first start {rel(s[1, lo, 0])}  last end {rel(s[1, hi, 1])}"
              f"   epilogue first got {rel(s[2, lo, 0])}  last released {rel(s[2, hi, 1])}")
    print(f"   stager: targets of the next job [{rel(st[j, 0, 0])}, {rel(st[j, 0, 1])}]  units (got, half 0 stored, next half 0 loads issued, stored): "
          + "  ".join(f"({rel(st[j, 1 + 2 * u, 0])}, {rel(st[j, 1 + 2 * u, 1])}, {rel(st[j, 2 + 2 * u, 0])}, {rel(st[j, 2 + 2 * u, 1])})" for u in range(nseg)))
    print(f"   stager: pos loads: before {rel(st[j, 12, 0])}  issued {rel(st[j, 12, 1])}  (debug 512: all landed) {rel(st[j, 13, 0])}")
    tile += nt
print("epilogue units (warp 2): arrives at the window claim, claimed, handed to the stager")
for u in range(32):
    if int(s[5, 2 * u, 0]) == 0: break
    print(f"  unit {u}: {rel(s[5, 2 * u, 0])} {rel(s[5, 2 * u, 1])} {rel(s[5, 2 * u + 1, 0])}")
print("tile  prod_issue | mma_start mma_end | epi_got epi_rel")
for i in range(64):
    r = [s[0, i, 0], s[1, i, 0], s[1, i, 1], s[2, i, 0], s[2, i, 1]]
    print(f"{i:3d}  " + "  ".join(f"{rel(x):8d}" for x in r))
