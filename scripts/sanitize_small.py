"""Small invocations of every kernel added in round 2 (for `compute-sanitizer --tool memcheck python scripts/sanitize_small.py`)."""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import comet_pose_estimation_b200 as cb
from comet_pose_estimation_b200 import update_former as uf, update_former_tc as tc

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
tc.USE_CUDA_GRAPH = False
# fine tracker on the source map (up2 lookup + pyramid), odd batch, queries on / off the map
src = torch.randn(5, 3, 16, 16, 32, device=dev, generator=g).permute(0, 1, 4, 2, 3)
co = torch.rand(5, 3, 1, 2, device=dev, generator=g) * 40 - 5
ft = torch.randn(5, 3, 1, 32, device=dev, generator=g)
blk = cb.CorrBlock.from_upsampled(src, num_levels=3, radius=3)
blk.corr(ft)
a = blk.sample(co)
tk = cb.TrackTokenizer(blk, co[:, 0], cb.transformer_dim(3, 3, 32, True))
x = tk.tokens(co, ft)
with torch.autocast("cuda", dtype=torch.bfloat16):
    tk.tokens(co, ft)
# coarse tensor path (PDL chain), N not a multiple of 128
fm = torch.randn(1, 2, 128, 64, 64, device=dev, generator=g)
c2 = torch.rand(1, 2, 130, 2, device=dev, generator=g) * 70 - 3
f2 = torch.randn(1, 2, 130, 128, device=dev, generator=g)
b2 = cb.CorrBlock(fm, num_levels=5, radius=4)
t2 = cb.TrackTokenizer(b2, c2[:, 0], cb.transformer_dim(5, 4, 128, False))
t2.tokens(c2, f2)
# transformer: every GEMM tile width, both precisions, all three attention kernels
for kw, B, N, T in ((dict(space_depth=1, time_depth=1, input_dim=160, hidden_size=32, output_dim=18), 1, 7, 4),
                    (dict(space_depth=1, time_depth=1, input_dim=664, hidden_size=384, output_dim=130), 1, 70, 5),
                    (dict(space_depth=0, time_depth=1, input_dim=216, hidden_size=256, output_dim=34, add_space_attn=False), 9, 1, 16)):
    m = uf.EfficientUpdateFormer(**kw).to(dev).eval()
    xx = torch.randn(B, N, T, kw["input_dim"], device=dev, generator=g)
    with torch.no_grad():
        y = m(xx)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            m(xx)
# long-key attention (tiled kernel with several key tiles) and few-key / many-query attention
run = tc._Run(tc._Weights(), 3, dev)
for (Bq, H, Lq, Lk, dh) in ((2, 4, 70, 130, 64), (2, 8, 200, 64, 48), (3, 8, 16, 16, 48)):
    D = H * dh
    q = torch.randn(Bq, Lq, D, device=dev, generator=g); k = torch.randn(Bq, Lk, D, device=dev, generator=g); v = torch.randn(Bq, Lk, D, device=dev, generator=g)
    run.attention(q.view(-1, D), k.view(-1, D), v.view(-1, D), Bq, H, Lq, Lk, dh, Lq * D, D, Lk * D, D, Bq * Lq, D, Lq * D, D)
torch.cuda.synchronize()
print("sanitize_small: ok", float(a.abs().max()), float(x.abs().max()), float(y.abs().max()))
