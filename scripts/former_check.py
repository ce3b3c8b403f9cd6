"""GPU bring-up of the update-transformer kernels: each piece against float64 torch, next to torch's own float32 error."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import comet_pose_estimation_b200 as cb
from comet_pose_estimation_b200 import update_former_tc as tc, update_former as uf

dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
g = torch.Generator(device=dev).manual_seed(0)


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


def check_linear(M, K, N, np_, gelu=False, resid=False):
    x = torch.randn(M, K, device=dev, generator=g)
    w = torch.randn(N, K, device=dev, generator=g) / K ** 0.5
    b = torch.randn(N, device=dev, generator=g)
    r = torch.randn(M, N, device=dev, generator=g) if resid else None
    run = tc._Run(tc._Weights(), np_, dev)
    xp = run.split(x)
    out, op = run.linear(xp, w, b, resid=r, gelu=gelu, want_planes=True)
    torch.cuda.synchronize()
    ref = x.double() @ w.double().T + b.double()
    if gelu:
        ref = torch.nn.functional.gelu(ref)
    if resid:
        ref = ref + r.double()
    t32 = x @ w.T + b
    if gelu:
        t32 = torch.nn.functional.gelu(t32)
    if resid:
        t32 = t32 + r
    planes = op.float().sum(0)[:, :N]
    print(f"linear M={M} K={K} N={N} np={np_} gelu={gelu} resid={resid}: ours {rel(out, ref):.2e}  planes {rel(planes, ref):.2e}  torch-fp32 {rel(t32, ref):.2e}")


for np_ in (3, 1):
    check_linear(128, 64, 128, np_)
    check_linear(300, 664, 384, np_)
    check_linear(1000, 384, 130, np_, resid=True)
    check_linear(9216, 384, 1536, np_, gelu=True)
    check_linear(9216, 1536, 384, np_, resid=True)
    check_linear(77, 216, 34, np_)

# LayerNorm + attention
x = torch.randn(777, 384, device=dev, generator=g) * 3 + 1
ln = torch.nn.LayerNorm(384, elementwise_affine=False, eps=1e-6).to(dev)
run = tc._Run(tc._Weights(), 3, dev)
o, p = run.layernorm(x, ln, True)
print("layernorm", rel(o, ln.double()(x.double())), "planes", rel(p.float().sum(0), ln.double()(x.double())), "torch", rel(ln.float()(x), ln.double()(x.double())))
lna = torch.nn.LayerNorm(384).to(dev)
lna.weight.data.normal_(generator=g); lna.bias.data.normal_(generator=g)
o, p = run.layernorm(x, lna, True)
print("layernorm affine", rel(o, lna.double()(x.double())))
lna.float()

for (Bq, H, Lq, Lk, dh) in ((40, 8, 16, 16, 48), (16, 8, 64, 512, 48), (16, 8, 512, 64, 48), (3, 8, 5, 7, 4), (16, 8, 64, 64, 32)):
    D = H * dh
    q = torch.randn(Bq, Lq, D, device=dev, generator=g); k = torch.randn(Bq, Lk, D, device=dev, generator=g); v = torch.randn(Bq, Lk, D, device=dev, generator=g)
    op = run.attention(q.view(-1, D), k.view(-1, D), v.view(-1, D), Bq, H, Lq, Lk, dh, Lq * D, D, Lk * D, D, Bq * Lq, D, Lq * D, D)
    qq = q.double().view(Bq, Lq, H, dh).transpose(1, 2); kk = k.double().view(Bq, Lk, H, dh).transpose(1, 2); vv = v.double().view(Bq, Lk, H, dh).transpose(1, 2)
    ref = torch.softmax(qq @ kk.transpose(-1, -2) / dh ** 0.5, -1) @ vv
    ref = ref.transpose(1, 2).reshape(Bq * Lq, D)
    print(f"attention B={Bq} H={H} Lq={Lq} Lk={Lk} dh={dh}:", rel(op.float().sum(0)[:, :D], ref))

# whole module
for (name, kw, B, N, T) in (("tiny", dict(space_depth=1, time_depth=1, input_dim=160, hidden_size=32, output_dim=18), 1, 7, 4),
                             ("coarse", dict(space_depth=6, time_depth=6, input_dim=664, hidden_size=384, output_dim=130), 1, 512, 16),
                             ("fine", dict(space_depth=0, time_depth=4, input_dim=216, hidden_size=256, output_dim=34, add_space_attn=False), 512, 1, 16)):
    torch.manual_seed(1)
    m = uf.EfficientUpdateFormer(**kw).to(dev).eval()
    x = torch.randn(B, N, T, kw["input_dim"], device=dev, generator=g)
    with torch.no_grad():
        ours = m(x)
        uf.USE_TC_KERNELS = False
        t32 = m(x)
        ref = m.double()(x.double())
        m.float()
        uf.USE_TC_KERNELS = True
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ours_bf = m(x)
            uf.USE_TC_KERNELS = False
            t_bf = m(x)
            uf.USE_TC_KERNELS = True
        torch.cuda.synchronize()
        def timeit(fn, n=5):
            fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(n): fn()
            torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
        ms_ours = timeit(lambda: m(x))
        uf.USE_TC_KERNELS = False
        ms_torch = timeit(lambda: m(x))
        uf.USE_TC_KERNELS = True
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ms_ours_bf = timeit(lambda: m(x))
            uf.USE_TC_KERNELS = False
            ms_torch_bf = timeit(lambda: m(x))
            uf.USE_TC_KERNELS = True
    print(f"former {name}: fp32 ours {rel(ours, ref):.2e} torch {rel(t32, ref):.2e} | bf16 ours {rel(ours_bf, ref):.2e} torch-autocast {rel(t_bf.float(), ref):.2e}"
          f" | ms fp32 ours {ms_ours:.2f} torch {ms_torch:.2f} | bf16 ours {ms_ours_bf:.2f} torch {ms_torch_bf:.2f}")
