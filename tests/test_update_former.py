"""The track-update transformer on the sm_100a kernels (csrc/gemm_tc.cu, csrc/transformer.cu, update_former_tc.py)
against float64 torch -- each kernel, then the whole ``EfficientUpdateFormer`` (reference: comet/models/track_modules/
blocks.py:205-348, comet/models/modules.py:119-154, :248-344).

Bars.  float32 mode (three bf16 planes per operand): the error against float64 must not exceed 3x the error torch's own
float32 path makes on the same inputs (and 1e-4 in any case) -- "float32-grade", which the tracker loop needs because it
amplifies rounding noise ~200x per iteration (tests/test_full_size.py).  autocast mode (one plane): 2e-2, the bf16 bar
of BASELINE.md section 5, and not worse than 3x torch.autocast(bf16) itself."""
import pytest
import torch

import comet_pose_estimation_b200.update_former as uf

FORMERS = {
    "tiny": (dict(space_depth=1, time_depth=1, input_dim=160, hidden_size=32, output_dim=18), 1, 7, 4),
    "coarse": (dict(space_depth=6, time_depth=6, input_dim=664, hidden_size=384, output_dim=130), 1, 512, 16),
    "coarse_b2_small": (dict(space_depth=2, time_depth=2, input_dim=160, hidden_size=64, output_dim=18), 2, 37, 5),
    "fine": (dict(space_depth=0, time_depth=4, input_dim=216, hidden_size=256, output_dim=34, add_space_attn=False), 512, 1, 16),
}


def rel(a, b):
    return float((a.detach().double() - b.detach().double()).abs().max() / b.detach().double().abs().max())


def test_cpu_tensors_take_the_torch_definition():
    torch.manual_seed(0)
    m = uf.EfficientUpdateFormer(space_depth=1, time_depth=1, input_dim=24, hidden_size=32, output_dim=6).eval()
    x = torch.randn(1, 3, 4, 24)
    with torch.no_grad():
        assert torch.equal(m(x), m._forward_torch(x))


def test_unsupported_head_dim_is_reported():
    from comet_pose_estimation_b200 import update_former_tc as tc

    m = uf.EfficientUpdateFormer(space_depth=0, time_depth=1, input_dim=24, hidden_size=16, output_dim=6, add_space_attn=False)
    assert not tc.supported(m, torch.zeros(1, 1, 2, 24))          # CPU tensor, and head_dim 2


@pytest.fixture
def strict_fp32():
    a = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32 = a


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(128, 64, 128), (300, 664, 384), (1000, 384, 130), (9216, 1536, 384), (77, 216, 34),
                                   (129, 72, 129)])
@pytest.mark.parametrize("np_", [3, 1])
def test_linear_tc_against_float64(strict_fp32, shape, np_):
    from comet_pose_estimation_b200 import update_former_tc as tc

    M, K, N = shape
    g = torch.Generator(device="cuda").manual_seed(M + K + N)
    x = torch.randn(M, K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(N, device="cuda", generator=g)
    r = torch.randn(M, N, device="cuda", generator=g)
    run = tc._Run(tc._Weights(), np_, x.device)
    xp = run.split(x)
    for gelu, resid in ((False, None), (True, None), (False, r), (True, r)):
        out, op = run.linear(xp, w, b, resid=resid, gelu=gelu, want_planes=True)
        ref = x.double() @ w.double().T + b.double()
        t32 = x @ w.T + b
        if gelu:
            ref, t32 = torch.nn.functional.gelu(ref), torch.nn.functional.gelu(t32)
        if resid is not None:
            ref, t32 = ref + r.double(), t32 + r
        bar = max(3 * rel(t32, ref), 2e-7) if np_ == 3 else 2e-2
        assert rel(out, ref) <= bar, (gelu, resid is not None)
        assert rel(op.float().sum(0)[:, :N], ref) <= (bar if np_ == 3 else 2e-2)       # the planes carry the same result
    # weight rows [n0, n1) only (the q / kv halves of MultiheadAttention.in_proj_weight)
    if N >= 32:
        out, _ = run.linear(xp, w, b, 16, N - 8)
        ref = x.double() @ w.double()[16:N - 8].T + b.double()[16:N - 8]
        assert rel(out, ref) <= (2e-6 if np_ == 3 else 2e-2)


@pytest.mark.gpu
def test_layernorm_and_attention_against_float64():
    from comet_pose_estimation_b200 import update_former_tc as tc

    g = torch.Generator(device="cuda").manual_seed(5)
    run = tc._Run(tc._Weights(), 3, torch.device("cuda"))
    for D in (32, 256, 384):
        x = torch.randn(333, D, device="cuda", generator=g) * 3 + 1
        ln = torch.nn.LayerNorm(D, elementwise_affine=False, eps=1e-6).cuda()
        o, p = run.layernorm(x, ln, True)
        ref = torch.nn.functional.layer_norm(x.double(), (D,), eps=1e-6)
        assert rel(o, ref) < 1e-6 and rel(p.float().sum(0), ref) < 1e-6
        lna = torch.nn.LayerNorm(D).cuda()
        lna.weight.data.normal_(generator=g)
        lna.bias.data.normal_(generator=g)
        o, _ = run.layernorm(x, lna, True, want_planes=False)
        assert rel(o, torch.nn.functional.layer_norm(x.double(), (D,), lna.weight.double(), lna.bias.double(), 1e-5)) < 1e-6
    for (B, H, Lq, Lk, dh) in ((40, 8, 16, 16, 48), (16, 8, 64, 512, 48), (16, 8, 512, 64, 48), (3, 8, 5, 7, 4),
                                (5, 8, 64, 64, 32), (2, 4, 70, 130, 64), (64, 8, 64, 4160, 48)):
        D = H * dh
        q = torch.randn(B, Lq, D, device="cuda", generator=g)
        k = torch.randn(B, Lk, D, device="cuda", generator=g) * 2
        v = torch.randn(B, Lk, D, device="cuda", generator=g)
        op = run.attention(q.view(-1, D), k.view(-1, D), v.view(-1, D), B, H, Lq, Lk, dh, Lq * D, D, Lk * D, D, B * Lq, D,
                           Lq * D, D)
        qq, kk, vv = (t.double().view(B, -1, H, dh).transpose(1, 2) for t in (q, k, v))
        ref = (torch.softmax(qq @ kk.transpose(-1, -2) / dh ** 0.5, -1) @ vv).transpose(1, 2).reshape(B * Lq, D)
        assert rel(op.float().sum(0)[:, :D], ref) < (2e-6 if Lk <= 512 else 6e-6), (B, H, Lq, Lk, dh)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(FORMERS))
def test_update_former_against_float64(strict_fp32, name):
    from comet_pose_estimation_b200 import update_former_tc as tc

    kw, B, N, T = FORMERS[name]
    torch.manual_seed(1)
    m = uf.EfficientUpdateFormer(**kw).cuda().eval()
    x = torch.randn(B, N, T, kw["input_dim"], device="cuda")
    with torch.no_grad():
        assert tc.supported(m, x)
        ours = m(x)
        again = m(x)                                   # second call: CUDA-graph replay of the first
        tc.USE_CUDA_GRAPH = False
        eager = m(x)
        tc.USE_CUDA_GRAPH = True
        t32 = m._forward_torch(x)
        ref = m.double()._forward_torch(x.double())
        m.float()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ours_bf = m(x)
            t_bf = m._forward_torch(x).float()
    assert ours.shape == t32.shape and ours.dtype == torch.float32
    assert torch.equal(ours, again) and torch.equal(ours, eager)
    assert rel(ours, ref) <= max(3 * rel(t32, ref), 1e-6)
    assert rel(ours_bf, ref) <= min(2e-2, max(3 * rel(t_bf, ref), 5e-3))


@pytest.mark.gpu
def test_weight_update_invalidates_cached_planes_and_graphs():
    kw, B, N, T = FORMERS["tiny"]
    torch.manual_seed(2)
    m = uf.EfficientUpdateFormer(**kw).cuda().eval()
    x = torch.randn(B, N, T, kw["input_dim"], device="cuda")
    with torch.no_grad():
        a = m(x)
        m.flow_head.weight.mul_(2.0)
        m.flow_head.bias.mul_(2.0)
        b = m(x)
        assert rel(b, 2 * a) < 1e-5
        sd = {k: v.clone() for k, v in m.state_dict().items()}
        sd["input_transform.weight"] = sd["input_transform.weight"] * 0.5
        m.load_state_dict(sd)
        assert rel(m(x), m._forward_torch(x)) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("np_", [3, 1])
def test_gemm_data_path_variants_are_bit_identical(np_):
    """Per-lane stores vs bulk tensor stores from the swizzled staging buffers, 8 vs 16 epilogue warps, 128- vs 96-column
    tiles, single CTAs vs CTA pairs that multicast the W tile (an odd number of row blocks: the last pair has a phantom
    tile): the same sums in the same order, so the results must agree bit for bit."""
    from comet_pose_estimation_b200 import _lib
    from comet_pose_estimation_b200 import update_former_tc as tc

    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(3)
    run = tc._Run(tc._Weights(), np_, dev)
    opts = (_lib.OPT_GEMM_TMA_STORE, _lib.OPT_GEMM_EW16, _lib.OPT_GEMM_BN96, _lib.OPT_GEMM_PAIR)
    saved = [_lib.lib.comet_get_option(o) for o in opts]
    try:
        for (M, K, N, kw) in ((9088, 1536, 384, dict(resid=True)), (9216, 384, 1152, dict()),
                              (9216, 384, 1536, dict(gelu=True, want_planes=True)), (300, 664, 130, dict(want_planes=True))):
            x = torch.randn(M, K, device=dev, generator=g)
            w = torch.randn(N, K, device=dev, generator=g) / K ** 0.5
            b = torch.randn(N, device=dev, generator=g)
            r = torch.randn(M, N, device=dev, generator=g) if kw.get("resid") else None
            kw2 = {k: v for k, v in kw.items() if k != "resid"}
            xp = run.split(x)
            ref = None
            for cfg in ((0, 0, 0, 0), (3, 0, 0, 0), (3, 1, 0, 0), (3, 1, 1, 0), (3, 1, 1, 7), (1, 0, 1, 7), (2, 1, 0, 5)):
                for o, v in zip(opts, cfg):
                    _lib.check(_lib.lib.comet_set_option(o, v))
                out, planes = run.linear(xp, w, b, resid=r, **kw2)
                torch.cuda.synchronize()
                got = (out.clone(), None if planes is None else planes[..., :N].clone())
                if ref is None:
                    ref = got
                else:
                    assert torch.equal(got[0], ref[0]), (M, K, N, cfg)
                    assert ref[1] is None or torch.equal(got[1], ref[1]), (M, K, N, cfg)
    finally:
        for o, v in zip(opts, saved):
            _lib.check(_lib.lib.comet_set_option(o, v))


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(16, 8, 512, 64, 48), (16, 8, 64, 512, 48), (2, 8, 100, 37, 32), (3, 4, 200, 5, 64),
                                   (2, 8, 130, 150, 32), (576, 8, 16, 16, 48), (40, 8, 16, 16, 32), (5, 8, 7, 11, 48)])
def test_autocast_attention_on_tensor_cores(shape):
    """Autocast mode, at least 64 queries or the short time attention: S = q k^T and O = P v as bf16 mma.sync with float32 accumulation and a float32
    softmax on the accumulator registers (what the reference's attention does under torch.autocast) -- within the bf16
    bar of the float64 result, next to the float32 lane-per-query kernels; key counts that are not multiples of 64 and
    the multi-tile online softmax included."""
    from comet_pose_estimation_b200 import _lib
    from comet_pose_estimation_b200 import update_former_tc as tc

    Bq, H, Lq, Lk, dh = shape
    D = H * dh
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(4)
    q = torch.randn(Bq, Lq, D, device=dev, generator=g)
    k = torch.randn(Bq, Lk, D, device=dev, generator=g)
    v = torch.randn(Bq, Lk, D, device=dev, generator=g)
    qq, kk, vv = (z.double().view(Bq, -1, H, dh).transpose(1, 2) for z in (q, k, v))
    ref = (torch.softmax(qq @ kk.transpose(-1, -2) / dh ** 0.5, -1) @ vv).transpose(1, 2).reshape(Bq * Lq, D)
    run = tc._Run(tc._Weights(), 1, dev)
    saved = _lib.lib.comet_get_option(_lib.OPT_ATTN_MMA)
    errs = []
    try:
        for on in (0, 3):
            _lib.check(_lib.lib.comet_set_option(_lib.OPT_ATTN_MMA, on))
            n0 = _lib.lib.comet_launch_count()
            op = run.attention(q.view(-1, D), k.view(-1, D), v.view(-1, D), Bq, H, Lq, Lk, dh, Lq * D, D, Lk * D, D, Bq * Lq, D,
                               Lq * D, D)
            torch.cuda.synchronize()
            assert _lib.lib.comet_launch_count() == n0 + 1
            errs.append(rel(op.float().sum(0)[:, :D], ref))
    finally:
        _lib.check(_lib.lib.comet_set_option(_lib.OPT_ATTN_MMA, saved))
    assert errs[0] < 1e-2 and errs[1] < 2e-2, errs          # one bf16 rounding / bf16 operands: the bar of BASELINE.md 5
