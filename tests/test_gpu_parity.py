"""Parity of the CUDA path (through the Python mirror -> ctypes -> C ABI) against the reference goldens and the
CPU oracle.  Bars: fp32 max|a-b|/max|b| <= 1e-4; bf16-autocast <= 2e-2."""
import numpy as np
import contextlib

import pytest
import torch

import cases
from conftest import rel_to_max
from oracle import comet_oracle as O

pytestmark = pytest.mark.gpu

FP32_BAR = 1e-4
BF16_BAR = 2e-2


@pytest.fixture(scope="module")
def cb():
    import comet_pose_estimation_b200 as m

    return m


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().float().cpu().numpy()


# ------------------------------------------------------------------ CorrBlock
@pytest.mark.parametrize("name", list(cases.CORR_CASES))
def test_corrblock_vs_reference_golden(cb, golden, name):
    g = golden("corr_blocks")
    kw, L, r = cases.CORR_CASES[name]
    fmaps, targets, coords = cases.corr_case(**kw)
    blk = cb.CorrBlock(dev(fmaps), num_levels=L, radius=r)
    blk.corr(dev(targets))
    out = blk.sample(dev(coords))
    assert out.shape == g[name + "/zeros"].shape and out.is_contiguous() and out.dtype == torch.float32
    assert rel_to_max(host(out), g[name + "/zeros"]) < FP32_BAR
    blk = cb.CorrBlock(dev(fmaps), num_levels=L, radius=r, padding_mode="border")
    blk.corr(dev(targets))
    assert rel_to_max(host(blk.sample(dev(coords))), g[name + "/border"]) < FP32_BAR
    eff = cb.EfficientCorrBlock(dev(fmaps), num_levels=L, radius=r)
    assert rel_to_max(host(eff.sample(dev(coords), dev(targets))), g[name + "/efficient"]) < FP32_BAR


@pytest.mark.parametrize("name", ["small_ragged", "tiny_odd"])
def test_pyramid_and_lazy_volumes(cb, golden, name):
    g = golden("corr_blocks")
    kw, L, r = cases.CORR_CASES[name]
    fmaps, targets, _ = cases.corr_case(**kw)
    blk = cb.CorrBlock(dev(fmaps), num_levels=L, radius=r)
    with pytest.raises(AttributeError):
        blk.corrs_pyramid
    blk.corr(dev(targets))
    assert len(blk.fmaps_pyramid) == L and len(blk.corrs_pyramid) == L
    for l in range(L):
        assert tuple(blk.fmaps_pyramid[l].shape) == g[f"{name}/pyr{l}"].shape
        assert rel_to_max(host(blk.fmaps_pyramid[l]), g[f"{name}/pyr{l}"]) < 1e-6
        assert rel_to_max(host(blk.corrs_pyramid[l]), g[f"{name}/vol{l}"]) < FP32_BAR


@pytest.mark.parametrize("name", list(cases.CORR_CASES))
def test_bf16_autocast_mode(cb, golden, name):
    g = golden("corr_blocks")
    kw, L, r = cases.CORR_CASES[name]
    fmaps, targets, coords = cases.corr_case(**kw)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        blk = cb.CorrBlock(dev(fmaps), num_levels=L, radius=r)
        blk.corr(dev(targets))
        out = blk.sample(dev(coords))
    assert out.dtype == torch.float32
    assert rel_to_max(host(out), g[name + "/zeros_bf16"]) < BF16_BAR
    # and it really is the bf16 variant, not the fp32 one
    assert rel_to_max(host(out), O.corr_lookup_bf16_autocast(fmaps, targets, coords, L, r)) < 4e-3


def test_strided_views_and_multiple_track_feats(cb):
    kw, L, r = cases.CORR_CASES["small_ragged"]
    fmaps, targets, coords = cases.corr_case(**kw)
    want = O.corr_lookup(fmaps, targets, coords, L, r)
    # targets / coords handed over as permuted (B,N,S,.) storage, like track_feats in the tracker loop
    t_perm = dev(targets.transpose(0, 2, 1, 3)).permute(0, 2, 1, 3)
    c_perm = dev(coords.transpose(0, 2, 1, 3)).permute(0, 2, 1, 3)
    assert not t_perm.is_contiguous()
    blk = cb.CorrBlock(dev(fmaps), num_levels=L, radius=r)
    blk.corr(t_perm)
    assert rel_to_max(host(blk.sample(c_perm)), want) < FP32_BAR
    # multiple_track_feats: level l correlates channels [l*C, (l+1)*C) of the target (blocks.py:411-425)
    rng = np.random.default_rng(5)
    B, S, N, C = targets.shape
    big = rng.standard_normal((B, S, N, C * L)).astype(np.float32)
    pyr = O.build_pyramid(fmaps, L)
    vols = [O.corr_volumes(big[..., l * C:(l + 1) * C], [pyr[l]])[0] for l in range(L)]
    want = O.lookup_volumes(vols, coords, r, "zeros")
    blk = cb.CorrBlock(dev(fmaps), num_levels=L, radius=r, multiple_track_feats=True)
    blk.corr(dev(big))
    assert rel_to_max(host(blk.sample(dev(coords))), want) < FP32_BAR
    for l in range(L):
        assert rel_to_max(host(blk.corrs_pyramid[l]), vols[l]) < FP32_BAR


def test_edge_cases(cb):
    # empty track set
    f = torch.randn(1, 2, 8, 8, 8, device="cuda")
    blk = cb.CorrBlock(f, num_levels=2, radius=2)
    blk.corr(torch.zeros(1, 2, 0, 8, device="cuda"))
    out = blk.sample(torch.zeros(1, 2, 0, 2, device="cuda"))
    assert out.shape == (1, 2, 0, 50)
    # reference assertions (blocks.py:379, :415-416)
    with pytest.raises(AssertionError):
        blk.corr(torch.zeros(1, 2, 3, 7, device="cuda"))
    with pytest.raises(AssertionError):
        blk.corr(torch.zeros(1, 3, 3, 8, device="cuda"))
    blk.corr(torch.zeros(1, 2, 3, 8, device="cuda"))
    with pytest.raises(AssertionError):
        blk.sample(torch.zeros(1, 2, 3, 3, device="cuda"))
    with pytest.raises(AssertionError):  # pyramid deeper than the map
        cb.CorrBlock(torch.randn(1, 1, 4, 4, 4, device="cuda"), num_levels=4)
    # wild coordinates: far outside / huge -> zeros under zero padding, finite under border
    kw, L, r = cases.CORR_CASES["deep_pyramid"]
    fmaps, targets, coords = cases.corr_case(**kw)
    coords[..., 0] = 1.0e9
    coords[0, 0, 0] = (-1.0e9, 3.0)
    blk = cb.CorrBlock(dev(fmaps), num_levels=L, radius=r)
    blk.corr(dev(targets))
    assert float(blk.sample(dev(coords)).abs().max()) == 0.0
    eff = cb.EfficientCorrBlock(dev(fmaps), num_levels=L, radius=r)
    assert bool(torch.isfinite(eff.sample(dev(coords), dev(targets))).all())
    # maximum radius / level count the kernels accept
    f = torch.randn(1, 1, 4, 128, 128, device="cuda")
    t = torch.randn(1, 1, 3, 4, device="cuda")
    c = torch.rand(1, 1, 3, 2, device="cuda") * 127
    blk = cb.CorrBlock(f, num_levels=8, radius=7)
    blk.corr(t)
    want = O.corr_lookup(host(f), host(t), host(c), 8, 7)
    assert rel_to_max(host(blk.sample(c)), want) < FP32_BAR


def test_edge_cases_tensor_and_channels_last_paths(cb):
    """Empty inputs through the specialised paths: no queries on the tcgen05 path, no maps / no queries on the
    channels-last TMA path, empty batches through the encoder kernels."""
    f = torch.randn(1, 2, 128, 64, 64, device="cuda")
    blk = cb.CorrBlock(f, num_levels=5, radius=4)
    blk.corr(torch.zeros(1, 2, 0, 128, device="cuda"))
    c0 = torch.zeros(1, 2, 0, 2, device="cuda")
    assert blk.sample(c0).shape == (1, 2, 0, 405)
    x = cb.TrackTokenizer(blk, c0[:, 0], 664).tokens(c0, torch.zeros(1, 2, 0, 128, device="cuda"))
    assert x.shape == (1, 0, 2, 664)
    fcl = _channels_last_view(torch.randn(3, 2, 32, 31, 31, device="cuda"))
    blk = cb.CorrBlock(fcl, num_levels=3, radius=3)
    blk.corr(torch.zeros(3, 2, 0, 32, device="cuda"))
    assert blk.sample(torch.zeros(3, 2, 0, 2, device="cuda")).shape == (3, 2, 0, 147)
    empty = torch.zeros(0, 2, 32, 31, 31, device="cuda")
    blk = cb.CorrBlock(empty, num_levels=3, radius=3)
    blk.corr(torch.zeros(0, 2, 1, 32, device="cuda"))
    assert blk.sample(torch.zeros(0, 2, 1, 2, device="cuda")).shape == (0, 2, 1, 147)
    assert cb.upsample_bilinear_align_corners(torch.zeros(0, 32, 16, 16, device="cuda"), (31, 31)).shape == (0, 32, 31, 31)
    assert cb.instance_norm(torch.zeros(0, 32, 8, 8, device="cuda"), relu=True).shape == (0, 32, 8, 8)
    assert cb.sample_features4d(fcl[:0, 0], torch.zeros(0, 4, 2, device="cuda")).shape == (0, 4, 32)
    # a single query and a single frame on the TMA path; coordinates far outside / non-finite -> zeros (zero padding)
    one = _channels_last_view(torch.randn(1, 1, 32, 31, 31, device="cuda"))
    blk = cb.CorrBlock(one, num_levels=3, radius=3)
    blk.corr(torch.randn(1, 1, 1, 32, device="cuda"))
    for bad in (1.0e9, -1.0e9, float("inf")):
        assert float(blk.sample(torch.full((1, 1, 1, 2), bad, device="cuda")).abs().max()) == 0.0


# ------------------------------------------------------------------ samplers / encodings
def test_samplers_vs_reference_golden(cb, golden):
    g = golden("samplers_encodings")
    inp, xy = cases.sampler_case(21, B=2, C=3, H=9, W=7, Ho=4, Wo=5)
    for pm in ("zeros", "border"):
        for ac in (True, False):
            got = cb.bilinear_sampler(dev(inp), dev(xy), align_corners=ac, padding_mode=pm)
            assert rel_to_max(host(got), g[f"bs4/{pm}/{int(ac)}"]) < FP32_BAR
    inp5, txy = cases.sampler_case(22, B=2, C=3, H=6, W=8, Ho=3, Wo=4, T=3)
    for pm in ("zeros", "border"):
        assert rel_to_max(host(cb.bilinear_sampler(dev(inp5), dev(txy), padding_mode=pm)), g[f"bs5/{pm}"]) < FP32_BAR
    inp1, txy1 = cases.sampler_case(23, B=2, C=4, H=6, W=8, Ho=3, Wo=4, T=1)
    txy1[..., 0] = 0
    assert rel_to_max(host(cb.bilinear_sampler(dev(inp1), dev(txy1))), g["bs5_t1/border"]) < FP32_BAR
    assert rel_to_max(host(cb.sample_features4d(dev(inp), dev(g["sf4d/pts"]))), g["sf4d/out"]) < FP32_BAR
    # batch-strided view (fmaps[:, 0]) and batch-expanded table
    big = torch.randn(2, 3, 3, 9, 7, device="cuda")
    pts = dev(g["sf4d/pts"])
    assert rel_to_max(host(cb.sample_features4d(big[:, 1], pts)), O.sample_features4d(host(big[:, 1]), host(pts))) < FP32_BAR
    tab = torch.randn(1, 3, 9, 7, device="cuda").expand(2, -1, -1, -1)
    assert rel_to_max(host(cb.sample_features4d(tab, pts)), O.sample_features4d(host(tab), host(pts))) < FP32_BAR
    with pytest.raises(AssertionError):
        cb.bilinear_sampler(torch.zeros(1, 2, 3, device="cuda"), torch.zeros(1, 1, 1, 2, device="cuda"))


def test_encodings_vs_reference_golden(cb, golden):
    g = golden("samplers_encodings")
    for C, scale in ((64, 3.0), (16, 0.7), (64, 40.0)):
        xy = g[f"emb2d/{C}/{scale}/xy"]
        # same float32 argument, accurate sinf/cosf: absolute error of a few ulp of 1.0
        assert np.abs(host(cb.get_2d_embedding(dev(xy), C, cat_coords=False)) - g[f"emb2d/{C}/{scale}/nocat"]).max() < 2e-6
        assert np.abs(host(cb.get_2d_embedding(dev(xy), C, cat_coords=True)) - g[f"emb2d/{C}/{scale}/cat"]).max() < 2e-6
    assert np.abs(host(cb.get_2d_sincos_pos_embed(216, (31, 31))) - g["sincos2d/216_31"]).max() < 2e-7
    full = host(cb.get_2d_sincos_pos_embed(664, (64, 64)))
    assert full.shape == (1, 664, 64, 64)
    assert np.abs(full[:, :, ::9, :] - g["sincos2d/664_64/rows"]).max() < 2e-7
    assert abs(full.astype(np.float64).sum() - g["sincos2d/664_64/sum"][0]) < 1e-2
    assert np.abs(host(cb.get_2d_sincos_pos_embed(12, (3, 5))) - g["sincos2d/12_h3w5"]).max() < 2e-7
    pe, grid = cb.get_2d_sincos_pos_embed(8, 4, return_grid=True)
    assert np.abs(host(pe) - g["sincos2d/8_4"]).max() < 2e-7 and np.array_equal(host(grid), g["sincos2d/8_4/grid"])
    assert np.abs(host(cb.get_1d_sincos_pos_embed(768, 16)) - g["sincos1d/768_16"]).max() < 2e-7
    assert np.abs(host(cb.get_1d_sincos_pos_embed(768, 64)) - g["sincos1d/768_64"]).max() < 2e-7
    pe1, g1 = cb.get_1d_sincos_pos_embed(10, 7, return_grid=True)
    assert np.abs(host(pe1) - g["sincos1d/10_7"]).max() < 2e-7 and np.array_equal(host(g1), g["sincos1d/10_7/grid"])
    got = cb.get_1d_sincos_pos_embed_from_grid(14, dev(g["sincos1dgrid/pos"]))
    assert np.abs(host(got) - g["sincos1dgrid/14"]).max() < 2e-7
    grid = torch.stack([torch.arange(6.0), torch.arange(6.0) * 0.5]).cuda()
    got = cb.get_2d_sincos_pos_embed_from_grid(12, grid)
    want = O.get_2d_sincos_pos_embed_from_grid(12, host(grid))
    assert got.shape == want.shape and np.abs(host(got) - want).max() < 2e-7


# ------------------------------------------------------------------ fused track tokens
TOKEN_CASES = {
    "coarse_full_it0": (dict(seed=44, B=1, S=3, C=128, H=64, W=64, N=16, stride=4, down_ratio=2), 5, 4, 128, False, 8.0),
    "fine_full_it0": (dict(seed=45, B=8, S=3, C=32, H=31, W=31, N=1, stride=1, down_ratio=1), 3, 3, 32, True, 1.0),
}


@pytest.mark.parametrize("name", list(TOKEN_CASES))
def test_first_iteration_tokens_vs_reference_golden(cb, golden, name):
    g = golden("tracker")
    kw, L, r, latent, fine, qscale = TOKEN_CASES[name]
    fmaps, q = cases.tracker_case(**kw)
    B, S = fmaps.shape[:2]
    N = q.shape[1]
    q = dev(q) / qscale
    coords = q.reshape(B, 1, N, 2).repeat(1, S, 1, 1)
    f = dev(fmaps)
    qfeat = cb.sample_features4d(f[:, 0], coords[:, 0])
    feats = qfeat.unsqueeze(1).repeat(1, S, 1, 1)
    tdim = cb.transformer_dim(L, r, latent, fine)
    want = g[name + "/tok0"]
    assert tdim == want.shape[-1]
    tok = cb.TrackTokenizer(cb.CorrBlock(f, num_levels=L, radius=r), coords[:, 0], tdim)
    x = tok.tokens(coords, feats)
    assert tuple(x.shape) == want.shape
    assert rel_to_max(host(x), want) < FP32_BAR


def test_tokens_with_motion_vs_oracle(cb):
    """Later iterations: coords differ per frame (non-zero flows with sin/cos arguments up to ~1e4), track_feats
    arrive as a permuted (B,N,S,C) view, as in base_track_predictor.py:243-252."""
    rng = np.random.default_rng(3)
    # (L, r, C, fine): token widths 216 (fine rule) and 132 (coarse rule, no padding needed)
    for L, r, C, fine in ((3, 3, 32, True), (2, 3, 16, False), (5, 2, 16, False)):
        B, S, N, H, W = 2, 5, 11, 24, 20
        fmaps = rng.standard_normal((B, S, C, H, W)).astype(np.float32)
        feats = rng.standard_normal((B, S, N, C)).astype(np.float32)
        c0 = np.stack([rng.uniform(0, W - 1, (B, N)), rng.uniform(0, H - 1, (B, N))], -1).astype(np.float32)
        coords = c0[:, None] + (rng.standard_normal((B, S, N, 2)) * 3).astype(np.float32)
        coords[:, 0] = c0
        tdim = O.transformer_dim(L, r, C, fine)
        want = O.track_tokens(O.corr_lookup(fmaps, feats, coords, L, r), coords, feats, (H, W), tdim)
        f = dev(fmaps)
        feats_view = dev(feats.transpose(0, 2, 1, 3)).permute(0, 2, 1, 3)
        tok = cb.TrackTokenizer(f, dev(c0), tdim, num_levels=L, radius=r)
        x = tok.tokens(dev(coords), feats_view)
        assert rel_to_max(host(x), want) < FP32_BAR
        # EfficientCorrBlock flavour (border padding)
        want_b = O.track_tokens(O.efficient_corr_lookup(fmaps, feats, coords, L, r), coords, feats, (H, W), tdim)
        tok_b = cb.TrackTokenizer(cb.EfficientCorrBlock(f, num_levels=L, radius=r), dev(c0), tdim)
        assert rel_to_max(host(tok_b.tokens(dev(coords), dev(feats))), want_b) < FP32_BAR


# ------------------------------------------------------------------ full-size properties (BASELINE configs)
def _full_coarse(S, N, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    fmaps = torch.randn(1, S, 128, 64, 64, device="cuda", generator=g)
    feats = torch.randn(1, S, N, 128, device="cuda", generator=g)
    coords = torch.rand(1, S, N, 2, device="cuda", generator=g) * 63
    wild = torch.rand(1, S, N, 1, device="cuda", generator=g) < 0.05
    coords = torch.where(wild, torch.rand(1, S, N, 2, device="cuda", generator=g) * 76 - 6, coords)
    return fmaps, feats, coords


def test_full_size_coarse_config_subset_vs_oracle(cb):
    """Config 1/2 shape (S=16, N=512, C=128, 64x64, L=5, r=4): queries are independent, so the oracle checks a
    slice of them in seconds while the kernel runs the full problem."""
    fmaps, feats, coords = _full_coarse(16, 512)
    blk = cb.CorrBlock(fmaps, num_levels=5, radius=4)
    blk.corr(feats)
    out = blk.sample(coords)
    assert out.shape == (1, 16, 512, 405)
    sel = slice(0, 512, 37)
    want = O.corr_lookup(host(fmaps), host(feats[:, :, sel]), host(coords[:, :, sel]), 5, 4)
    assert rel_to_max(host(out[:, :, sel]), want) < FP32_BAR


def test_full_size_properties(cb):
    """Size-independent properties at full size: linearity in the target, pooling commutes with correlation
    (level l of the output == level 0 of a CorrBlock built on the pooled map), and consistency between the
    lookup-only and the token kernels."""
    fmaps, feats, coords = _full_coarse(16, 512, seed=1)
    L, r = 5, 4
    blk = cb.CorrBlock(fmaps, num_levels=L, radius=r)
    blk.corr(feats)
    a = blk.sample(coords)
    blk.corr(feats * -2.5)
    b = blk.sample(coords)
    assert rel_to_max(host(b), host(a) * -2.5) < 1e-5
    # level 1 of the pyramid == level 0 of a block built on the pooled maps, at halved coordinates
    pooled = blk.fmaps_pyramid[1].contiguous()
    blk1 = cb.CorrBlock(pooled, num_levels=1, radius=r)
    blk1.corr(feats)
    lvl1 = blk1.sample(coords / 2)
    assert rel_to_max(host(lvl1), host(a[..., 81:162])) < 1e-5
    # tokens carry the same correlation features, transposed to (B,N,S,.) and offset by the position embedding
    tdim = cb.transformer_dim(L, r, 128, False)
    tok = cb.TrackTokenizer(blk, coords[:, 0], tdim)
    x = tok.tokens(coords, feats)
    corr_part = x[..., 130:130 + 405] - tok.pos[:, :, None, 130:130 + 405]
    assert rel_to_max(host(corr_part.permute(0, 2, 1, 3)), host(a)) < 1e-5
    assert float((x[..., 663] - tok.pos[:, :, None, 663]).abs().max()) == 0.0  # zero pad channel


def test_config4_long_sequence_dense_grid(cb):
    """BASELINE.json configs[3]: S=64, N=4096 = dense query grid at the pixel centres (8i+4, 8j+4) of a 512^2 image
    (= cell coordinates (i+.5, j+.5) of the 64x64 map), displaced per frame.  The reference would materialise a
    5.7 GB volume per iteration; here nothing larger than the 696 MB token tensor exists.  The oracle checks a slice
    of (frame, query) pairs; linearity and tokens-vs-lookup consistency cover the rest."""
    S, N, L, r = 64, 4096, 5, 4
    g = torch.Generator(device="cuda").manual_seed(4)
    fmaps = torch.randn(1, S, 128, 64, 64, device="cuda", generator=g)
    feats = torch.randn(1, S, N, 128, device="cuda", generator=g)
    ii, jj = torch.meshgrid(torch.arange(64, device="cuda"), torch.arange(64, device="cuda"), indexing="ij")
    q = torch.stack([jj.flatten() + 0.5, ii.flatten() + 0.5], -1).float()            # (4096, 2) = (x, y)
    coords = q[None, None] + torch.randn(1, S, N, 2, device="cuda", generator=g) * 1.5
    coords[:, 0] = q
    blk = cb.CorrBlock(fmaps, num_levels=L, radius=r)
    blk.corr(feats)
    out = blk.sample(coords)
    assert out.shape == (1, S, N, 405)
    if cb._lib.lib.comet_has_tensor_path():
        assert cb._lib.lib.comet_tc_status() == 0
    fs, qs = [0, 17, 46], slice(5, N, 397)   # frame 0 first: the oracle's flows / position embedding refer to it
    fm_s, ft_s, co_s = host(fmaps[:, fs]), host(feats[:, fs][:, :, qs]), host(coords[:, fs][:, :, qs])
    want = O.corr_lookup(fm_s, ft_s, co_s, L, r)
    assert rel_to_max(host(out[:, fs][:, :, qs]), want) < FP32_BAR
    tdim = cb.transformer_dim(L, r, 128, False)
    tok = cb.TrackTokenizer(blk, coords[:, 0], tdim)
    x = tok.tokens(coords, feats)
    assert x.shape == (1, N, S, 664)
    corr_part = x[..., 130:535] - tok.pos[:, :, None, 130:535]
    assert float((corr_part.permute(0, 2, 1, 3) - out).abs().max()) < 1e-4 * float(out.abs().max())
    want_x = O.track_tokens(want, co_s, ft_s, (64, 64), tdim)
    assert rel_to_max(host(x[:, qs][:, :, fs]), want_x) < FP32_BAR
    # linearity in the target at full size
    blk.corr(feats * 0.5)
    assert rel_to_max(host(blk.sample(coords)), host(out) * 0.5) < 1e-5


def test_full_size_fine_config_subset_vs_oracle(cb):
    """Fine tracker shape: B' = 512 patches, S=16, one query per 31x31 patch, C=32, L=3, r=3."""
    g = torch.Generator(device="cuda").manual_seed(2)
    fmaps = torch.randn(512, 16, 32, 31, 31, device="cuda", generator=g)
    feats = torch.randn(512, 16, 1, 32, device="cuda", generator=g)
    coords = torch.rand(512, 16, 1, 2, device="cuda", generator=g) * 30
    blk = cb.CorrBlock(fmaps, num_levels=3, radius=3)
    blk.corr(feats)
    out = blk.sample(coords)
    assert out.shape == (512, 16, 1, 147)
    sel = slice(0, 512, 61)
    want = O.corr_lookup(host(fmaps[sel]), host(feats[sel]), host(coords[sel]), 3, 3)
    assert rel_to_max(host(out[sel]), want) < FP32_BAR


def _channels_last_view(fmaps):
    """Same values, (B,S,C,H,W) shape, memory laid out (B,S,H,W,C): what a torch.channels_last encoder returns."""
    return fmaps.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)


@pytest.mark.parametrize("shape", [(24, 4, 32, 31, 31, 3, 3), (3, 2, 32, 31, 31, 2, 2), (2, 3, 16, 20, 13, 3, 3),
                                   (2, 2, 32, 31, 31, 1, 3), (5, 2, 8, 9, 9, 4, 1)])
def test_channels_last_fmaps_are_used_zero_copy(cb, shape):
    """A channels-last fmaps view (the producer's native layout) goes through the same kernels without a copy and
    gives the reference's numbers: pyramid, lookup (both paddings), tokens, lazily materialised volumes."""
    B, S, C, H, W, L, r = shape
    g = torch.Generator(device="cuda").manual_seed(11)
    fmaps = torch.randn(B, S, C, H, W, device="cuda", generator=g)
    feats = torch.randn(B, S, 2, C, device="cuda", generator=g)
    coords = torch.rand(B, S, 2, 2, device="cuda", generator=g) * torch.tensor([W + 6.0, H + 6.0], device="cuda") - 3
    fcl = _channels_last_view(fmaps)
    assert not fcl.is_contiguous() and torch.equal(fcl, fmaps)
    blk = cb.CorrBlock(fcl, num_levels=L, radius=r)
    assert blk._pyr.cl_input and blk._pyr.fmaps0.data_ptr() == fcl.data_ptr()
    want_levels = O.build_pyramid(host(fmaps), L)
    for l in range(L):
        assert rel_to_max(host(blk.fmaps_pyramid[l]), want_levels[l]) < 1e-6
    blk.corr(feats)
    want = O.corr_lookup(host(fmaps), host(feats), host(coords), L, r, "zeros")
    assert rel_to_max(host(blk.sample(coords)), want) < FP32_BAR
    eff = cb.EfficientCorrBlock(fcl, num_levels=L, radius=r)
    want_b = O.corr_lookup(host(fmaps), host(feats), host(coords), L, r, "border")
    assert rel_to_max(host(eff.sample(coords, feats)), want_b) < FP32_BAR
    ref = cb.CorrBlock(fmaps, num_levels=L, radius=r)
    ref.corr(feats)
    for a, b in zip(blk.corrs_pyramid, ref.corrs_pyramid):
        assert rel_to_max(host(a), host(b)) < 1e-5
    tdim = (2 * C + 2 + L * (2 * r + 1) ** 2 + 3) // 4 * 4 + 4   # any multiple of 4 that holds the channels
    x = cb.TrackTokenizer(blk, coords[:, 0], tdim).tokens(coords, feats)
    x_ref = cb.TrackTokenizer(ref, coords[:, 0], tdim).tokens(coords, feats)
    assert rel_to_max(host(x), host(x_ref)) < 1e-5
    # the autocast (bf16 rounding) variant of the same kernels, both layouts, against the oracle's bf16 restatement
    with torch.autocast("cuda", dtype=torch.bfloat16):
        a = cb.CorrBlock(fcl, num_levels=L, radius=r)
        a.corr(feats)
        out_cl = a.sample(coords)
        b = cb.CorrBlock(fmaps, num_levels=L, radius=r)
        b.corr(feats)
        out_nchw = b.sample(coords)
        x_bf = cb.TrackTokenizer(a, coords[:, 0], tdim).tokens(coords, feats)
    want_bf = O.corr_lookup_bf16_autocast(host(fmaps), host(feats), host(coords), L, r)
    assert rel_to_max(host(out_cl), want_bf) < 4e-3
    assert rel_to_max(host(out_cl), host(out_nchw)) < 4e-3
    WW = (2 * r + 1) ** 2
    pos = cb.TrackTokenizer(a, coords[:, 0], tdim).pos
    assert rel_to_max(host((x_bf[..., C + 2:C + 2 + L * WW] - pos[:, :, None, C + 2:C + 2 + L * WW]).permute(0, 2, 1, 3)),
                      host(out_cl)) < 1e-5


def test_channels_last_point_sampler_and_cached_pos_emb(cb, golden):
    """sample_features4d on a channels-last view == on the contiguous tensor == the oracle; the sampled position
    embedding from the cached channel-last sin/cos table == the on-the-fly evaluation == the oracle."""
    from comet_pose_estimation_b200.track_tokens import sampled_pos_emb

    g = torch.Generator(device="cuda").manual_seed(13)
    x = torch.randn(6, 32, 31, 31, device="cuda", generator=g)
    pts = torch.rand(6, 9, 2, device="cuda", generator=g) * 36 - 3
    xcl = x.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    a, b = cb.sample_features4d(x, pts), cb.sample_features4d(xcl, pts)
    assert rel_to_max(host(b), O.sample_features4d(host(x), host(pts))) < 1e-5
    assert rel_to_max(host(b), host(a)) < 1e-5
    view = _channels_last_view(torch.randn(4, 3, 32, 31, 31, device="cuda", generator=g))[:, 0]   # fmaps[:, 0]
    assert rel_to_max(host(cb.sample_features4d(view, pts[:4])), O.sample_features4d(host(view), host(pts[:4]))) < 1e-5
    for (D, H, W) in ((664, 64, 64), (216, 31, 31), (12, 3, 5)):
        c0 = torch.rand(2, 17, 2, device="cuda", generator=g) * torch.tensor([W + 4.0, H + 4.0], device="cuda") - 2
        t1 = sampled_pos_emb(c0, D, H, W, cached_table=True)
        t2 = sampled_pos_emb(c0, D, H, W, cached_table=False)
        tab = O.get_2d_sincos_pos_embed(D, (H, W))
        want = O.sample_features4d(np.broadcast_to(tab, (2,) + tab.shape[1:]), host(c0))
        assert rel_to_max(host(t1), want) < 1e-5
        assert rel_to_max(host(t2), want) < 1e-5


@pytest.mark.parametrize("shape", [(37, 32, 16, 16, 31, 31), (5, 32, 8, 8, 16, 16), (3, 32, 4, 4, 16, 16),
                                   (2, 6, 7, 5, 12, 9), (4, 8, 3, 3, 1, 1), (2, 4, 1, 1, 5, 4)])
def test_bilinear_resize_matches_interpolate(cb, shape):
    """upsample_bilinear_align_corners == F.interpolate(..., mode="bilinear", align_corners=True), contiguous and
    channels-last (the three resizes of the fine tracker's patch encoder, blocks.py:176-190, and odd shapes)."""
    N, C, Hi, Wi, Ho, Wo = shape
    g = torch.Generator(device="cuda").manual_seed(17)
    x = torch.randn(N, C, Hi, Wi, device="cuda", generator=g)
    want = torch.nn.functional.interpolate(x, (Ho, Wo), mode="bilinear", align_corners=True)
    got = cb.upsample_bilinear_align_corners(x, (Ho, Wo))
    assert got.is_contiguous() and rel_to_max(host(got), host(want)) < 1e-6
    xcl = x.contiguous(memory_format=torch.channels_last)
    got_cl = cb.upsample_bilinear_align_corners(xcl, (Ho, Wo))
    assert got_cl.shape == want.shape and rel_to_max(host(got_cl), host(want)) < 1e-6
    if C % 4 == 0 and C > 1 and Ho * Wo > 1 and not xcl.is_contiguous():
        assert got_cl.is_contiguous(memory_format=torch.channels_last)   # the memory format is preserved


@pytest.mark.parametrize("shape", [(9, 32, 16, 16), (5, 32, 8, 8), (3, 32, 4, 4), (2, 5, 7, 3), (4, 1, 6, 6)])
def test_instance_norm_matches_torch(cb, shape):
    g = torch.Generator(device="cuda").manual_seed(19)
    x = torch.randn(*shape, device="cuda", generator=g) * 3 + 1.5
    norm = torch.nn.InstanceNorm2d(shape[1])
    for relu in (False, True):
        want = norm(x)
        want = torch.relu(want) if relu else want
        assert rel_to_max(host(cb.instance_norm(x, relu=relu)), host(want)) < 2e-6
        xcl = x.contiguous(memory_format=torch.channels_last)
        got = cb.instance_norm(xcl, relu=relu)
        assert got.stride() == xcl.stride() and rel_to_max(host(got), host(want)) < 2e-6


def test_full_size_fine_config_channels_last(cb):
    """Fine tracker at full size with channels-last patch features == the NCHW result (same oracle slice)."""
    g = torch.Generator(device="cuda").manual_seed(2)
    fmaps = torch.randn(512, 16, 32, 31, 31, device="cuda", generator=g)
    feats = torch.randn(512, 16, 1, 32, device="cuda", generator=g)
    coords = torch.rand(512, 16, 1, 2, device="cuda", generator=g) * 30
    blk = cb.CorrBlock(_channels_last_view(fmaps), num_levels=3, radius=3)
    assert blk._pyr.cl_input
    blk.corr(feats)
    out = blk.sample(coords)
    sel = slice(0, 512, 61)
    want = O.corr_lookup(host(fmaps[sel]), host(feats[sel]), host(coords[sel]), 3, 3)
    assert rel_to_max(host(out[sel]), want) < FP32_BAR


# ------------------------------------------------------------------ tensor-core path (tcgen05) vs SIMT path
def test_tensor_path_degenerate_query_sets(cb):
    """The tensor path sorts the queries of a frame by row and only streams the band of map rows a tile of 128
    queries needs.  Degenerate bands: every query on one row, every query off the map (wild / non-finite
    coordinates), a single query, bands narrower than the row split, many query tiles per frame."""
    if not cb._lib.lib.comet_has_tensor_path():
        pytest.skip("no sm_100 tensor path on this device")
    g = torch.Generator(device="cuda").manual_seed(21)
    S, L, r = 2, 5, 4
    fmaps = torch.randn(1, S, 128, 64, 64, device="cuda", generator=g)
    sets = {}
    c = torch.rand(1, S, 300, 2, device="cuda", generator=g) * 63
    c[..., 1] = 17.25
    sets["one_row"] = c
    c = torch.rand(1, S, 140, 2, device="cuda", generator=g) * 63
    c[:, :, ::2] = 1.0e7
    c[:, :, 1::4] = -3.0e5
    c[0, 0, 5, 0] = float("inf")
    c[0, 1, 7, 1] = float("-inf")
    sets["mostly_off_map"] = c
    sets["all_off_map"] = torch.full((1, S, 130, 2), -500.0, device="cuda")
    sets["single"] = torch.tensor([[[[63.0, 63.0]], [[0.0, 0.0]]]], device="cuda")
    c = torch.rand(1, S, 1500, 2, device="cuda", generator=g) * 70 - 3
    sets["many_tiles"] = c
    c = torch.rand(1, S, 257, 2, device="cuda", generator=g) * 63
    c[..., 1] = c[..., 1] * 0.02 + 62.5   # band hugging the bottom border
    sets["bottom_band"] = c
    # more than 256 query tiles per frame: the plan kernel stops tracking bands and every job covers the full maps
    sets["more_tiles_than_tracked"] = torch.rand(1, S, 256 * 128 + 77, 2, device="cuda", generator=g) * 66 - 1.5
    for name, coords in sets.items():
        N = coords.shape[2]
        feats = torch.randn(1, S, N, 128, device="cuda", generator=g)
        blk = cb.CorrBlock(fmaps, num_levels=L, radius=r)
        assert blk._pyr.split is not None
        blk.corr(feats)
        got = blk.sample(coords)
        tdim = cb.transformer_dim(L, r, 128, False)
        x = cb.TrackTokenizer(blk, coords[:, 0], tdim).tokens(coords, feats)
        assert cb._lib.lib.comet_tc_status() == 0, name
        finite = torch.isfinite(coords).all(-1)
        sel = torch.arange(0, N, max(1, N // 40), device="cuda")
        want = O.corr_lookup(host(fmaps), host(feats[:, :, sel]), host(torch.nan_to_num(coords[:, :, sel], posinf=1e7, neginf=-1e7)),
                             L, r)
        err = np.abs(host(got[:, :, sel]) - want)[host(finite[:, :, sel]).astype(bool)]
        assert err.size == 0 or err.max() / max(np.abs(want).max(), 1.0) < FP32_BAR, name
        corr_part = x[..., 130:130 + 405] - cb.TrackTokenizer(blk, coords[:, 0], tdim).pos[:, :, None, 130:535]
        ok = finite.permute(0, 2, 1)
        assert rel_to_max(host(corr_part[ok]), host(got.permute(0, 2, 1, 3)[ok])) < 1e-5, name


def test_tensor_path_matches_simt_and_oracle(cb):
    """Coarse COMET shape (C=128, 64x64, L=5, r=4, zero padding) runs on the tcgen05 kernel; the same call with
    the tensor path switched off (comet_set_option) runs the SIMT kernel.  Both must sit inside the fp32 bar against the oracle, for
    N that is not a multiple of the 128-query tile, for strided targets, and for smaller L / r."""
    if not cb._lib.lib.comet_has_tensor_path():
        pytest.skip("no sm_100 tensor path on this device")
    g = torch.Generator(device="cuda").manual_seed(7)
    for (S, N, L, r) in ((3, 200, 5, 4), (2, 129, 3, 2), (1, 5, 1, 0), (2, 384, 4, 3)):
        fmaps = torch.randn(2, S, 128, 64, 64, device="cuda", generator=g)
        feats = torch.randn(2, N, S, 128, device="cuda", generator=g).permute(0, 2, 1, 3)  # strided view
        coords = torch.rand(2, S, N, 2, device="cuda", generator=g) * 80 - 8
        coords[0, 0, 0] = torch.tensor([63.0, 0.0], device="cuda")
        coords[1, 0, 0] = torch.tensor([-30.0, 31.25], device="cuda")
        cb._lib.set_option(cb._lib.OPT_TENSOR_PATH, True)
        blk = cb.CorrBlock(fmaps, num_levels=L, radius=r)
        assert blk._pyr.split is not None
        blk.corr(feats)
        tc = blk.sample(coords)
        tdim = cb.transformer_dim(L, r, 128, False)
        if tdim >= 2 * 128 + 2 + L * (2 * r + 1) ** 2:
            x_tc = cb.TrackTokenizer(blk, coords[:, 0], tdim).tokens(coords, feats)
        else:
            x_tc = None
        vols_tc = [v.clone() for v in blk.corrs_pyramid]
        cb._lib.set_option(cb._lib.OPT_TENSOR_PATH, False)
        blk2 = cb.CorrBlock(fmaps, num_levels=L, radius=r)
        assert blk2._pyr.split is None
        blk2.corr(feats)
        simt = blk2.sample(coords)
        assert rel_to_max(host(tc), host(simt)) < 2e-5
        for a, b in zip(vols_tc, blk2.corrs_pyramid):
            assert rel_to_max(host(a), host(b)) < 2e-5
        if x_tc is not None:
            x_simt = cb.TrackTokenizer(blk2, coords[:, 0], tdim).tokens(coords, feats)
            assert rel_to_max(host(x_tc), host(x_simt)) < 2e-5
        sel = slice(0, N, max(1, N // 9))
        want = O.corr_lookup(host(fmaps), host(feats[:, :, sel]), host(coords[:, :, sel]), L, r)
        assert rel_to_max(host(tc[:, :, sel]), want) < FP32_BAR
        cb._lib.set_option(cb._lib.OPT_TENSOR_PATH, True)
        assert cb._lib.lib.comet_tc_status() == 0
    cb._lib.set_option(cb._lib.OPT_TENSOR_PATH, True)


def test_tensor_path_bf16_autocast(cb):
    if not cb._lib.lib.comet_has_tensor_path():
        pytest.skip("no sm_100 tensor path on this device")
    g = torch.Generator(device="cuda").manual_seed(8)
    fmaps = torch.randn(1, 2, 128, 64, 64, device="cuda", generator=g)
    feats = torch.randn(1, 2, 150, 128, device="cuda", generator=g)
    coords = torch.rand(1, 2, 150, 2, device="cuda", generator=g) * 70 - 3
    with torch.autocast("cuda", dtype=torch.bfloat16):
        blk = cb.CorrBlock(fmaps, num_levels=5, radius=4)
        blk.corr(feats)
        out = blk.sample(coords)
    want = O.corr_lookup_bf16_autocast(host(fmaps), host(feats), host(coords), 5, 4)
    assert rel_to_max(host(out), want) < 8e-3          # same rounding points (reciprocal multiply vs divide)
    assert rel_to_max(host(out), O.corr_lookup(host(fmaps), host(feats), host(coords), 5, 4)) < BF16_BAR


def test_forward_only_contract_and_corr_aliasing(cb):
    """(i) an input that requires grad under enabled autograd is refused (the kernels build no autograd
    graph; the reference runs the tracker under no_grad); (ii) corr() keeps a reference to `targets`, so an in-place
    update before sample() -- which the reference, computing the volume inside corr(), would not see -- is detected."""
    fmaps = torch.randn(1, 2, 16, 12, 12, device="cuda")
    targets = torch.randn(1, 2, 5, 16, device="cuda")
    coords = torch.rand(1, 2, 5, 2, device="cuda") * 11
    blk = cb.CorrBlock(fmaps, num_levels=2, radius=2)
    blk.corr(targets)
    want = blk.sample(coords).clone()
    with torch.enable_grad():
        with pytest.raises(RuntimeError, match="forward-only"):
            cb.CorrBlock(fmaps.clone().requires_grad_(True), num_levels=2, radius=2)
        blk.corr(targets.clone().requires_grad_(True))
        with pytest.raises(RuntimeError, match="forward-only"):
            blk.sample(coords)
        with pytest.raises(RuntimeError, match="forward-only"):
            cb.sample_features4d(fmaps[:, 0].clone().requires_grad_(True), coords[:, 0])
    with torch.no_grad():   # the same tensors under no_grad are accepted
        blk.corr(targets.clone().requires_grad_(True))
        assert torch.equal(blk.sample(coords), want)
    blk.corr(targets)
    targets.mul_(2.0)
    with pytest.raises(RuntimeError, match="modified in place"):
        blk.sample(coords)
    blk.corr(targets)
    assert rel_to_max(host(blk.sample(coords)), host(want) * 2.0) < 1e-5


def test_extract_patches_clamps_corners_on_non_square_images(cb):
    """corners are clamped per axis (x with W, y with H); the reference assumes H == W."""
    imgs = torch.rand(1, 2, 3, 40, 33, device="cuda")            # H=40, W=33
    tl = torch.tensor([[[[-5, 38], [10, 3]], [[30, 0], [2, 9]]]], device="cuda", dtype=torch.int32)   # (1,2,2,2) = (x, y)
    p = cb.extract_patches(imgs, tl, 31)
    assert p.shape == (4, 3, 31, 31)
    for n in range(2):
        for s in range(2):
            x0 = int(tl[0, s, n, 0].clamp(0, 33 - 31)); y0 = int(tl[0, s, n, 1].clamp(0, 40 - 31))
            assert torch.equal(p[n * 2 + s], imgs[0, s, :, y0:y0 + 31, x0:x0 + 31])


# ------------------------------------------------------------------ fine tracker on the half-resolution source map
def _cl5(t):
    """(B,S,C,H,W) tensor re-laid out channels-last (same shape)."""
    return t.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)


@pytest.mark.parametrize("hs", [16, 9])
def test_up2_lookup_matches_materialised_upsampling_and_oracle(cb, hs):
    """COMET_PYR_UP2_SOURCE: CorrBlock.from_upsampled(S) == CorrBlock(F.interpolate(S, 2Hs-1, bilinear, align_corners))
    -- fcorrs and tokens, queries on, near and off the border (zero padding is applied on each level's own tap grid),
    against the materialised path (itself held to the reference goldens) and against the oracle."""
    import torch.nn.functional as F

    g = torch.Generator(device="cuda").manual_seed(21)
    Bp, S, C = 37, 3, 32
    H = 2 * hs - 1
    src = _cl5(torch.randn(Bp, S, C, hs, hs, device="cuda", generator=g))
    full = F.interpolate(src.reshape(Bp * S, C, hs, hs), (H, H), mode="bilinear", align_corners=True).reshape(Bp, S, C, H, H)
    feats = torch.randn(Bp, S, 1, C, device="cuda", generator=g)
    coords = torch.rand(Bp, S, 1, 2, device="cuda", generator=g) * (H + 5) - 3
    coords[0, 0, 0] = torch.tensor([0.0, H - 1.0], device="cuda")
    coords[1, 0, 0] = torch.tensor([H - 0.5, -0.5], device="cuda")
    coords[2, 0, 0] = torch.tensor([H - 1.0, H - 1.0], device="cuda")
    coords[3, 0, 0] = torch.tensor([-50.0, 1e9], device="cuda")
    up = cb.CorrBlock.from_upsampled(src, num_levels=3, radius=3)
    assert type(up._pyr).__name__ == "_PyramidUp2"
    ref = cb.CorrBlock(_cl5(full), num_levels=3, radius=3)
    up.corr(feats)
    ref.corr(feats)
    a, b = up.sample(coords), ref.sample(coords)
    assert rel_to_max(host(a), host(b)) < 1e-5
    want = O.corr_lookup(host(full), host(feats), host(coords), 3, 3)
    assert rel_to_max(host(a), want) < FP32_BAR
    tdim = cb.transformer_dim(3, 3, C, True)
    xa = cb.TrackTokenizer(up, coords[:, 0], tdim).tokens(coords, feats)
    xb = cb.TrackTokenizer(ref, coords[:, 0], tdim).tokens(coords, feats)
    assert rel_to_max(host(xa), host(xb)) < 1e-5
    # the attributes the reference exposes still work (materialised lazily)
    assert [tuple(l.shape[-2:]) for l in up.fmaps_pyramid] == [tuple(l.shape[-2:]) for l in ref.fmaps_pyramid]
    assert rel_to_max(host(up.fmaps_pyramid[2]), host(ref.fmaps_pyramid[2])) < 1e-6
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ab, bb = up.sample(coords), ref.sample(coords)
    assert rel_to_max(host(ab), host(bb)) < BF16_BAR
    # shapes the specialised kernel does not serve fall back to the materialised tensor
    other = cb.CorrBlock.from_upsampled(src, num_levels=2, radius=2)
    assert type(other._pyr).__name__ == "_Pyramid"
    other.corr(feats)
    assert rel_to_max(host(other.sample(coords)), O.corr_lookup(host(full), host(feats), host(coords), 2, 2)) < FP32_BAR


def test_up2_pyramid_level2_matches_pooled_upsampling(cb):
    import torch.nn.functional as F

    src = _cl5(torch.randn(5, 2, 32, 16, 16, device="cuda"))
    p = cb.CorrBlock.from_upsampled(src, num_levels=3, radius=3)._pyr
    got = p.pyr[: 10 * 7 * 7 * 32].view(10, 7, 7, 32).permute(0, 3, 1, 2)
    full = F.interpolate(src.reshape(10, 32, 16, 16), (31, 31), mode="bilinear", align_corners=True)
    want = F.avg_pool2d(F.avg_pool2d(full, 2, 2), 2, 2)
    assert rel_to_max(host(got), host(want)) < 1e-6

def test_tensor_path_output_modes_agree(cb):
    """Token rows of the tcgen05 path: windows added by bulk reductions onto the pre-kernel's rows (default for 16-byte
    aligned rows) == windows stored entry by entry (COMET_OPT_TC_REDUCE_STORE off, or a token buffer that is not 16-byte
    aligned), bit for bit -- float32 and autocast, N not a multiple of the tile, queries off the map and non-finite."""
    if not cb._lib.lib.comet_has_tensor_path():
        pytest.skip("no sm_100 tensor path on this device")
    g = torch.Generator(device="cuda").manual_seed(11)
    for (B, S, N, L, r, autocast) in ((2, 3, 200, 5, 4, False), (1, 2, 129, 3, 2, False), (1, 4, 512, 5, 4, True), (1, 1, 7, 2, 1, False)):
        fmaps = torch.randn(B, S, 128, 64, 64, device="cuda", generator=g)
        feats = torch.randn(B, S, N, 128, device="cuda", generator=g)
        coords = torch.rand(B, S, N, 2, device="cuda", generator=g) * 80 - 8
        coords[0, 0, 0] = torch.tensor([float("nan"), 5.0], device="cuda")
        coords[0, 0, 1] = torch.tensor([1.0e9, -1.0e9], device="cuda")
        # the reference's token width where it holds the channels, else the channel count rounded up to a multiple of 4
        tdim = max(cb.transformer_dim(L, r, 128, False), (2 * 128 + 2 + L * (2 * r + 1) ** 2 + 3) // 4 * 4)
        ctx = torch.autocast("cuda", dtype=torch.bfloat16) if autocast else contextlib.nullcontext()
        with ctx:
            blk = cb.CorrBlock(fmaps, num_levels=L, radius=r)
            assert blk._pyr.split is not None
            tk = cb.TrackTokenizer(blk, coords[:, 0].nan_to_num(0.0).clamp(0, 63), tdim)
            red = tk.tokens(coords, feats).clone()
            cb._lib.set_option(cb._lib.OPT_TC_REDUCE_STORE, False)
            try:
                stored = tk.tokens(coords, feats).clone()
            finally:
                cb._lib.set_option(cb._lib.OPT_TC_REDUCE_STORE, True)
            # a token buffer whose rows start 4 bytes off a 16-byte boundary: the library must take the store path itself
            flat = torch.empty(B * N * S * tdim + 1, device="cuda")
            off = flat[1:].view(B, N, S, tdim)
            assert off.data_ptr() % 16 != 0
            tk.tokens(coords, feats, out=off)
        assert torch.equal(red.nan_to_num(1234.5), stored.nan_to_num(1234.5))
        assert torch.equal(red.nan_to_num(1234.5), off.nan_to_num(1234.5))
        assert cb._lib.lib.comet_tc_status() == 0


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(3, 64, 128, 128), (2, 5, 40, 30), (1, 2, 200, 200)])
def test_instance_norm_large_planes_and_bf16(cb, shape):
    """One CTA per plane (parked in shared memory up to 128x128 positions, re-read beyond): float32 against torch, and
    the bf16 in / bf16 out form (float32 statistics of the bf16 values) against torch's own bf16 instance norm."""
    g = torch.Generator(device="cuda").manual_seed(8)
    x = torch.randn(*shape, device="cuda", generator=g) * 2.5 + 0.7
    norm = torch.nn.InstanceNorm2d(shape[1])
    for relu in (False, True):
        want = norm(x.double())
        want = torch.relu(want) if relu else want
        got = cb.instance_norm(x, relu=relu)
        assert got.dtype == torch.float32 and rel_to_max(host(got), host(want)) < 2e-6
        xb = x.to(torch.bfloat16)
        wb = norm(xb.double())
        wb = torch.relu(wb) if relu else wb
        gb = cb.instance_norm(xb, relu=relu)
        assert gb.dtype == torch.bfloat16 and gb.shape == xb.shape
        assert rel_to_max(host(gb.float()), host(wb)) < 5e-3          # one bf16 rounding of the result
        gcl = cb.instance_norm(xb.contiguous(memory_format=torch.channels_last), relu=relu)
        assert gcl.dtype == torch.bfloat16 and rel_to_max(host(gcl.float()), host(wb)) < 5e-3


@pytest.mark.gpu
def test_bf16_resize_is_one_rounding_away_from_the_exact_resize(cb):
    """bf16 in / bf16 out resize: float32 arithmetic on the bf16 values, one rounding of the result (the float32 kernel
    between two casts gives the same values up to the contraction of the interpolation's multiply-adds)."""
    import torch.nn.functional as F

    g = torch.Generator(device="cuda").manual_seed(9)
    x = (torch.randn(3, 7, 20, 33, device="cuda", generator=g) * 3).to(torch.bfloat16)
    got = cb.upsample_bilinear_align_corners(x, (41, 50))
    assert got.dtype == torch.bfloat16 and tuple(got.shape) == (3, 7, 41, 50)
    exact = F.interpolate(x.double(), (41, 50), mode="bilinear", align_corners=True)
    err = (got.double() - exact).abs()
    assert bool((err <= exact.abs() * 2.0 ** -8 + 5e-5).all())   # a bf16 rounding + the float32 noise of ATen's source-index arithmetic
    between = cb.upsample_bilinear_align_corners(x.float(), (41, 50)).to(torch.bfloat16)
    assert float((got != between).float().mean()) < 0.01                      # the odd tie broken the other way
