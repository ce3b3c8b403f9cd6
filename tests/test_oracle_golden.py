"""Pin the CPU oracle (oracle/comet_oracle.py) to tensors produced by the
unmodified reference (tests/golden/*.npz, made by tests/golden/make_golden.py).

Tolerances: fp32 1e-4 relative-to-max is the product bar; the oracle itself is
held to 2e-6 so that it leaves the whole budget to the kernels."""
import hashlib

import numpy as np
import pytest

import cases
from conftest import rel_to_max
from oracle import comet_oracle as O

TIGHT = 2e-6


def _digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:16]


@pytest.mark.parametrize("name", list(cases.CORR_CASES))
def test_corr_lookup_matches_reference(golden, name):
    g = golden("corr_blocks")
    kw, L, r = cases.CORR_CASES[name]
    fmaps, targets, coords = cases.corr_case(**kw)
    assert _digest(fmaps, targets, coords) == bytes(g[name + "/digest"]).decode(), "input generator drifted"
    assert rel_to_max(O.corr_lookup(fmaps, targets, coords, L, r, "zeros"), g[name + "/zeros"]) < TIGHT
    assert rel_to_max(O.corr_lookup(fmaps, targets, coords, L, r, "border"), g[name + "/border"]) < TIGHT
    assert rel_to_max(O.efficient_corr_lookup(fmaps, targets, coords, L, r), g[name + "/efficient"]) < TIGHT


@pytest.mark.parametrize("name", ["small_ragged", "tiny_odd"])
def test_pyramid_and_volumes(golden, name):
    g = golden("corr_blocks")
    kw, L, r = cases.CORR_CASES[name]
    fmaps, targets, _ = cases.corr_case(**kw)
    pyr = O.build_pyramid(fmaps, L)
    vols = O.corr_volumes(targets, pyr)
    for l in range(L):
        assert pyr[l].shape == g[f"{name}/pyr{l}"].shape
        assert rel_to_max(pyr[l], g[f"{name}/pyr{l}"]) < TIGHT
        assert rel_to_max(vols[l], g[f"{name}/vol{l}"]) < TIGHT


@pytest.mark.parametrize("name", list(cases.CORR_CASES))
def test_bf16_autocast_semantics(golden, name):
    g = golden("corr_blocks")
    kw, L, r = cases.CORR_CASES[name]
    fmaps, targets, coords = cases.corr_case(**kw)
    got = O.corr_lookup_bf16_autocast(fmaps, targets, coords, L, r)
    # bf16 rounding points restated exactly -> far inside the 2e-2 product bar
    assert rel_to_max(got, g[name + "/zeros_bf16"]) < 4e-3
    # and the reference's own bf16 path sits inside the stated bf16 bar vs fp32
    assert rel_to_max(g[name + "/zeros_bf16"], g[name + "/zeros"]) < 2e-2


def test_transposed_window_quirk():
    """a correlation peak at (x+1, y) lands at flat index
    (r+1)*Wr + r of the window (x offset is the slow index)."""
    C, H, W, r = 1, 9, 9, 2
    fmaps = np.zeros((1, 1, C, H, W), np.float32)
    fmaps[0, 0, 0, 4, 5] = 1.0  # y=4, x=5
    targets = np.ones((1, 1, 1, C), np.float32)
    coords = np.array([[[[4.0, 4.0]]]], np.float32)
    out = O.corr_lookup(fmaps, targets, coords, 1, r)[0, 0, 0]
    Wr = 2 * r + 1
    assert np.argmax(out) == (r + 1) * Wr + r


def test_bilinear_sampler(golden):
    g = golden("samplers_encodings")
    inp, xy = cases.sampler_case(21, B=2, C=3, H=9, W=7, Ho=4, Wo=5)
    for pm in ("zeros", "border"):
        for ac in (True, False):
            got = O.bilinear_sampler(inp, xy, align_corners=ac, padding_mode=pm)
            assert rel_to_max(got, g[f"bs4/{pm}/{int(ac)}"]) < TIGHT
    inp5, txy = cases.sampler_case(22, B=2, C=3, H=6, W=8, Ho=3, Wo=4, T=3)
    for pm in ("zeros", "border"):
        assert rel_to_max(O.bilinear_sampler(inp5, txy, padding_mode=pm), g[f"bs5/{pm}"]) < TIGHT
    inp1, txy1 = cases.sampler_case(23, B=2, C=4, H=6, W=8, Ho=3, Wo=4, T=1)
    txy1[..., 0] = 0
    assert rel_to_max(O.bilinear_sampler(inp1, txy1), g["bs5_t1/border"]) < TIGHT
    assert rel_to_max(O.sample_features4d(inp, g["sf4d/pts"]), g["sf4d/out"]) < TIGHT


def test_encodings(golden):
    g = golden("samplers_encodings")
    for C, scale in ((64, 3.0), (16, 0.7), (64, 40.0)):
        xy = g[f"emb2d/{C}/{scale}/xy"]
        # arguments reach |xy|*1000: sin/cos of the same float32 argument, <= few ulp apart
        assert np.abs(O.get_2d_embedding(xy, C, False) - g[f"emb2d/{C}/{scale}/nocat"]).max() < 1e-6
        assert np.abs(O.get_2d_embedding(xy, C, True) - g[f"emb2d/{C}/{scale}/cat"]).max() < 1e-6
    assert np.array_equal(O.get_2d_sincos_pos_embed(216, (31, 31)), g["sincos2d/216_31"])
    full = O.get_2d_sincos_pos_embed(664, (64, 64))
    assert full.shape == (1, 664, 64, 64)
    assert np.abs(full[:, :, ::9, :] - g["sincos2d/664_64/rows"]).max() < 1e-7
    assert abs(full.astype(np.float64).sum() - g["sincos2d/664_64/sum"][0]) < 1e-2
    assert np.abs(O.get_2d_sincos_pos_embed(12, (3, 5)) - g["sincos2d/12_h3w5"]).max() < 1e-7
    pe, grid = O.get_2d_sincos_pos_embed(8, 4, return_grid=True)
    assert np.abs(pe - g["sincos2d/8_4"]).max() < 1e-7 and np.array_equal(grid, g["sincos2d/8_4/grid"])
    assert np.abs(O.get_1d_sincos_pos_embed(768, 16) - g["sincos1d/768_16"]).max() < 1e-7
    assert np.abs(O.get_1d_sincos_pos_embed(768, 64) - g["sincos1d/768_64"]).max() < 1e-7
    pe1, g1 = O.get_1d_sincos_pos_embed(10, 7, return_grid=True)
    assert np.abs(pe1 - g["sincos1d/10_7"]).max() < 1e-7 and np.array_equal(g1, g["sincos1d/10_7/grid"])
    assert np.abs(O.get_1d_sincos_pos_embed_from_grid(14, g["sincos1dgrid/pos"]) - g["sincos1dgrid/14"]).max() < 1e-7


TOKEN_CASES = {
    # name: (case kwargs, L, r, latent, fine, down_ratio, stride)
    "coarse_full_it0": (dict(seed=44, B=1, S=3, C=128, H=64, W=64, N=16, stride=4, down_ratio=2), 5, 4, 128, False, 2, 4),
    "fine_full_it0": (dict(seed=45, B=8, S=3, C=32, H=31, W=31, N=1, stride=1, down_ratio=1), 3, 3, 32, True, 1, 1),
}


@pytest.mark.parametrize("name", list(TOKEN_CASES))
def test_first_iteration_tokens_full_size(golden, name):
    g = golden("tracker")
    kw, L, r, latent, fine, dr, stride = TOKEN_CASES[name]
    fmaps, q = cases.tracker_case(**kw)
    assert _digest(fmaps, q) == bytes(g[name + "/digest"]).decode()

    class Stop(Exception):
        pass

    def upd(x):
        raise Stop

    toks = []
    try:
        O.tracker_forward(q, fmaps, lambda x: (toks.append(x), upd(x))[1], None, iters=1, stride=stride,
                          corr_levels=L, corr_radius=r, latent_dim=latent, fine=fine, down_ratio=dr)
    except Stop:
        pass
    want = g[name + "/tok0"]
    assert toks[0].shape == want.shape
    assert O.transformer_dim(L, r, latent, fine) == want.shape[-1]
    assert rel_to_max(toks[0], want) < 1e-5


# ---- the torch CPU port (what bench.py times as the CPU baseline) is pinned to the same goldens ----
@pytest.mark.parametrize("name", list(cases.CORR_CASES))
def test_torch_port_corr_lookup(golden, name):
    import torch

    from oracle import torch_port as P

    g = golden("corr_blocks")
    kw, L, r = cases.CORR_CASES[name]
    fmaps, targets, coords = (torch.from_numpy(a) for a in cases.corr_case(**kw))
    lv = P.pyramid(fmaps, L)
    for pm in ("zeros", "border"):
        got = P.lookup(P.volumes(lv, targets), coords, r, pm).numpy()
        assert rel_to_max(got, g[f"{name}/{pm}"]) < TIGHT


@pytest.mark.parametrize("name", list(TOKEN_CASES))
def test_torch_port_tokens(golden, name):
    import torch

    from oracle import torch_port as P

    g = golden("tracker")
    kw, L, r, latent, fine, dr, stride = TOKEN_CASES[name]
    fmaps, q = (torch.from_numpy(a) for a in cases.tracker_case(**kw))
    if dr > 1:
        q = q / float(dr) / float(stride)
    B, S = fmaps.shape[:2]
    coords = q[:, None].repeat(1, S, 1, 1)
    feats = P.point_sample(fmaps[:, 0], coords[:, 0])[:, None].repeat(1, S, 1, 1)
    tdim = O.transformer_dim(L, r, latent, fine)
    x = P.hot_path_iteration(P.pyramid(fmaps, L), coords, feats, r, fmaps.shape[-2:], tdim)
    assert rel_to_max(x.numpy(), g[name + "/tok0"]) < 1e-5
