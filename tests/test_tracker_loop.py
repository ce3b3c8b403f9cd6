"""The refinement loop (tokens -> update transformer -> state update) against the reference's BaseTrackerPredictor
outputs (tests/golden/tracker.npz: random-init weights saved as the reference's own state dict).

CPU part: the oracle loop driven by this package's torch EfficientUpdateFormer -> pins the loop restatement and the
transformer plumbing (state-dict compatibility included).  GPU part: the drop-in BaseTrackerPredictor on CUDA."""
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

import cases
from conftest import rel_to_max
from oracle import comet_oracle as O

SPECS = {
    "coarse_tiny": (dict(stride=4, corr_levels=5, corr_radius=2, latent_dim=16, hidden_size=32, depth=1,
                         use_spaceatt=True, fine=False),
                    dict(seed=41, B=1, S=4, C=16, H=16, W=16, N=7, stride=4, down_ratio=2), 3, 2, False),
    "coarse_tiny_eff": (dict(stride=4, corr_levels=2, corr_radius=3, latent_dim=16, hidden_size=32, depth=1,
                             use_spaceatt=True, fine=False),
                        dict(seed=42, B=1, S=4, C=16, H=16, W=16, N=7, stride=4, down_ratio=2), 2, 2, True),
    "fine_tiny": (dict(stride=1, corr_levels=3, corr_radius=3, latent_dim=32, hidden_size=32, depth=1,
                       use_spaceatt=False, fine=True),
                  dict(seed=43, B=5, S=3, C=32, H=31, W=31, N=1, stride=1, down_ratio=1), 2, 1, False),
}


# Per-iteration bars.  From iteration 1 on the tokens contain sin/cos(flow * k * 1000/C) (utils.py:84-96): rounding
# differences in the previous iteration's coordinates are multiplied by up to ~1e3 inside the sine argument, so ANY two
# float32 implementations drift apart.  How much is not asserted but MEASURED: tracker.npz holds the reference's own
# float32 and float64 runs of every case, and an implementation passes iteration i when
#     |impl - ref64| / max|ref64| <= max(1e-4, 3 * |ref32 - ref64| / max|ref64|)
# -- the north-star float32 tolerance, or three times the reference's own float32 noise where that is larger.  Every
# kernel is held to 1e-4 on identical inputs by the other tests; tests/test_full_size.py applies the same criterion at
# the shipped sizes.
SPEC = 1e-4


def bar(g, name, what, i=None):
    k32 = f"{name}/{what}{i}" if i is not None else f"{name}/{what}"
    k64 = f"{name}/{what}64_{i}" if i is not None else f"{name}/{what}64"
    return max(SPEC, 3.0 * rel_to_max(g[k32], g[k64])), g[k64]


def cfg(eff):
    return NS(track_conf=False, MODEL=NS(TRACK=NS(efficient_corr=eff)))


def load_reference_weights(module, g, name):
    sd = {k[len(name) + 4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(name + "/sd/")}
    missing, unexpected = module.load_state_dict(sd, strict=True), None
    return module


def build(name):
    from comet_pose_estimation_b200.base_track_predictor import BaseTrackerPredictor  # imports the built library

    ck, case_kw, iters, dr, eff = SPECS[name]
    m = BaseTrackerPredictor(cfg=cfg(eff), **ck).eval()
    return m, case_kw, iters, dr, eff, ck


@pytest.mark.parametrize("name", list(SPECS))
def test_state_dict_keys_match_reference(golden, name):
    g = golden("tracker")
    m, *_ = build(name)
    want = sorted(k[len(name) + 4:] for k in g.files if k.startswith(name + "/sd/"))
    assert sorted(m.state_dict().keys()) == want
    load_reference_weights(m, g, name)  # strict


@pytest.mark.parametrize("name", list(SPECS))
def test_oracle_loop_with_torch_updateformer(golden, name):
    g = golden("tracker")
    m, case_kw, iters, dr, eff, ck = build(name)
    load_reference_weights(m, g, name)
    fmaps, q = cases.tracker_case(**case_kw)
    with torch.no_grad():
        preds, feats, qfeat, toks = O.tracker_forward(
            q, fmaps,
            lambda x: m.updateformer(torch.from_numpy(x)).numpy(),
            lambda d: m.ffeat_updater(m.norm(torch.from_numpy(d))).numpy(),
            iters=iters, stride=ck["stride"], corr_levels=ck["corr_levels"], corr_radius=ck["corr_radius"],
            latent_dim=ck["latent_dim"], fine=ck["fine"], down_ratio=dr, efficient_corr=eff)
    for i in range(iters):
        b, ref = bar(g, name, "tok", i)
        assert rel_to_max(toks[i], ref) <= b, (i, "tokens")
        b, ref = bar(g, name, "pred", i)
        assert rel_to_max(preds[i], ref) <= b, (i, "tracks")
    b, ref = bar(g, name, "track_feats")
    assert rel_to_max(feats, ref) <= b
    assert rel_to_max(qfeat, g[name + "/query_feat"]) < 1e-5


def test_fixture_holds_the_reference_drift(golden):
    """The fixtures demonstrate the drift statement above: iteration 0 of the reference agrees between float32 and
    float64 to ~1e-7, and the difference grows by orders of magnitude per iteration."""
    g = golden("tracker")
    d = [rel_to_max(g[f"coarse_tiny/pred{i}"], g[f"coarse_tiny/pred64_{i}"]) for i in range(3)]
    assert d[0] < 1e-6 and d[2] > 20 * d[0]


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(SPECS))
def test_dropin_predictor_matches_reference(golden, name):
    """Predicted tracks check-point of BASELINE.md section 5 (fp32 bar 1e-4)."""
    g = golden("tracker")
    m, case_kw, iters, dr, eff, ck = build(name)
    load_reference_weights(m, g, name)
    m = m.cuda()
    fmaps, q = cases.tracker_case(**case_kw)
    toks = []
    h = m.updateformer.register_forward_pre_hook(lambda mod, a: toks.append(a[0].detach().cpu().numpy()))
    with torch.no_grad():
        preds, vis, feats, qfeat, conf = m(query_points=torch.from_numpy(q).cuda(), fmaps=torch.from_numpy(fmaps).cuda(),
                                           iters=iters, return_feat=True, down_ratio=dr, TRACKorPOSE=False)
    h.remove()
    assert conf is None and len(preds) == iters
    for i in range(iters):
        b, ref = bar(g, name, "tok", i)
        assert rel_to_max(toks[i], ref) <= b, (i, "tokens")
        b, ref = bar(g, name, "pred", i)
        assert rel_to_max(preds[i].cpu().numpy(), ref) <= b, (i, "tracks")
    b, ref = bar(g, name, "track_feats")
    assert rel_to_max(feats.cpu().numpy(), ref) <= b
    assert rel_to_max(qfeat.cpu().numpy(), g[name + "/query_feat"]) < 1e-5
    if ck["fine"]:
        assert vis is None
    else:
        b, ref = bar(g, name, "vis")
        assert rel_to_max(vis.cpu().numpy(), ref) <= b


@pytest.mark.gpu
def test_dropin_predictor_coarse_shape_uses_tensor_path():
    """Full coarse configuration (C=128, 64x64, L=5, r=4): tcgen05 path and SIMT path give the same tracks."""
    import comet_pose_estimation_b200 as cb

    torch.manual_seed(0)
    m = cb.BaseTrackerPredictor(cfg=cfg(False), hidden_size=64, depth=2).eval().cuda()
    fmaps = torch.randn(1, 4, 128, 64, 64, device="cuda")
    q = torch.rand(1, 96, 2, device="cuda") * 480 + 16
    with torch.no_grad():
        a = m(query_points=q, fmaps=fmaps, iters=3, down_ratio=2, TRACKorPOSE=False)[0]
        cb._lib.set_option(cb._lib.OPT_TENSOR_PATH, False)
        try:
            b = m(query_points=q, fmaps=fmaps, iters=3, down_ratio=2, TRACKorPOSE=False)[0]
        finally:
            cb._lib.set_option(cb._lib.OPT_TENSOR_PATH, True)
    # two float32 implementations of the same loop: iteration 0 within the spec, later ones bounded by the drift the
    # reference shows between its own float32 and float64 runs at this size (tracker_full.npz: 2.5e-5, 5.7e-3)
    for i, (x, y) in enumerate(zip(a, b)):
        assert rel_to_max(x.cpu().numpy(), y.cpu().numpy()) < [1e-4, 1e-3, 5e-2][i]
