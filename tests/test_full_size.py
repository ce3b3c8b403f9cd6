"""Shipped-size parity of the tracker loop, refine_track and the final pose (BASELINE.md section 5 check-points
"predicted tracks" and "final pred_pose_enc"), against fixtures produced by executing the reference in float32 AND
float64 (tests/golden/make_golden_full.py; weights rebuilt by name from ``cases.seeded_state_dict``).

The bar.  From the second refinement iteration on the tokens contain sin/cos(flow * k * 1000/C) (utils.py:84-96), so
rounding differences are amplified ~200x per iteration -- measured on the reference itself: its float32 run differs
from its float64 run by 1.4e-7 / 2.5e-5 / 5.7e-3 / 1.8e-2 (relative to max) after coarse iterations 0..3.  A
conforming float32 implementation therefore has to satisfy, per iteration,

    |impl - ref64| / max|ref64|  <=  max(1e-4, 3 * |ref32 - ref64| / max|ref64|)

i.e. the north-star float32 tolerance, or three times the reference's own float32 noise where that is larger.
"""
import json
import os
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

import cases
from conftest import ROOT, rel_to_max

SPEC = 1e-4


def cfg():
    return NS(track_conf=False, MODEL=NS(TRACK=NS(efficient_corr=False)))


def load_seeded(module, seed):
    sd = cases.seeded_state_dict({k: tuple(v.shape) for k, v in module.state_dict().items()}, seed)
    module.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    return module


def bar(ref32, ref64):
    return max(SPEC, 3.0 * rel_to_max(ref32, ref64))


def _record(name, rows):
    """Measured numbers next to their bars, for profiles/ (written on the GPU box under gpurun_out/)."""
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    path = os.path.join(out, "parity_full.json")
    try:
        with open(path) as f:
            d = json.load(f)
    except Exception:
        d = {}
    d[name] = rows
    with open(path, "w") as f:
        json.dump(d, f, indent=1)


# ------------------------------------------------------------------------------------------------ CPU: fixtures
def test_seeded_weights_are_name_keyed_and_stable():
    a = cases.seeded_state_dict({"x.weight": (4, 3), "x.bias": (4,), "n.weight": (5,)}, 7)
    b = cases.seeded_state_dict({"n.weight": (5,), "x.bias": (4,), "x.weight": (4, 3), "extra.weight": (2, 2)}, 7)
    for k in a:
        assert np.array_equal(a[k], b[k])          # independent of the other entries / of the order
    assert abs(float(a["x.weight"].max())) <= 1 / np.sqrt(3) + 1e-6
    assert abs(float(a["n.weight"].mean()) - 1.0) < 0.3


def test_full_fixtures_document_reference_fp32_drift(golden):
    """The yard-stick itself: the reference's float32 run drifts from its float64 run by orders of magnitude more than
    1e-4 after two iterations (the statement DESIGN.md section 3 makes), while the final pose barely moves."""
    g = golden("tracker_full")
    d = [rel_to_max(g[f"pred32_{i}"], g[f"pred64_{i}"]) for i in range(cases.FULL_COARSE_ITERS)]
    assert d[0] < 1e-6 and d[1] < SPEC          # iterations 0 and 1: float32 noise is still below the spec
    assert d[2] > SPEC and d[3] > SPEC          # from iteration 2 on the reference itself exceeds 1e-4
    assert all(b > a for a, b in zip(d, d[1:]))
    r = golden("refine_full")
    assert rel_to_max(r["refined32"], r["refined64"]) > SPEC
    p = golden("pose")
    for k in ("pred_pose_enc", "R", "T"):
        assert rel_to_max(p["ref64/" + k], p["ref32/" + k]) < SPEC


def test_pose_checkpoint_with_cuda_tracks(golden):
    """Final-pose check-point: the reference CameraPredictor fed with tracks / confidence produced by this repository's
    CUDA path on a B200 (tests/golden/cuda_tracks_full.npz, written by the GPU test below and run through
    make_golden_full.py --pose-from in the build container) gives the reference's pose within 1e-4."""
    p = golden("pose")
    if "cuda/pred_pose_enc" not in p.files:
        pytest.skip("pose.npz has no CUDA arm yet (run the GPU test, then make_golden_full.py --pose-from)")
    for k in ("pred_pose_enc", "R", "T"):
        # tracker isolated (refine_track fed with the reference's coarse prediction): the north-star 1e-4
        assert rel_to_max(p["cuda/" + k], p["ref32/" + k]) < SPEC, k
        # whole chain (own coarse prediction, which carries the loop's drift): 1e-4, or three times how far the
        # reference's own float64 chain moves the pose away from its float32 chain (measured 3.4e-4)
        assert rel_to_max(p["cuda_chain/" + k], p["ref32/" + k]) <= bar(p["ref64_chain/" + k], p["ref32/" + k]), k
    # frame 0 is the identity transform by construction (camera_predictor10.py:457-460)
    assert np.allclose(p["cuda/pred_pose_enc"][0], [0, 0, 0, 1, 0, 0, 0])


def test_inverted_score_matches_reference(golden):
    from comet_pose_estimation_b200.refine_track import inverted_score

    r = golden("refine_full")
    got = inverted_score(torch.from_numpy(r["score32"])).numpy()
    assert rel_to_max(got, r["inverted32"]) < 1e-6


# ------------------------------------------------------------------------------------------------ GPU
def _coarse_model():
    import comet_pose_estimation_b200 as cb

    return load_seeded(cb.BaseTrackerPredictor(cfg=cfg(), **cases.FULL_COARSE_CTOR).eval(), cases.FULL_SEEDS["coarse"]).cuda()


def _fine_models():
    import comet_pose_estimation_b200 as cb

    fnet = load_seeded(cb.ShallowEncoder(input_dim=3).eval(), cases.FULL_SEEDS["fnet"]).cuda()
    ftr = load_seeded(cb.BaseTrackerPredictor(cfg=cfg(), **cases.FULL_FINE_CTOR).eval(), cases.FULL_SEEDS["fine"]).cuda()
    return fnet.to(memory_format=torch.channels_last), ftr


def _run_coarse(m):
    fmaps, q = cases.tracker_case(**cases.FULL_COARSE_CASE)
    with torch.no_grad():
        return m(query_points=torch.from_numpy(q).cuda(), fmaps=torch.from_numpy(fmaps).cuda(),
                 iters=cases.FULL_COARSE_ITERS, return_feat=True, down_ratio=cases.FULL_COARSE_CASE["down_ratio"],
                 TRACKorPOSE=False)


@pytest.fixture
def no_tf32():
    a, b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = a, b


@pytest.mark.gpu
def test_full_coarse_tracker_against_fp64_reference(golden, no_tf32):
    """hidden 384, depth 6, S=16, N=512, 4 iterations (abl_ours.yaml:99, 399-412) through the drop-in predictor."""
    g = golden("tracker_full")
    preds, vis, feats, qfeat, _ = _run_coarse(_coarse_model())
    rows = []
    for i in range(cases.FULL_COARSE_ITERS):
        r32, r64 = g[f"pred32_{i}"], g[f"pred64_{i}"]
        e = rel_to_max(preds[i].cpu().numpy(), r64)
        rows.append({"iteration": i, "impl_vs_ref64": e, "ref32_vs_ref64": rel_to_max(r32, r64), "bar": bar(r32, r64),
                     "impl_vs_ref32": rel_to_max(preds[i].cpu().numpy(), r32)})
    _record("coarse_tracker_full", rows)
    for r in rows:
        assert r["impl_vs_ref64"] <= r["bar"], rows
    assert rel_to_max(qfeat.cpu().numpy(), g["query_feat32"]) < 1e-5
    assert rel_to_max(vis.cpu().numpy(), g["vis64"]) <= bar(g["vis32"], g["vis64"])
    assert rel_to_max(feats[:, :, ::37].cpu().numpy(), g["track_feats64_slice"]) <= \
        bar(g["track_feats32_slice"], g["track_feats64_slice"])


@pytest.mark.gpu
def test_full_refine_track_against_fp64_reference(golden, no_tf32):
    """refine_track at 512 tracks x 16 frames of 512x512 images, fine tracker hidden 256 / depth 4 / 6 iterations
    (abl_ours.yaml:414-428, refine_track.py:136), fed with the reference's coarse prediction."""
    import comet_pose_estimation_b200 as cb

    r = golden("refine_full")
    fnet, ftr = _fine_models()
    images, _ = cases.refine_case(**cases.FULL_REFINE_CASE)
    with torch.no_grad():
        refined, score = cb.refine_track(torch.from_numpy(images).cuda(), fnet, ftr,
                                         torch.from_numpy(r["coarse_pred"]).cuda(), compute_score=True)
        inv = cb.inverted_score(score)
    rows = {}
    for k, got in (("refined", refined), ("score", score), ("inverted", inv)):
        e = rel_to_max(got.cpu().numpy(), r[k + "64"])
        rows[k] = {"impl_vs_ref64": e, "ref32_vs_ref64": rel_to_max(r[k + "32"], r[k + "64"]),
                   "bar": bar(r[k + "32"], r[k + "64"]), "impl_vs_ref32": rel_to_max(got.cpu().numpy(), r[k + "32"])}
    _record("refine_track_full", rows)
    for k, v in rows.items():
        assert v["impl_vs_ref64"] <= v["bar"], rows


@pytest.mark.gpu
def test_full_chain_writes_tracks_for_pose_golden(golden, no_tf32):
    """coarse tracker -> refine_track -> inverted score on the GPU, (a) from the reference's coarse prediction and
    (b) from this path's own coarse prediction; saved for make_golden_full.py --pose-from (pose check-point), and
    compared with the committed copy the pose fixture was computed from."""
    import comet_pose_estimation_b200 as cb

    r = golden("refine_full")
    fnet, ftr = _fine_models()
    images = torch.from_numpy(cases.refine_case(**cases.FULL_REFINE_CASE)[0]).cuda()
    HW = cases.FULL_REFINE_CASE["HW"]
    with torch.no_grad():
        refined, score = cb.refine_track(images, fnet, ftr, torch.from_numpy(r["coarse_pred"]).cuda(), compute_score=True)
        own = _run_coarse(_coarse_model())[0][-1].clamp(0.0, HW - 1.001)
        refined_c, score_c = cb.refine_track(images, fnet, ftr, own, compute_score=True)
        out = dict(refined=refined.cpu().numpy(), inverted=cb.inverted_score(score).cpu().numpy(),
                   refined_chain=refined_c.cpu().numpy(), inverted_chain=cb.inverted_score(score_c).cpu().numpy())
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "cuda_tracks_full.npz"), **out)
    committed = os.path.join(ROOT, "tests", "golden", "cuda_tracks_full.npz")
    if os.path.exists(committed):
        c = np.load(committed)
        # same device code, same inputs: only run-to-run scheduling differences (none expected) amplified by the loop
        assert rel_to_max(out["refined"], c["refined"]) <= bar(r["refined32"], r["refined64"])
        assert rel_to_max(out["refined_chain"], c["refined_chain"]) <= 10 * bar(r["refined32"], r["refined64"])
