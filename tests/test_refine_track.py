"""``refine_track`` (the fine tracker's caller: patch gather -> ShallowEncoder -> fine BaseTrackerPredictor -> score)
against tensors produced by executing the reference (tests/golden/refine.npz).

CPU part: the host logic -- patch extraction in (b, n, s) order, the channels-last encoder mirror with the
reference's state dict, coordinate bookkeeping and ``compute_score_fn`` with the reference's indexing quirk -- driven
by the numpy oracle as the fine tracker.  GPU part: the full drop-in (fused kernels on channels-last patch features),
the "refined tracks / score" check-point of BASELINE.md section 5."""
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

import cases
from conftest import rel_to_max
from oracle import comet_oracle as O

NAME = "refine_small"
CASE = dict(seed=51, B=1, S=3, N=5, HW=64)
FINE = dict(stride=1, corr_levels=3, corr_radius=3, latent_dim=32, hidden_size=32, depth=1, use_spaceatt=False, fine=True)


# tokens of iteration i >= 1 contain sin/cos(flow * k * 1000/16): a 1e-6 px difference in the previous iteration's
# coordinates is ~1e-3 in those channels (see test_tracker_loop.py); the tracks and scores themselves are held to 1e-4
TOK_BAR = 5e-3


def cfg():
    return NS(track_conf=False, MODEL=NS(TRACK=NS(efficient_corr=False)))


def modules(g):
    from comet_pose_estimation_b200.base_track_predictor import BaseTrackerPredictor
    from comet_pose_estimation_b200.refine_track import ShallowEncoder

    fnet = ShallowEncoder(input_dim=3).eval()
    ftr = BaseTrackerPredictor(cfg=cfg(), **FINE).eval()
    for mod, tag in ((fnet, "fnet"), (ftr, "ftr")):
        pre = f"{NAME}/{tag}/"
        sd = {k[len(pre):]: torch.from_numpy(g[k]) for k in g.files if k.startswith(pre)}
        assert sorted(sd) == sorted(mod.state_dict().keys())  # same parameter names as the reference
        mod.load_state_dict(sd, strict=True)
    return fnet, ftr


def test_shallow_encoder_matches_reference(golden):
    g = golden("refine")
    fnet, _ = modules(g)
    x = torch.from_numpy(g[NAME + "/enc_in"])
    with torch.no_grad():
        y = fnet(x)
        y_cl = fnet.to(memory_format=torch.channels_last)(x.contiguous(memory_format=torch.channels_last))
    assert rel_to_max(y.numpy(), g[NAME + "/enc_out"]) < 1e-5
    assert rel_to_max(y_cl.numpy(), g[NAME + "/enc_out"]) < 1e-5


def test_fused_encoder_parameter_order_is_the_state_dict_order():
    """comet_shallow_encoder_pack_f32 takes the 16 parameter tensors in state-dict order: the mirror's list must be it."""
    from comet_pose_estimation_b200.refine_track import ShallowEncoder

    fnet = ShallowEncoder(input_dim=3)
    want = [f"{n}.{s}" for n in fnet._PARAM_ORDER for s in ("weight", "bias")]
    assert want == list(fnet.state_dict().keys())
    shapes = [tuple(fnet.state_dict()[k].shape) for k in want[::2]]
    assert shapes == [(32, 3, 3, 3), (32, 32, 3, 3), (32, 32, 3, 3), (32, 32, 1, 1),
                      (32, 32, 3, 3), (32, 32, 3, 3), (32, 32, 1, 1), (32, 32, 1, 1)]
    # CPU tensors, other patch sizes, training: the per-operator path
    x = torch.zeros(2, 3, 31, 31)
    assert not fnet._fused_ok(x, 31, 31)


def test_patch_extraction_order_and_layout():
    from comet_pose_estimation_b200.refine_track import extract_patches

    images, coarse = cases.refine_case(**CASE)
    B, S, N = coarse.shape[:3]
    tl = (np.floor(coarse).astype(np.int32) - 15).clip(0, 64 - 31)
    p = extract_patches(torch.from_numpy(images), torch.from_numpy(tl), 31)
    assert p.shape == (B * N * S, 3, 31, 31)
    assert p.permute(0, 2, 3, 1).is_contiguous()  # channels-last memory
    for (b, n, s) in ((0, 0, 0), (0, 3, 2), (0, 4, 1)):
        x0, y0 = tl[b, s, n]
        want = images[b, s, :, y0:y0 + 31, x0:x0 + 31]
        assert np.array_equal(p[(b * N + n) * S + s].numpy(), want)


def test_refine_track_host_logic_with_oracle_tracker(golden):
    """refine_track + compute_score_fn on CPU; the fine tracker is the numpy oracle driven by this package's torch
    update transformer (reference weights)."""
    from comet_pose_estimation_b200.refine_track import refine_track

    g = golden("refine")
    fnet, ftr = modules(g)
    images, coarse = cases.refine_case(**CASE)
    toks = []

    def oracle_tracker(query_points, fmaps, iters, return_feat, TRACKorPOSE):
        assert not TRACKorPOSE and return_feat
        assert fmaps.permute(0, 1, 3, 4, 2).is_contiguous()  # channels-last view handed to the path, no copy
        preds, feats, qfeat, tk = O.tracker_forward(
            query_points.numpy(), fmaps.numpy(),
            lambda x: ftr.updateformer(torch.from_numpy(x)).numpy(),
            lambda d: ftr.ffeat_updater(ftr.norm(torch.from_numpy(d))).numpy(),
            iters=iters, stride=1, corr_levels=3, corr_radius=3, latent_dim=32, fine=True, down_ratio=1)
        toks.extend(tk)
        return [torch.from_numpy(p) for p in preds], None, torch.from_numpy(feats), torch.from_numpy(qfeat), None

    with torch.no_grad():
        fnet = fnet.to(memory_format=torch.channels_last)
        refined, score = refine_track(torch.from_numpy(images), fnet, oracle_tracker, torch.from_numpy(coarse),
                                      compute_score=True)
    assert rel_to_max(toks[0], g[NAME + "/tok0"]) < 1e-4      # patch gather + encoder + first tokens
    for i in range(1, 6):
        assert rel_to_max(toks[i], g[f"{NAME}/tok{i}"]) < TOK_BAR
    assert refined.shape == g[NAME + "/refined"].shape
    assert np.array_equal(refined[:, 0].numpy(), coarse[:, 0])  # frame 0 is pinned to the query points
    assert rel_to_max(refined.numpy(), g[NAME + "/refined"]) < 1e-4
    assert rel_to_max(score.numpy(), g[NAME + "/score"]) < 1e-4
    assert float(score[:, 0].min()) == 1.0 and float(score[:, 0].max()) == 1.0


def test_compute_score_indexing_quirk(golden):
    """with B == 1 every score is computed from patch (n=0, s=0); randomising every other patch
    leaves the score unchanged."""
    from comet_pose_estimation_b200.refine_track import compute_score_fn

    rng = np.random.default_rng(5)
    B, N, S, C, P = 1, 4, 3, 32, 31
    pf = torch.from_numpy(rng.standard_normal((B * N, S, C, P, P)).astype(np.float32))
    qf = torch.from_numpy(rng.standard_normal((B, N, C)).astype(np.float32))
    tr = torch.from_numpy(rng.uniform(3, 27, (B * N, S, 1, 2)).astype(np.float32))
    a = compute_score_fn(qf, pf, tr, 2, P, B, N, S, C)
    pf2 = pf.clone()
    pf2[1:] = torch.from_numpy(rng.standard_normal((B * N - 1, S, C, P, P)).astype(np.float32))
    pf2[0, 1:] = 0.0
    b = compute_score_fn(qf, pf2, tr, 2, P, B, N, S, C)
    assert torch.equal(a, b)
    assert a.shape == (B, S, N)


@pytest.mark.gpu
def test_patch_gather_kernel_matches_torch_indexing():
    import importlib

    rt = importlib.import_module("comet_pose_estimation_b200.refine_track")
    g = torch.Generator(device="cuda").manual_seed(3)
    images = torch.rand(2, 3, 3, 70, 70, device="cuda", generator=g)
    tl = torch.randint(0, 70 - 31 + 1, (2, 3, 9, 2), device="cuda", generator=g, dtype=torch.int32)
    tl[0, 0, 0] = 0
    tl[1, 2, 8] = 70 - 31
    a = rt.extract_patches(images, tl, 31)
    rt.USE_LIBRARY_KERNELS = False
    try:
        b = rt.extract_patches(images, tl, 31)
    finally:
        rt.USE_LIBRARY_KERNELS = True
    assert a.shape == b.shape == (2 * 9 * 3, 3, 31, 31) and a.stride() == b.stride()
    assert torch.equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("defer", [True, False])
def test_dropin_refine_track_matches_reference(golden, defer):
    """defer=True: the fine tracker reads the encoder's half-resolution map (blocks.Upsampled2x, the default);
    defer=False: the up-sampled patch features are materialised and reach the kernels as a channels-last view."""
    import importlib

    rt = importlib.import_module("comet_pose_estimation_b200.refine_track")
    refine_track = rt.refine_track
    g = golden("refine")
    fnet, ftr = modules(g)
    fnet = fnet.cuda().to(memory_format=torch.channels_last)
    ftr = ftr.cuda()
    images, coarse = cases.refine_case(**CASE)
    toks, layouts = [], []
    h = ftr.updateformer.register_forward_pre_hook(lambda mod, a: toks.append(a[0].detach().cpu().numpy()))
    import comet_pose_estimation_b200.blocks as blk

    orig = blk._Pyramid.__init__

    def spy(self, fmaps, num_levels):
        orig(self, fmaps, num_levels)
        layouts.append(self.cl_input)

    blk._Pyramid.__init__ = spy
    orig_up = blk._PyramidUp2.__init__

    def spy_up(self, up, num_levels):
        orig_up(self, up, num_levels)
        layouts.append("up2")

    blk._PyramidUp2.__init__ = spy_up
    rt.DEFER_UPSAMPLE = defer
    # the golden comes from the reference on CPU (strict fp32); cuDNN would otherwise run the encoder's convolutions
    # in TF32 (PyTorch's default), which is an encoder-precision choice outside the path under test
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            refined, score = refine_track(torch.from_numpy(images).cuda(), fnet, ftr, torch.from_numpy(coarse).cuda(),
                                          compute_score=True)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
        blk._Pyramid.__init__ = orig
        blk._PyramidUp2.__init__ = orig_up
        rt.DEFER_UPSAMPLE = True
        h.remove()
    # defer: the half-resolution source reached the kernels; else: the encoder output as a channels-last view, zero-copy
    assert layouts == (["up2"] if defer else [True])
    assert rel_to_max(toks[0], g[NAME + "/tok0"]) < 1e-4
    for i in range(1, 6):
        assert rel_to_max(toks[i], g[f"{NAME}/tok{i}"]) < TOK_BAR
    assert rel_to_max(refined.cpu().numpy(), g[NAME + "/refined"]) < 1e-4   # "refined tracks" check-point, fp32 bar
    assert np.array_equal(refined[:, 0].cpu().numpy(), coarse[:, 0])
    assert rel_to_max(score.cpu().numpy(), g[NAME + "/score"]) < 1e-4       # "pred_score" check-point


def _per_operator(rt, fnet, x, **kw):
    """The encoder on the per-operator path (cuDNN float32 convolutions + the library's norm / resize kernels)."""
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    rt.USE_FUSED_ENCODER = False
    try:
        with torch.no_grad():
            return fnet(x, **kw)
    finally:
        rt.USE_FUSED_ENCODER = True
        torch.backends.cudnn.allow_tf32 = tf32


@pytest.mark.gpu
def test_fused_shallow_encoder_matches_reference(golden):
    """csrc/shallow_encoder.cu (all layers of one patch in shared memory, float32 FMA) against the encoder output of the
    executed reference, for NCHW and channels-last patch tensors."""
    import importlib

    rt = importlib.import_module("comet_pose_estimation_b200.refine_track")
    g = golden("refine")
    fnet, _ = modules(g)
    fnet = fnet.cuda()
    x = torch.from_numpy(g[NAME + "/enc_in"]).cuda()
    from comet_pose_estimation_b200 import _lib

    n0 = _lib.lib.comet_launch_count()
    with torch.no_grad():
        y = fnet(x)
        y_cl = fnet(x.contiguous(memory_format=torch.channels_last))
        half, size = fnet(x, defer_upsample=True)
    assert _lib.lib.comet_launch_count() - n0 >= 3            # the library's kernels ran, not torch.nn
    assert size == (31, 31) and tuple(half.shape) == (4, 32, 16, 16)
    assert half.permute(0, 2, 3, 1).is_contiguous()           # channel-last memory: what Upsampled2x consumes
    assert rel_to_max(y.cpu().numpy(), g[NAME + "/enc_out"]) < 1e-5
    assert torch.equal(y, y_cl)
    ref_half, _ = _per_operator(rt, fnet, x, defer_upsample=True)
    assert rel_to_max(half.cpu().numpy(), ref_half.cpu().numpy()) < 1e-5


@pytest.mark.gpu
def test_fused_shallow_encoder_from_images_odd_count_and_repacking():
    """Gather fused into the encoder: same values as extract_patches + forward, bit for bit; an odd patch count
    (the last CTA carries one patch); parameters modified in place are re-packed."""
    import importlib

    rt = importlib.import_module("comet_pose_estimation_b200.refine_track")
    torch.manual_seed(11)
    fnet = rt.ShallowEncoder(input_dim=3).eval().cuda()
    with torch.no_grad():
        for prm in fnet.parameters():
            if prm.dim() == 1:
                prm.uniform_(-0.5, 0.5)                       # biases are zero-initialised by default: exercise them
    gen = torch.Generator(device="cuda").manual_seed(5)
    B, S, N, H, W = 1, 3, 7, 80, 64                            # 21 patches
    images = torch.rand(B, S, 3, H, W, device="cuda", generator=gen)
    tl = torch.stack([torch.randint(0, W - 31 + 1, (B, S, N), device="cuda", generator=gen),
                      torch.randint(0, H - 31 + 1, (B, S, N), device="cuda", generator=gen)], -1).int()
    tl[0, 0, 0] = 0
    tl[0, 2, 6, 0], tl[0, 2, 6, 1] = W - 31, H - 31
    with torch.no_grad():
        patches = rt.extract_patches(images, tl, 31)
        a, _ = fnet(patches, defer_upsample=True)
        b, _ = fnet.encode_patches_of(images, tl, defer_upsample=True)
    assert tuple(a.shape) == (21, 32, 16, 16)
    assert torch.equal(a, b)
    ref, _ = _per_operator(rt, fnet, patches, defer_upsample=True)
    assert rel_to_max(a.cpu().numpy(), ref.cpu().numpy()) < 1e-5
    with torch.no_grad():
        fnet.layer2.conv2.weight.mul_(-1.5)
        c, _ = fnet(patches, defer_upsample=True)
    ref2, _ = _per_operator(rt, fnet, patches, defer_upsample=True)
    assert not torch.equal(a, c)
    assert rel_to_max(c.cpu().numpy(), ref2.cpu().numpy()) < 1e-5
