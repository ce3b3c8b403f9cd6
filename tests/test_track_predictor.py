"""``BasicEncoder`` / ``TrackerPredictor`` mirror (the step before the hot path: comet/models/track_modules/blocks.py:27-111,
comet/models/track_predictor.py:16-151) against tensors produced by executing the reference (tests/golden/encoders.npz;
weights rebuilt by name from ``cases.seeded_state_dict``)."""
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

import cases
from conftest import rel_to_max


def encoder():
    from comet_pose_estimation_b200.track_predictor import BasicEncoder

    enc = BasicEncoder(input_dim=3, output_dim=128, stride=4).eval()
    sd = cases.seeded_state_dict({k: tuple(v.shape) for k, v in enc.state_dict().items()}, 91)
    enc.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    return enc


def test_basic_encoder_state_dict_keys_and_cpu_forward(golden):
    g = golden("encoders")
    enc = encoder()
    assert sorted(enc.state_dict().keys()) == sorted(g["basic/keys"].tolist())      # reference parameter names
    img = torch.from_numpy(g["basic/images"])
    with torch.no_grad():
        y = enc(img.reshape(2, 3, 96, 80))
    assert rel_to_max(y.numpy(), g["basic/encoder_only"]) < 1e-5


def test_process_images_to_fmaps_cpu(golden):
    from comet_pose_estimation_b200.track_predictor import TrackerPredictor

    g = golden("encoders")
    tp = TrackerPredictor(coarse_fnet=encoder()).eval()
    with torch.no_grad():
        fm = tp.process_images_to_fmaps(torch.from_numpy(g["basic/images"]))
    assert fm.shape == g["basic/fmaps"].shape == (1, 2, 128, 12, 10)
    assert rel_to_max(fm.numpy(), g["basic/fmaps"]) < 1e-5
    with pytest.raises(AssertionError):
        tp.process_images_to_fmaps(torch.zeros(2, 2, 3, 16, 16))      # inference is one scene at a time (track_predictor.py:130)
    # attribute names of the reference container (checkpoint prefixes track_predictor.{coarse,fine}_{fnet,predictor}.*)
    keys = tp.state_dict().keys()
    for prefix in ("coarse_fnet.conv1.weight", "coarse_predictor.updateformer.input_transform.weight",
                   "fine_fnet.layer1.downsample.0.weight", "fine_predictor.ffeat_updater.0.weight"):
        assert prefix in keys


@pytest.mark.gpu
def test_basic_encoder_on_library_kernels(golden):
    g = golden("encoders")
    from comet_pose_estimation_b200.track_predictor import TrackerPredictor

    tp = TrackerPredictor(coarse_fnet=encoder()).eval().cuda()
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            fm = tp.process_images_to_fmaps(torch.from_numpy(g["basic/images"]).cuda())
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    assert rel_to_max(fm.cpu().numpy(), g["basic/fmaps"]) < 1e-4


@pytest.mark.gpu
def test_track_equals_manual_composition():
    """``TrackerPredictor.track`` == process_images_to_fmaps -> coarse predictor -> refine_track -> inverted_score
    (the tracker part of COMET.forward_all, E2Epose2.py:176-239)."""
    import comet_pose_estimation_b200 as cb

    cfg = NS(track_conf=False, MODEL=NS(TRACK=NS(efficient_corr=False)))
    torch.manual_seed(3)
    tp = cb.TrackerPredictor(
        coarse_predictor=cb.BaseTrackerPredictor(cfg=cfg, hidden_size=64, depth=2),
        fine_predictor=cb.BaseTrackerPredictor(cfg=cfg, stride=1, depth=1, corr_levels=3, corr_radius=3, latent_dim=32,
                                               hidden_size=64, fine=True, use_spaceatt=False), cfg=cfg).eval().cuda()
    tp.fine_fnet.to(memory_format=torch.channels_last)
    images = torch.rand(1, 3, 3, 512, 512, device="cuda")
    q = torch.rand(1, 20, 2, device="cuda") * 400 + 56
    out = tp.track(images, q, coarse_iters=2)
    assert out["refine_pred_track"].shape == (1, 3, 20, 2) and out["pred_score"].shape == (1, 3, 20)
    assert torch.equal(out["refine_pred_track"][:, 0], q)
    with torch.no_grad():
        fm = tp.process_images_to_fmaps(images)
        assert fm.shape == (1, 3, 128, 64, 64)
        coarse = tp.coarse_predictor(query_points=q, fmaps=fm, iters=2, down_ratio=2, TRACKorPOSE=False)[0][-1]
        refined, score = cb.refine_track(images, tp.fine_fnet, tp.fine_predictor, coarse, compute_score=True)
    assert torch.equal(out["coarse_pred_track"], coarse)
    assert torch.equal(out["refine_pred_track"], refined)
    assert torch.equal(out["pred_score"], cb.inverted_score(score))
    assert float(out["pred_score"].max()) <= 1.0 + 1e-6
