"""Drop-in ``TrackerPredictor`` container and ``BasicEncoder`` (comet/models/track_predictor.py:16-151,
comet/models/track_modules/blocks.py:27-111) -- the step right before the hot path -- and
the tracker part of ``COMET.forward_all`` (comet/models/E2Epose2.py:176-239) as one call.

What is B200-specific: its four bilinear resizes to 1/stride resolution
(blocks.py:99-102, ``_bilinear_intepolate`` :198-201) and its instance norms run in this library's kernels
(``comet_upsample_bilinear_ac_f32``, ``comet_instance_norm_f32``: ATen's up-sampling kernel walks batch x channels inside
every thread); the convolutions stay cuDNN.  Parameter names are the reference's (``conv1``, ``layer{1..4}.{0,1}.conv{1,2}``,
``layer*.0.downsample.0``, ``conv2``, ``conv3``), so ``track_predictor.coarse_fnet.*`` checkpoints load unchanged.

The reference builds the four sub-modules with ``hydra.utils.instantiate`` from the ``COARSE`` / ``FINE`` config nodes
(track_predictor.py:48-62); here they are passed in (or built with the shipped defaults of abl_ours.yaml:399-428), so no
hydra dependency is needed.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F

from .base_track_predictor import BaseTrackerPredictor
from .refine_track import ShallowEncoder, _inorm, _ResidualBlock, _resize, inverted_score, refine_track


class BasicEncoder(nn.Module):
    """blocks.py:27-111 (``norm_fn="instance"``, the only value the reference uses)."""

    def __init__(self, input_dim=3, output_dim=128, stride=4, use_trans=False, cfg=None):
        super().__init__()
        self.stride = stride
        self.norm_fn = "instance"
        self.in_planes = output_dim // 2
        self.norm1 = nn.InstanceNorm2d(self.in_planes)
        self.norm2 = nn.InstanceNorm2d(output_dim * 2)
        self.conv1 = nn.Conv2d(input_dim, self.in_planes, kernel_size=7, stride=2, padding=3)
        self.layer1 = self._make_layer(output_dim // 2, stride=1)
        self.layer2 = self._make_layer(output_dim // 4 * 3, stride=2)
        self.layer3 = self._make_layer(output_dim, stride=2)
        self.layer4 = self._make_layer(output_dim, stride=2)
        self.conv2 = nn.Conv2d(output_dim * 3 + output_dim // 4, output_dim * 2, kernel_size=3, padding=1)
        self.conv3 = nn.Conv2d(output_dim * 2, output_dim, kernel_size=1)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")

    def _make_layer(self, dim, stride=1):
        layers = (_ResidualBlock(self.in_planes, dim, stride=stride), _ResidualBlock(dim, dim, stride=1))
        self.in_planes = dim
        return nn.Sequential(*layers)

    def forward(self, x):
        _, _, H, W = x.shape
        size = (H // self.stride, W // self.stride)
        x = _inorm(self.norm1, self.conv1(x), True)
        a = self.layer1(x)
        b = self.layer2(a)
        c = self.layer3(b)
        d = self.layer4(c)
        x = torch.cat([_resize(a, size), _resize(b, size), _resize(c, size), _resize(d, size)], dim=1)
        x = _inorm(self.norm2, self.conv2(x), True)
        return self.conv3(x)


def _default_cfg():
    return SimpleNamespace(track_conf=False, MODEL=SimpleNamespace(TRACK=SimpleNamespace(efficient_corr=False)))


class TrackerPredictor(nn.Module):
    """track_predictor.py:16-151.  Attribute names (``coarse_fnet``, ``coarse_predictor``, ``fine_fnet``,
    ``fine_predictor``, ``coarse_down_ratio``) are the reference's; sub-modules not passed in are built with the shipped
    configuration (abl_ours.yaml:399-428)."""

    def __init__(self, coarse_fnet=None, coarse_predictor=None, fine_fnet=None, fine_predictor=None, coarse_stride=4,
                 coarse_down_ratio=2, cfg=None):
        super().__init__()
        self.cfg = cfg if cfg is not None else _default_cfg()
        self.coarse_down_ratio = coarse_down_ratio
        self.coarse_fnet = coarse_fnet if coarse_fnet is not None else BasicEncoder(stride=coarse_stride, cfg=self.cfg)
        self.coarse_predictor = coarse_predictor if coarse_predictor is not None else BaseTrackerPredictor(
            stride=coarse_stride, cfg=self.cfg)
        self.fine_fnet = fine_fnet if fine_fnet is not None else ShallowEncoder(input_dim=3, stride=1, cfg=self.cfg)
        self.fine_predictor = fine_predictor if fine_predictor is not None else BaseTrackerPredictor(
            stride=1, depth=4, corr_levels=3, corr_radius=3, latent_dim=32, hidden_size=256, fine=True,
            use_spaceatt=False, cfg=self.cfg)

    def process_images_to_fmaps(self, images, training=False):
        """track_predictor.py:117-151: (B,S,3,H,W) -> (B,S,128,H/(stride*down_ratio),W/(stride*down_ratio))."""
        B, S, C, H, W = images.shape
        if not training:
            assert B == 1, "now we only support processing one scene during inference"
        x = images.reshape(B * S, C, H, W)
        if self.coarse_down_ratio > 1:
            # F.interpolate(scale_factor=1/down_ratio, bilinear, align_corners=True): output size floor(H / down_ratio)
            x = _resize(x, (int(H * (1.0 / self.coarse_down_ratio)), int(W * (1.0 / self.coarse_down_ratio))))
        # NCHW on purpose: the channels-last instance-norm kernel of this library is built for the patch encoder's many tiny
        # planes (one thread per (sample, channel)); at 16 x 64 planes of 256 x 256 it would take 20 ms (measured), the
        # NCHW kernel (one warp per plane) 2.9 ms for the whole encoder (torch ops: 5.5 ms)
        fmaps = self.coarse_fnet(x.contiguous())
        return fmaps.reshape(B, S, -1, fmaps.shape[-2], fmaps.shape[-1])

    @torch.no_grad()
    def track(self, images, query_points, coarse_iters=4, fine_tracking=True):
        """The tracker part of ``COMET.forward_all`` (E2Epose2.py:176-239): images (B,S,3,H,W) in the model's
        normalisation, ``query_points`` (B,N,2) pixels of frame 0 -> dict with ``coarse_pred_track`` /
        ``refine_pred_track`` (B,S,N,2), ``pred_score`` (the inverted, normalised track confidence the camera predictor
        consumes), ``vis`` and the per-iteration coarse list."""
        fmaps = self.process_images_to_fmaps(images)
        coarse_list, vis, track_feats, query_feat, conf = self.coarse_predictor(
            query_points=query_points, fmaps=fmaps, iters=coarse_iters, down_ratio=self.coarse_down_ratio, is_train=False,
            return_feat=True, TRACKorPOSE=False)
        coarse = coarse_list[-1]
        out = {"coarse_pred_track_list": coarse_list, "coarse_pred_track": coarse, "vis": vis, "conf": conf}
        if fine_tracking:
            refined, score = refine_track(images, self.fine_fnet, self.fine_predictor, coarse, compute_score=True)
            out["refine_pred_track"] = refined
            out["track_score"] = score
            out["pred_score"] = inverted_score(score)
        else:
            out["refine_pred_track"] = coarse
            out["pred_score"] = torch.ones_like(vis)
        return out
