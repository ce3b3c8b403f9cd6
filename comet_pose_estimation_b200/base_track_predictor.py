"""Drop-in ``BaseTrackerPredictor`` (comet/models/track_modules/base_track_predictor.py:15-284).

Same constructor, ``forward`` signature, return tuple and state-dict keys (``updateformer.*``, ``norm.*``,
``ffeat_updater.0.*``, ``vis_predictor.0.*``, ``conf_predictor.0.*``).  Per refinement iteration the reference
issues ~25 ATen kernels plus a host-built, re-uploaded 10.9 MB sin/cos table; here the hot path is:

    once per call    pyramid (+ hi/lo split for the tensor path), sampled position embedding, query features
    per iteration    ONE fused kernel -> tokens (B,N,S,D), then the update transformer and the state update

The loop-state arithmetic (coords += delta, track_feats += ffeat_updater(norm(delta_feat)), frame 0 pinned,
rescale to pixels, base_track_predictor.py:229-262) is kept operation for operation.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .blocks import CorrBlock, EfficientCorrBlock, Upsampled2x
from .track_tokens import TrackTokenizer, transformer_dim
from .update_former import EfficientUpdateFormer
from .utils import sample_features4d


class BaseTrackerPredictor(nn.Module):
    def __init__(self, stride=4, corr_levels=5, corr_radius=4, latent_dim=128, hidden_size=384, use_spaceatt=True,
                 depth=6, fine=False, cfg=None, updateformer_cls=EfficientUpdateFormer):
        super().__init__()
        self.cfg = cfg
        self.stride = stride
        self.latent_dim = latent_dim
        self.corr_levels = corr_levels
        self.corr_radius = corr_radius
        self.hidden_size = hidden_size
        self.fine = fine
        self.flows_emb_dim = latent_dim // 2
        self.transformer_dim = transformer_dim(corr_levels, corr_radius, latent_dim, fine)
        self.efficient_corr = cfg.MODEL.TRACK.efficient_corr  # config key MODEL.TRACK.efficient_corr (abl_ours.yaml:401)
        space_depth = depth if use_spaceatt else 0
        self.updateformer = updateformer_cls(
            space_depth=space_depth, time_depth=depth, input_dim=self.transformer_dim, hidden_size=hidden_size,
            output_dim=latent_dim + 2, mlp_ratio=4.0, add_space_attn=use_spaceatt)
        self.norm = nn.GroupNorm(1, latent_dim)
        self.ffeat_updater = nn.Sequential(nn.Linear(latent_dim, latent_dim), nn.GELU())
        if self.cfg.track_conf:  # config key track_conf (abl_ours.yaml:380)
            self.conf_predictor = nn.Sequential(nn.Linear(latent_dim, 1))
        if not self.fine:
            self.vis_predictor = nn.Sequential(nn.Linear(latent_dim, 1))

    def forward(self, query_points, fmaps=None, iters=4, return_feat=False, down_ratio=1, is_train=False,
                track_feats=None, TRACKorPOSE=True, ind=0):
        if TRACKorPOSE:
            B, S, N, D = query_points.shape
        else:
            B, N, D = query_points.shape
        B, S, C, HH, WW = fmaps.shape
        assert D == 2
        if down_ratio > 1:  # guards BOTH divisions (SURVEY A.6 i)
            query_points = query_points / float(down_ratio)
            query_points = query_points / float(self.stride)
        if TRACKorPOSE:
            coords = query_points.clone()
        else:
            coords = query_points.clone().reshape(B, 1, N, 2).repeat(1, S, 1, 1)

        if isinstance(fmaps, Upsampled2x):
            # bilinear sample (border) of the 2x-1 up-sampling at p == bilinear sample of its source at p / 2
            query_track_feat = sample_features4d(fmaps.src[:, 0], coords[:, 0] * 0.5)
        else:
            query_track_feat = sample_features4d(fmaps[:, 0], coords[:, 0])
        track_feats = query_track_feat.unsqueeze(1).repeat(1, S, 1, 1)
        coords_backup = coords.clone()

        if self.efficient_corr:
            fcorr_fn = EfficientCorrBlock(fmaps, num_levels=self.corr_levels, radius=self.corr_radius)
        else:
            fcorr_fn = CorrBlock(fmaps, num_levels=self.corr_levels, radius=self.corr_radius)
        # sampled_pos_emb depends on coords[:, 0] only, which the loop pins -> computed once, not per iteration
        tokenizer = TrackTokenizer(fcorr_fn, coords[:, 0], self.transformer_dim)

        coord_preds = []
        for _ in range(iters):
            coords = coords.detach()
            x = tokenizer.tokens(coords, track_feats)  # (B, N, S, transformer_dim), one fused kernel

            delta = self.updateformer(x).reshape(B * N, S, self.latent_dim + 2)
            delta_coords_ = delta[:, :, :2]
            delta_feats_ = delta[:, :, 2:].reshape(B * N * S, self.latent_dim)

            track_feats_ = track_feats.permute(0, 2, 1, 3).reshape(B * N * S, self.latent_dim)
            track_feats_ = self.ffeat_updater(self.norm(delta_feats_)) + track_feats_
            track_feats = track_feats_.reshape(B, N, S, self.latent_dim).permute(0, 2, 1, 3)

            coords = coords + delta_coords_.reshape(B, N, S, 2).permute(0, 2, 1, 3)
            coords[:, 0] = coords_backup[:, 0]
            if down_ratio > 1:
                coord_preds.append(coords * self.stride * down_ratio)
            else:
                coord_preds.append(coords * self.stride)

        if not self.fine:
            vis_e = torch.sigmoid(self.vis_predictor(track_feats.reshape(B * S * N, self.latent_dim)).reshape(B, S, N))
        else:
            vis_e = None
        conf_e = None
        if self.cfg.track_conf:
            conf_e = torch.sigmoid(self.conf_predictor(track_feats.reshape(B * S * N, self.latent_dim)).reshape(B, S, N))
        if return_feat:
            return coord_preds, vis_e, track_feats, query_track_feat, conf_e
        return coord_preds, vis_e, conf_e
