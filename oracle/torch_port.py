"""Multi-threaded CPU port of the hot path -- TEST / BASELINE INFRASTRUCTURE ONLY.

Same status as oracle/comet_oracle.py (never imported by the product).  Where comet_oracle.py restates the
arithmetic from scratch in numpy, this file expresses the same steps with the ATen library calls the reference
itself makes on CPU (``torch.matmul``, ``F.grid_sample``, ``F.avg_pool2d``, ``torch.sin/cos``) so that it runs as
fast as the reference does on the host cores.  It is what ``bench.py`` times for ``cpu_baseline`` (kind "port")
and for ``--impl reference`` (the reference itself is Python source that cannot travel to the GPU box).
Pinned to the reference goldens in tests/test_oracle_golden.py::test_torch_port_*.

Reference: comet/models/track_modules/blocks.py:351-429, comet/models/utils.py:65-101, :724-974,
comet/models/track_modules/base_track_predictor.py:153-224.
"""
from __future__ import annotations

import math
from typing import List

import torch
import torch.nn.functional as F


def pyramid(fmaps: torch.Tensor, num_levels: int) -> List[torch.Tensor]:
    B, S, C, H, W = fmaps.shape
    levels = [fmaps]
    cur = fmaps.reshape(B * S, C, H, W)
    for _ in range(num_levels - 1):
        cur = F.avg_pool2d(cur, 2, stride=2)
        levels.append(cur.reshape(B, S, C, *cur.shape[-2:]))
    return levels


def volumes(levels: List[torch.Tensor], targets: torch.Tensor) -> List[torch.Tensor]:
    B, S, N, C = targets.shape
    scale = math.sqrt(C)
    out = []
    for f in levels:
        h, w = f.shape[-2:]
        v = torch.matmul(targets, f.reshape(B, S, C, h * w)) / scale
        out.append(v.reshape(B, S, N, h, w))
    return out


def _px_to_grid(xy: torch.Tensor, h: int, w: int) -> torch.Tensor:
    sx = 2.0 / max(w - 1, 1)
    sy = 2.0 / max(h - 1, 1)
    return xy * xy.new_tensor([sx, sy]) - 1.0


def lookup(vols: List[torch.Tensor], coords: torch.Tensor, radius: int, padding_mode: str = "zeros") -> torch.Tensor:
    B, S, N, _ = coords.shape
    d = torch.linspace(-radius, radius, 2 * radius + 1)
    gi, gj = torch.meshgrid(d, d, indexing="ij")
    delta = torch.stack([gi, gj], dim=-1)[None]  # [...,0] added to x, slow index
    outs = []
    for l, v in enumerate(vols):
        h, w = v.shape[-2:]
        pts = coords.reshape(B * S * N, 1, 1, 2) / (2 ** l) + delta
        s = F.grid_sample(v.reshape(B * S * N, 1, h, w).float(), _px_to_grid(pts, h, w), align_corners=True,
                          padding_mode=padding_mode)
        outs.append(s.reshape(B, S, N, -1))
    return torch.cat(outs, dim=-1)


def point_sample(img: torch.Tensor, pts: torch.Tensor) -> torch.Tensor:
    """sample_features4d: (B,C,H,W) @ (B,R,2) -> (B,R,C), border."""
    h, w = img.shape[-2:]
    s = F.grid_sample(img, _px_to_grid(pts[:, :, None, :], h, w), align_corners=True, padding_mode="border")
    return s[..., 0].permute(0, 2, 1)


def flow_embedding(flow: torch.Tensor, C: int) -> torch.Tensor:
    div = torch.arange(0, C, 2, dtype=torch.float32) * (1000.0 / C)
    ax = flow[..., 0:1] * div
    ay = flow[..., 1:2] * div
    ex = torch.stack([ax.sin(), ax.cos()], dim=-1).flatten(-2)
    ey = torch.stack([ay.sin(), ay.cos()], dim=-1).flatten(-2)
    return torch.cat([ex, ey], dim=-1)


def sincos_table(D: int, h: int, w: int) -> torch.Tensor:
    q = D // 4
    om = 1.0 / 10000 ** (torch.arange(q, dtype=torch.double) / q)
    ax = torch.arange(w, dtype=torch.double)[:, None] * om  # (w,q)
    ay = torch.arange(h, dtype=torch.double)[:, None] * om  # (h,q)
    tx = torch.cat([ax.sin(), ax.cos()], dim=1).T[:, None, :].expand(2 * q, h, w)
    ty = torch.cat([ay.sin(), ay.cos()], dim=1).T[:, :, None].expand(2 * q, h, w)
    return torch.cat([tx, ty], dim=0)[None].float()


def tokens(fcorrs: torch.Tensor, coords: torch.Tensor, feats: torch.Tensor, hw, tdim: int) -> torch.Tensor:
    B, S, N, LW = fcorrs.shape
    latent = feats.shape[-1]
    flow = (coords - coords[:, 0:1]).permute(0, 2, 1, 3)
    parts = [flow_embedding(flow, latent // 2), flow, fcorrs.permute(0, 2, 1, 3), feats.permute(0, 2, 1, 3)]
    x = torch.cat(parts, dim=-1)
    if x.shape[-1] < tdim:
        x = F.pad(x, (0, tdim - x.shape[-1]))
    pos = point_sample(sincos_table(tdim, *hw).expand(B, -1, -1, -1), coords[:, 0])
    return x + pos[:, :, None, :]


def hot_path_iteration(levels, coords, feats, radius, hw, tdim):
    """One refinement iteration of the hot path (corr + lookup + token assembly) on the CPU."""
    fc = lookup(volumes(levels, feats), coords, radius)
    return tokens(fc, coords, feats, hw, tdim)
