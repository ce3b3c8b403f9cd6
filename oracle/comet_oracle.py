"""CPU oracle for the COMET tracking hot path -- TEST INFRASTRUCTURE ONLY.

This file is a from-scratch numpy restatement of the arithmetic that the
reference (wulibingbinglin/COMET-Pose-Estimation, mounted at /root/reference
while the fixtures were generated) performs on its point-tracking hot path.
It exists to *check* the CUDA kernels; it is never the thing measured or
shipped.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.
The product package (``comet_pose_estimation_b200``) must never import it and
has no CPU fallback.

Parity status: PINNED.  Every public function below is compared against
outputs of the unmodified reference modules (fixtures under ``tests/golden/``
produced by ``tests/golden/make_golden.py``, which imports the reference from
/root/reference) in ``tests/test_oracle_golden.py``.

The arithmetic lives partly in PyTorch ATen (``matmul``, ``grid_sample``,
``avg_pool2d``, ``sin``/``cos``; the reference pins torch==2.1.0 in
requirements.txt:30, fixtures were generated under torch 2.11.0).  Their
published semantics are restated here: bilinear ``grid_sample`` with
``align_corners`` and ``zeros``/``border`` padding, floor-mode 2x2 average
pooling, float32 accumulation.

Conventions: all arrays are numpy; float32 unless stated; coordinates are
``(x, y)`` in level-0 cell units.  Reference citations are
``file:line`` into /root/reference.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------
# grid_sample semantics (ATen, restated)
# --------------------------------------------------------------------------
def _unnormalize(g: np.ndarray, size: int, align_corners: bool) -> np.ndarray:
    """ATen grid_sampler_unnormalize: [-1,1] -> pixel units (float32)."""
    g = g.astype(F32)
    if align_corners:
        return ((g + F32(1)) / F32(2)) * F32(size - 1)
    return ((g + F32(1)) * F32(size) - F32(1)) / F32(2)


def _linear_taps(p: np.ndarray, size: int, border: bool):
    """Per-axis linear interpolation taps.

    Returns (i0, i1, w0, w1, ok0, ok1): integer taps, their float32 weights and
    in-bounds masks.  ``border`` clamps the *coordinate* first (ATen
    clip_coordinates); ``zeros`` leaves the coordinate alone and masks taps that
    fall outside [0, size-1] (weights still come from the unclamped value).
    """
    p = p.astype(F32)
    if border:
        p = np.minimum(F32(size - 1), np.maximum(p, F32(0)))
    f = np.floor(p)
    i0 = f.astype(np.int64)
    i1 = i0 + 1
    w1 = (p - f).astype(F32)
    w0 = (F32(1) - w1).astype(F32)
    ok0 = (i0 >= 0) & (i0 <= size - 1)
    ok1 = (i1 >= 0) & (i1 <= size - 1)
    # clamp indices for safe gathering; masked taps contribute exactly zero
    i0c = np.clip(i0, 0, size - 1)
    i1c = np.clip(i1, 0, size - 1)
    return i0c, i1c, w0, w1, ok0, ok1


def bilinear_sampler(
    inp: np.ndarray,
    coords: np.ndarray,
    align_corners: bool = True,
    padding_mode: str = "border",
) -> np.ndarray:
    """Restates ``bilinear_sampler`` comet/models/utils.py:874-939.

    ``inp`` (B,C,H,W) with ``coords`` (B,Ho,Wo,2) = (x,y), or ``inp``
    (B,C,T,H,W) with ``coords`` (B,D,Ho,Wo,3) = (t,x,y).  Coordinates are in
    pixel units; they are scaled to [-1,1] (utils.py:925-935) and handed to
    ``grid_sample`` which scales them back -- the round trip is kept so that
    float32 rounding matches.
    """
    assert padding_mode in ("zeros", "border")
    border = padding_mode == "border"
    inp = np.asarray(inp, dtype=F32)
    coords = np.asarray(coords, dtype=F32)
    sizes = inp.shape[2:]
    assert len(sizes) in (2, 3)
    if len(sizes) == 3:
        coords = coords[..., [1, 2, 0]]  # (t,x,y) -> (x,y,t)  utils.py:921-923
    rsz = list(reversed(sizes))  # (W,H[,T]) order of the coordinate tuple
    if align_corners:
        scale = np.array([2.0 / max(s - 1, 1) for s in rsz], dtype=F32)
    else:
        scale = np.array([2.0 / s for s in rsz], dtype=F32)
    g = coords * scale - F32(1)

    B, C = inp.shape[:2]
    taps = []
    for ax, s in enumerate(rsz):
        p = _unnormalize(g[..., ax], s, align_corners)
        taps.append(_linear_taps(p, s, border))

    bidx = np.arange(B).reshape((B,) + (1,) * (coords.ndim - 2))
    out_shape = (B, C) + coords.shape[1:-1]
    out = np.zeros(out_shape, dtype=F32)
    if len(sizes) == 2:
        (x0, x1, wx0, wx1, okx0, okx1), (y0, y1, wy0, wy1, oky0, oky1) = taps
        # ATen accumulation order: nw, ne, sw, se
        for yi, wy, oky in ((y0, wy0, oky0), (y1, wy1, oky1)):
            for xi, wx, okx in ((x0, wx0, okx0), (x1, wx1, okx1)):
                v = inp[bidx, :, yi, xi]  # (B,Ho,Wo,C)
                w = (wx * wy) * (okx & oky).astype(F32)
                out += np.moveaxis(v * w[..., None], -1, 1)
    else:
        (x0, x1, wx0, wx1, okx0, okx1), (y0, y1, wy0, wy1, oky0, oky1), (
            t0, t1, wt0, wt1, okt0, okt1) = taps
        for ti, wt, okt in ((t0, wt0, okt0), (t1, wt1, okt1)):
            for yi, wy, oky in ((y0, wy0, oky0), (y1, wy1, oky1)):
                for xi, wx, okx in ((x0, wx0, okx0), (x1, wx1, okx1)):
                    v = inp[bidx, :, ti, yi, xi]
                    w = (wx * wy * wt) * (okx & oky & okt).astype(F32)
                    out += np.moveaxis(v * w[..., None], -1, 1)
    return out


def sample_features4d(inp: np.ndarray, coords: np.ndarray) -> np.ndarray:
    """Restates ``sample_features4d`` comet/models/utils.py:942-974:
    (B,C,H,W) sampled at (B,R,2) -> (B,R,C), border padding, align_corners."""
    feats = bilinear_sampler(inp, np.asarray(coords)[:, :, None, :])  # B C R 1
    return np.ascontiguousarray(feats[..., 0].transpose(0, 2, 1))


# --------------------------------------------------------------------------
# feature pyramid and correlation
# --------------------------------------------------------------------------
def avg_pool2(x: np.ndarray) -> np.ndarray:
    """2x2 / stride 2 average pooling, floor output size, no padding
    (``F.avg_pool2d(x, 2, stride=2)``, blocks.py:371)."""
    H, W = x.shape[-2:]
    h, w = H // 2, W // 2
    x = x[..., : 2 * h, : 2 * w].astype(F32)
    s = (x[..., 0::2, 0::2] + x[..., 0::2, 1::2]) + (x[..., 1::2, 0::2] + x[..., 1::2, 1::2])
    return (s * F32(0.25)).astype(F32)


def build_pyramid(fmaps: np.ndarray, num_levels: int) -> List[np.ndarray]:
    """Restates ``CorrBlock.__init__`` blocks.py:352-374: level 0 is the
    input (B,S,C,H,W); level l is avg_pool2 of level l-1."""
    pyr = [np.asarray(fmaps, dtype=F32)]
    for _ in range(num_levels - 1):
        pyr.append(avg_pool2(pyr[-1]))
    return pyr


def corr_volumes(targets: np.ndarray, pyramid: Sequence[np.ndarray]) -> List[np.ndarray]:
    """Restates ``CorrBlock.corr`` blocks.py:409-429: per level
    ``V_l[b,s,n,y,x] = (sum_c T[b,s,n,c] * F_l[b,s,c,y,x]) / sqrt(C)``;
    the division happens after the matmul, by a float32 scalar."""
    targets = np.asarray(targets, dtype=F32)
    B, S, N, C = targets.shape
    out = []
    inv = F32(math.sqrt(C))  # torch.sqrt(torch.tensor(C).float())
    for f in pyramid:
        assert f.shape[2] == C and f.shape[1] == S
        H, W = f.shape[-2:]
        v = np.matmul(targets, f.reshape(B, S, C, H * W)).astype(F32)
        out.append((v / inv).reshape(B, S, N, H, W).astype(F32))
    return out


def _window_offsets(radius: int) -> np.ndarray:
    """The reference builds ``delta = stack(meshgrid(dy, dx, 'ij'), -1)``
    (blocks.py:385-391), i.e. delta[i,j] = (d_i, d_j): the value added to the
    *x* coordinate varies with the slow index i, the value added to *y* with
    the fast index j (the window is transposed)."""
    d = np.linspace(-radius, radius, 2 * radius + 1, dtype=F32)
    gi, gj = np.meshgrid(d, d, indexing="ij")
    return np.stack([gi, gj], axis=-1)  # (Wr,Wr,2): [...,0]->x  [...,1]->y


def lookup_volumes(
    volumes: Sequence[np.ndarray], coords: np.ndarray, radius: int, padding_mode: str = "zeros"
) -> np.ndarray:
    """Restates ``CorrBlock.sample`` blocks.py:376-407."""
    coords = np.asarray(coords, dtype=F32)
    B, S, N, D = coords.shape
    assert D == 2
    delta = _window_offsets(radius)[None]
    outs = []
    for lvl, v in enumerate(volumes):
        H, W = v.shape[-2:]
        cen = coords.reshape(B * S * N, 1, 1, 2) / F32(2 ** lvl)
        smp = bilinear_sampler(v.reshape(B * S * N, 1, H, W), cen + delta, padding_mode=padding_mode)
        outs.append(smp.reshape(B, S, N, -1))
    return np.ascontiguousarray(np.concatenate(outs, axis=-1))


def corr_lookup(
    fmaps: np.ndarray,
    targets: np.ndarray,
    coords: np.ndarray,
    num_levels: int,
    radius: int,
    padding_mode: str = "zeros",
) -> np.ndarray:
    """CorrBlock(fmaps).corr(targets); .sample(coords) in one call
    (blocks.py:351-429) -> (B,S,N,L*(2r+1)^2)."""
    pyr = build_pyramid(fmaps, num_levels)
    return lookup_volumes(corr_volumes(targets, pyr), coords, radius, padding_mode)


def efficient_corr_lookup(
    fmaps: np.ndarray, targets: np.ndarray, coords: np.ndarray, num_levels: int, radius: int
) -> np.ndarray:
    """Restates ``EfficientCorrBlock.sample`` blocks.py:446-484: sample the
    C-channel features at the window taps first (5-D grid_sample with a
    singleton T axis, *border* padding), then contract with the target over C
    and divide by sqrt(C)."""
    fmaps = np.asarray(fmaps, dtype=F32)
    targets = np.asarray(targets, dtype=F32)
    coords = np.asarray(coords, dtype=F32)
    B, S, N, D = coords.shape
    assert D == 2
    r = radius
    Wr = 2 * r + 1
    pyr = build_pyramid(fmaps, num_levels)
    d2 = _window_offsets(r)
    delta = np.concatenate([np.zeros_like(d2[..., :1]), d2], axis=-1)  # (t, x, y)
    outs = []
    for lvl, f in enumerate(pyr):
        C, H, W = f.shape[2:]
        c3 = np.concatenate([np.zeros_like(coords[..., :1]), coords], axis=-1)
        cen = c3.reshape(B * S, N, 1, 1, 3) / F32(2 ** lvl)
        smp = bilinear_sampler(f.reshape(B * S, C, 1, H, W), cen + delta[None, None])
        smp = smp.reshape(B, S, C, N, Wr * Wr)
        t = targets.transpose(0, 1, 3, 2)[..., None]  # B S C N 1
        corr = np.sum(t * smp, axis=2, dtype=F32)
        outs.append((corr / F32(math.sqrt(C))).astype(F32))
    return np.concatenate(outs, axis=-1)


def corr_lookup_bf16_autocast(
    fmaps: np.ndarray, targets: np.ndarray, coords: np.ndarray, num_levels: int, radius: int,
    padding_mode: str = "zeros",
) -> np.ndarray:
    """What ``CorrBlock`` computes under ``torch.autocast(bfloat16)`` (SURVEY
    A.6 vi): matmul operands rounded to bf16, float32 accumulation, the
    quotient by sqrt(C) rounded to bf16 (the volume is *stored* in bf16), then
    a float32 lookup.  The pyramid itself stays float32."""
    pyr = build_pyramid(fmaps, num_levels)
    t = round_bf16(np.asarray(targets, dtype=F32))
    B, S, N, C = t.shape
    vols = []
    for f in pyr:
        H, W = f.shape[-2:]
        v = np.matmul(t, round_bf16(f).reshape(B, S, C, H * W)).astype(F32)
        v = round_bf16(v)  # matmul output dtype is bf16
        v = round_bf16(v / F32(math.sqrt(C)))  # bf16 / f32-scalar stays bf16
        vols.append(v.reshape(B, S, N, H, W))
    return lookup_volumes(vols, coords, radius, padding_mode)


def round_bf16(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even float32 -> bfloat16 -> float32."""
    x = np.ascontiguousarray(x, dtype=F32)
    u = x.view(np.uint32).astype(np.uint64)
    lsb = (u >> 16) & 1
    u = (u + 0x7FFF + lsb) & 0xFFFF0000
    return u.astype(np.uint32).view(F32).reshape(x.shape)


# --------------------------------------------------------------------------
# sin/cos encodings
# --------------------------------------------------------------------------
def get_1d_sincos_pos_embed_from_grid(embed_dim: int, pos: np.ndarray) -> np.ndarray:
    """Restates comet/models/utils.py:37-62 (dup :807-832): float64
    ``omega_k = 10000^(-k/(D/2))``, ``[sin(p*omega) | cos(p*omega)]`` cast to
    float32, shape (1, M, D)."""
    assert embed_dim % 2 == 0
    omega = np.arange(embed_dim // 2, dtype=np.float64)
    omega /= embed_dim / 2.0
    omega = 1.0 / 10000 ** omega
    out = np.einsum("m,d->md", np.asarray(pos, dtype=np.float64).reshape(-1), omega)
    emb = np.concatenate([np.sin(out), np.cos(out)], axis=1)
    return emb[None].astype(F32)


def get_2d_sincos_pos_embed_from_grid(embed_dim: int, grid: np.ndarray) -> np.ndarray:
    """Restates comet/models/utils.py:780-804: first half of the channels
    encodes grid[0], second half grid[1]."""
    assert embed_dim % 2 == 0
    a = get_1d_sincos_pos_embed_from_grid(embed_dim // 2, grid[0])
    b = get_1d_sincos_pos_embed_from_grid(embed_dim // 2, grid[1])
    return np.concatenate([a, b], axis=2)


def get_2d_sincos_pos_embed(embed_dim: int, grid_size, return_grid: bool = False):
    """Restates comet/models/utils.py:724-755 -> (1, D, H, W).  grid[0] is the
    x mesh (``meshgrid(grid_w, grid_h, indexing='xy')``), so the channel order
    is [sin_x | cos_x | sin_y | cos_y]."""
    if isinstance(grid_size, tuple):
        gh, gw = grid_size
    else:
        gh = gw = grid_size
    ys = np.arange(gh, dtype=F32)
    xs = np.arange(gw, dtype=F32)
    gx, gy = np.meshgrid(xs, ys, indexing="xy")
    grid = np.stack([gx, gy], axis=0).reshape(2, 1, gh, gw)
    pe = get_2d_sincos_pos_embed_from_grid(embed_dim, grid)
    pe = pe.reshape(1, gh, gw, -1).transpose(0, 3, 1, 2)
    if return_grid:
        return pe, grid
    return pe


def get_1d_sincos_pos_embed(embed_dim: int, length: int, return_grid: bool = False):
    """Restates comet/models/utils.py:758-777 (time encoding used at
    camera_predictor10.py:365-371) -> (1, length, D)."""
    grid = np.arange(length, dtype=F32)
    pe = get_1d_sincos_pos_embed_from_grid(embed_dim, grid)
    if return_grid:
        return pe, grid[None]
    return pe


def get_2d_embedding(xy: np.ndarray, C: int, cat_coords: bool = True) -> np.ndarray:
    """Restates comet/models/utils.py:65-101 (dup :835-871): interleaved
    float32 sin/cos of coordinate * div_term, ``div = arange(0,C,2)*(1000/C)``."""
    xy = np.asarray(xy, dtype=F32)
    B, N, D = xy.shape
    assert D == 2
    div = (np.arange(0, C, 2, dtype=F32) * F32(1000.0 / C)).reshape(1, 1, C // 2)
    pe = np.zeros((B, N, 2 * C), dtype=F32)
    for k in range(2):
        arg = (xy[:, :, k : k + 1] * div).astype(F32)
        pe[:, :, k * C + 0 : (k + 1) * C : 2] = np.sin(arg, dtype=F32)
        pe[:, :, k * C + 1 : (k + 1) * C : 2] = np.cos(arg, dtype=F32)
    if cat_coords:
        pe = np.concatenate([xy, pe], axis=2)
    return pe


# --------------------------------------------------------------------------
# track tokens and the refinement loop
# --------------------------------------------------------------------------
def transformer_dim(corr_levels: int, corr_radius: int, latent_dim: int, fine: bool) -> int:
    """base_track_predictor.py:55-66."""
    d = corr_levels * (corr_radius * 2 + 1) ** 2 + latent_dim * 2
    if fine:
        d += 4 if d % 2 == 0 else 5
    else:
        d += (4 - d % 4) % 4
    return d


def track_tokens(
    fcorrs: np.ndarray,
    coords: np.ndarray,
    track_feats: np.ndarray,
    fmap_hw: Tuple[int, int],
    tdim: int,
) -> np.ndarray:
    """Restates the token assembly of ``BaseTrackerPredictor.forward``
    base_track_predictor.py:165-224.

    fcorrs (B,S,N,LW), coords (B,S,N,2) in level-0 cells, track_feats
    (B,S,N,latent) -> x (B,N,S,tdim) =
    [flow sin/cos (latent) | flow (2) | fcorrs | track_feats | 0-pad]
    + bilinear sample of the 2-D sincos table at coords[:,0] (broadcast over S).
    """
    fcorrs = np.asarray(fcorrs, dtype=F32)
    coords = np.asarray(coords, dtype=F32)
    track_feats = np.asarray(track_feats, dtype=F32)
    B, S, N, LW = fcorrs.shape
    latent = track_feats.shape[-1]
    fc = fcorrs.transpose(0, 2, 1, 3).reshape(B * N, S, LW)
    flows = (coords - coords[:, 0:1]).transpose(0, 2, 1, 3).reshape(B * N, S, 2)
    femb = get_2d_embedding(flows, latent // 2, cat_coords=False)
    femb = np.concatenate([femb, flows], axis=-1)
    tf = track_feats.transpose(0, 2, 1, 3).reshape(B * N, S, latent)
    x = np.concatenate([femb, fc, tf], axis=2)
    if x.shape[2] < tdim:
        x = np.concatenate([x, np.zeros((B * N, S, tdim - x.shape[2]), dtype=F32)], axis=2)
    pos = get_2d_sincos_pos_embed(tdim, (fmap_hw[0], fmap_hw[1]))
    spe = sample_features4d(np.broadcast_to(pos, (B,) + pos.shape[1:]), coords[:, 0])
    x = x + spe.reshape(B * N, 1, tdim)
    return x.reshape(B, N, S, tdim).astype(F32)


def tracker_forward(
    query_points: np.ndarray,
    fmaps: np.ndarray,
    update_fn: Callable[[np.ndarray], np.ndarray],
    feat_update_fn: Callable[[np.ndarray], np.ndarray],
    *,
    iters: int,
    stride: int,
    corr_levels: int,
    corr_radius: int,
    latent_dim: int,
    fine: bool,
    down_ratio: int = 1,
    efficient_corr: bool = False,
):
    """Restates the loop of ``BaseTrackerPredictor.forward``
    base_track_predictor.py:95-262 for ``TRACKorPOSE=False`` (every live
    caller).  ``update_fn`` stands for ``self.updateformer`` ((B,N,S,D) ->
    (B,N,S,latent+2)) and ``feat_update_fn`` for
    ``ffeat_updater(norm(.))`` ((B*N*S,latent) -> same); both are opaque to
    the hot path.  Returns (coord_preds, track_feats, query_track_feat,
    tokens_per_iteration)."""
    q = np.asarray(query_points, dtype=F32)
    fmaps = np.asarray(fmaps, dtype=F32)
    B, N, D = q.shape
    _, S, C, HH, WW = fmaps.shape
    assert D == 2
    if down_ratio > 1:  # quirk A.6(i): guards both divisions
        q = q / F32(float(down_ratio))
        q = q / F32(float(stride))
    coords = np.repeat(q.reshape(B, 1, N, 2), S, axis=1).copy()
    qfeat = sample_features4d(fmaps[:, 0], coords[:, 0])
    track_feats = np.repeat(qfeat[:, None], S, axis=1).copy()
    backup = coords.copy()
    tdim = transformer_dim(corr_levels, corr_radius, latent_dim, fine)
    preds, toks = [], []
    for _ in range(iters):
        if efficient_corr:
            fcorrs = efficient_corr_lookup(fmaps, track_feats, coords, corr_levels, corr_radius)
        else:
            fcorrs = corr_lookup(fmaps, track_feats, coords, corr_levels, corr_radius, "zeros")
        x = track_tokens(fcorrs, coords, track_feats, (HH, WW), tdim)
        toks.append(x)
        delta = np.asarray(update_fn(x), dtype=F32).reshape(B * N, S, latent_dim + 2)
        dco = delta[:, :, :2]
        dfe = delta[:, :, 2:].reshape(B * N * S, latent_dim)
        tf_ = track_feats.transpose(0, 2, 1, 3).reshape(B * N * S, latent_dim)
        tf_ = np.asarray(feat_update_fn(dfe), dtype=F32) + tf_
        track_feats = np.ascontiguousarray(tf_.reshape(B, N, S, latent_dim).transpose(0, 2, 1, 3))
        coords = coords + dco.reshape(B, N, S, 2).transpose(0, 2, 1, 3)
        coords[:, 0] = backup[:, 0]
        if down_ratio > 1:
            preds.append(coords * F32(stride) * F32(down_ratio))
        else:
            preds.append(coords * F32(stride))
    return preds, track_feats, qfeat, toks
