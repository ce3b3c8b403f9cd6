"""Seeded synthetic inputs shared by the fixture generator (which runs the
reference) and by the parity tests (which run the oracle and the CUDA path).

Nothing here imports the reference or the oracle.  Inputs follow SURVEY.md
section 8(d): fmaps ~ N(0,1), targets ~ N(0,1), coords ~ U[0, size-1) with a
fraction drawn from a wider range to exercise padding.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def corr_case(seed, B, S, C, H, W, N, oob_frac=0.2, margin=6.0):
    rng = np.random.default_rng(seed)
    fmaps = rng.standard_normal((B, S, C, H, W)).astype(F32)
    targets = rng.standard_normal((B, S, N, C)).astype(F32)
    x = rng.uniform(0, W - 1, (B, S, N)).astype(F32)
    y = rng.uniform(0, H - 1, (B, S, N)).astype(F32)
    oob = rng.uniform(0, 1, (B, S, N)) < oob_frac
    xo = rng.uniform(-margin, W - 1 + margin, (B, S, N)).astype(F32)
    yo = rng.uniform(-margin, H - 1 + margin, (B, S, N)).astype(F32)
    coords = np.stack([np.where(oob, xo, x), np.where(oob, yo, y)], axis=-1).astype(F32)
    # a few exact-integer and exact-border coordinates (weights 0/1 edge cases)
    flat = coords.reshape(-1, 2)
    if flat.shape[0] >= 4:
        flat[0] = (0.0, 0.0)
        flat[1] = (W - 1.0, H - 1.0)
        flat[2] = (np.floor(W / 2), np.floor(H / 3))
        flat[3] = (-0.5, H - 0.5)
    return fmaps, targets, coords


# name -> (kwargs of corr_case, num_levels, radius)
CORR_CASES = {
    # small, ragged map (H != W), odd sizes so pooling floors
    "small_ragged": (dict(seed=11, B=2, S=3, C=16, H=13, W=10, N=9), 3, 2),
    # slice of the coarse configuration (C=128, 64x64, L=5, r=4)
    "coarse_slice": (dict(seed=12, B=1, S=2, C=128, H=64, W=64, N=24), 5, 4),
    # fine-tracker configuration (one query per 31x31 patch, L=3, r=3)
    "fine_slice": (dict(seed=13, B=6, S=3, C=32, H=31, W=31, N=1, margin=4.0), 3, 3),
    # single level / radius 1 / C not a multiple of 4
    "tiny_odd": (dict(seed=14, B=1, S=1, C=6, H=5, W=7, N=5, margin=3.0), 1, 1),
    # map smaller than the window at the top levels
    "deep_pyramid": (dict(seed=15, B=1, S=2, C=8, H=16, W=16, N=6), 4, 3),
}


def sampler_case(seed, B, C, H, W, Ho, Wo, T=None):
    rng = np.random.default_rng(seed)
    if T is None:
        inp = rng.standard_normal((B, C, H, W)).astype(F32)
        xy = np.stack(
            [rng.uniform(-2, W + 1, (B, Ho, Wo)), rng.uniform(-2, H + 1, (B, Ho, Wo))], -1
        ).astype(F32)
        return inp, xy
    inp = rng.standard_normal((B, C, T, H, W)).astype(F32)
    D = 2
    txy = np.stack(
        [
            rng.uniform(-0.5, T - 0.5, (B, D, Ho, Wo)),
            rng.uniform(-2, W + 1, (B, D, Ho, Wo)),
            rng.uniform(-2, H + 1, (B, D, Ho, Wo)),
        ],
        -1,
    ).astype(F32)
    return inp, txy


def embed_case(seed, B, N, scale):
    rng = np.random.default_rng(seed)
    xy = (rng.standard_normal((B, N, 2)) * scale).astype(F32)
    xy[0, 0] = 0.0
    return xy


def tracker_case(seed, B, S, C, H, W, N, stride, down_ratio):
    """query points in *pixels* of the (virtual) input image."""
    rng = np.random.default_rng(seed)
    fmaps = rng.standard_normal((B, S, C, H, W)).astype(F32)
    scale = stride * (down_ratio if down_ratio > 1 else 1) if down_ratio > 1 else 1
    q = np.stack(
        [rng.uniform(1, (W - 2), (B, N)), rng.uniform(1, (H - 2), (B, N))], -1
    ).astype(F32) * F32(scale)
    return fmaps, q


def refine_case(seed, B, S, N, HW):
    """Images (B,S,3,HW,HW) ~ U[0,1) and a coarse track prediction (B,S,N,2) in pixels: query points inside the image,
    later frames displaced by a few pixels; some tracks close to the border so the patch corner clamps."""
    rng = np.random.default_rng(seed)
    images = rng.random((B, S, 3, HW, HW)).astype(F32)
    q = rng.uniform(3.0, HW - 4.0, (B, 1, N, 2))
    q[:, :, 0] = [[2.25, HW - 3.5]]           # corner track: top-left patch corner clamps on both axes
    coarse = (q + rng.normal(0.0, 2.0, (B, S, N, 2))).astype(F32)
    coarse[:, 0] = q[:, 0].astype(F32)
    coarse = np.clip(coarse, 0.0, HW - 1.001).astype(F32)
    return images, coarse


# ---- full-size (shipped) configurations: weights by name from a counter-based seed -------------------------------
def seeded_state_dict(shapes, seed, flow_head_gain=1.0):
    """Deterministic weights for a module, keyed by parameter NAME (not by construction order), so the fixture
    generator (reference classes) and the tests (this package's classes) build bit-identical 170 MB state dicts
    without committing them.  ``shapes``: {name: shape}.  Distribution: what torch's default initialisers give in
    magnitude -- matrices / conv kernels U(-1/sqrt(fan_in), 1/sqrt(fan_in)), biases U(-0.05, 0.05), 1-D ``weight``
    (affine norms) 1 + 0.1 N(0,1), the learned virtual tracks N(0,1)."""
    import zlib

    out = {}
    for name in sorted(shapes):
        shape = tuple(int(s) for s in shapes[name])
        rng = np.random.default_rng([int(seed), zlib.crc32(name.encode())])
        if name.endswith("virual_tracks"):
            a = rng.standard_normal(shape)
        elif len(shape) >= 2:
            fan_in = int(np.prod(shape[1:]))
            b = 1.0 / np.sqrt(fan_in)
            a = rng.uniform(-b, b, shape)
            if "flow_head" in name:
                a = a * flow_head_gain
        elif name.endswith("weight"):
            a = 1.0 + 0.1 * rng.standard_normal(shape)
        else:
            a = rng.uniform(-0.05, 0.05, shape)
            if "flow_head" in name:
                a = a * flow_head_gain
        out[name] = a.astype(F32)
    return out


# shipped tracker configuration (abl_ours.yaml:399-428): coarse 4 iterations, fine 6 iterations (refine_track.py:136)
FULL_COARSE_CTOR = dict(stride=4, corr_levels=5, corr_radius=4, latent_dim=128, hidden_size=384, depth=6,
                        use_spaceatt=True, fine=False)
FULL_FINE_CTOR = dict(stride=1, corr_levels=3, corr_radius=3, latent_dim=32, hidden_size=256, depth=4,
                      use_spaceatt=False, fine=True)   # abl_ours.yaml:419-428
FULL_COARSE_CASE = dict(seed=61, B=1, S=16, C=128, H=64, W=64, N=512, stride=4, down_ratio=2)
FULL_COARSE_ITERS = 4
FULL_REFINE_CASE = dict(seed=62, B=1, S=16, N=512, HW=512)
FULL_SEEDS = dict(coarse=71, fine=72, fnet=73, camera=74)
