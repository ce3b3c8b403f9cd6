"""Generate the golden fixtures by EXECUTING THE UNMODIFIED REFERENCE.

Run in the build container only (it needs /root/reference, which does not
exist on the GPU box):

    python tests/golden/make_golden.py

The reference hot-path modules are imported in place with the recipe of
SURVEY.md section 8(c): sys.path = [ref/comet/models, ref, ref/comet] and one
stub module (``train_eval_func``, imported by comet/models/utils.py:26 only
for a class that the hot path never touches).  No reference source is copied;
only tensors produced by running it are stored (``tests/golden/*.npz``).
"""
from __future__ import annotations

import hashlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("COMET_REFERENCE", "/root/reference")
sys.path.insert(0, HERE)
import cases  # noqa: E402


def import_reference():
    sys.path[:0] = [REF + "/comet/models", REF, REF + "/comet"]
    stub = types.ModuleType("train_eval_func")

    class QuaternionCameras:  # placeholder, never instantiated on this path
        pass

    stub.QuaternionCameras = QuaternionCameras
    sys.modules["train_eval_func"] = stub
    import torch  # noqa: F401
    from models.track_modules import blocks
    from models.track_modules import base_track_predictor as btp
    import utils as rutils

    return blocks, btp, rutils


def digest(*arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:16]


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"{name}.npz  {os.path.getsize(path) / 1024:.1f} KiB")


def main():
    import torch

    torch.manual_seed(0)
    torch.set_grad_enabled(False)
    blocks, btp, ru = import_reference()
    t = torch.from_numpy

    # ---- CorrBlock / EfficientCorrBlock ---------------------------------
    out = {}
    for name, (kw, L, r) in cases.CORR_CASES.items():
        fmaps, targets, coords = cases.corr_case(**kw)
        out[name + "/digest"] = np.frombuffer(digest(fmaps, targets, coords).encode(), dtype=np.uint8)
        cb = blocks.CorrBlock(t(fmaps), num_levels=L, radius=r)
        cb.corr(t(targets))
        out[name + "/zeros"] = cb.sample(t(coords)).numpy()
        if name in ("small_ragged", "tiny_odd"):
            for l in range(L):
                out[f"{name}/pyr{l}"] = cb.fmaps_pyramid[l].numpy()
                out[f"{name}/vol{l}"] = cb.corrs_pyramid[l].numpy()
        cbb = blocks.CorrBlock(t(fmaps), num_levels=L, radius=r, padding_mode="border")
        cbb.corr(t(targets))
        out[name + "/border"] = cbb.sample(t(coords)).numpy()
        eb = blocks.EfficientCorrBlock(t(fmaps), num_levels=L, radius=r)
        out[name + "/efficient"] = eb.sample(t(coords), t(targets)).numpy()
        with torch.autocast("cpu", dtype=torch.bfloat16):
            ca = blocks.CorrBlock(t(fmaps), num_levels=L, radius=r)
            ca.corr(t(targets))
            assert ca.corrs_pyramid[0].dtype == torch.bfloat16
            out[name + "/zeros_bf16"] = ca.sample(t(coords)).float().numpy()
    save("corr_blocks", **out)

    # ---- samplers / encodings ------------------------------------------
    out = {}
    inp, xy = cases.sampler_case(21, B=2, C=3, H=9, W=7, Ho=4, Wo=5)
    for pm in ("zeros", "border"):
        for ac in (True, False):
            out[f"bs4/{pm}/{int(ac)}"] = ru.bilinear_sampler(t(inp), t(xy), align_corners=ac, padding_mode=pm).numpy()
    inp5, txy = cases.sampler_case(22, B=2, C=3, H=6, W=8, Ho=3, Wo=4, T=3)
    for pm in ("zeros", "border"):
        out[f"bs5/{pm}"] = ru.bilinear_sampler(t(inp5), t(txy), padding_mode=pm).numpy()
    inp1, txy1 = cases.sampler_case(23, B=2, C=4, H=6, W=8, Ho=3, Wo=4, T=1)
    txy1[..., 0] = 0
    out["bs5_t1/border"] = ru.bilinear_sampler(t(inp1), t(txy1)).numpy()
    rng = np.random.default_rng(24)
    pts = np.stack([rng.uniform(-1, 8, (2, 11)), rng.uniform(-1, 10, (2, 11))], -1).astype(np.float32)
    out["sf4d/pts"] = pts
    out["sf4d/out"] = ru.sample_features4d(t(inp), t(pts)).numpy()
    for C, scale in ((64, 3.0), (16, 0.7), (64, 40.0)):
        xyv = cases.embed_case(30 + C, 3, 5, scale)
        out[f"emb2d/{C}/{scale}/xy"] = xyv
        out[f"emb2d/{C}/{scale}/nocat"] = ru.get_2d_embedding(t(xyv), C, cat_coords=False).numpy()
        out[f"emb2d/{C}/{scale}/cat"] = ru.get_2d_embedding(t(xyv), C, cat_coords=True).numpy()
    out["sincos2d/216_31"] = ru.get_2d_sincos_pos_embed(216, (31, 31)).numpy()
    full = ru.get_2d_sincos_pos_embed(664, (64, 64)).numpy()
    out["sincos2d/664_64/rows"] = full[:, :, ::9, :]  # every 9th row, all columns/channels
    out["sincos2d/664_64/sum"] = np.array([full.astype(np.float64).sum(), np.abs(full).astype(np.float64).sum()])
    out["sincos2d/12_h3w5"] = ru.get_2d_sincos_pos_embed(12, (3, 5)).numpy()
    pe, grid = ru.get_2d_sincos_pos_embed(8, 4, return_grid=True)
    out["sincos2d/8_4"], out["sincos2d/8_4/grid"] = pe.numpy(), grid.numpy()
    out["sincos1d/768_16"] = ru.get_1d_sincos_pos_embed(768, 16).numpy()
    out["sincos1d/768_64"] = ru.get_1d_sincos_pos_embed(768, 64).numpy()
    pe1, g1 = ru.get_1d_sincos_pos_embed(10, 7, return_grid=True)
    out["sincos1d/10_7"], out["sincos1d/10_7/grid"] = pe1.numpy(), g1.numpy()
    pos = np.array([0.0, 0.5, 3.25, 100.0, -2.0], dtype=np.float32)
    out["sincos1dgrid/pos"] = pos
    out["sincos1dgrid/14"] = ru.get_1d_sincos_pos_embed_from_grid(14, t(pos)).numpy()
    save("samplers_encodings", **out)

    # ---- BaseTrackerPredictor: tokens and refinement loop ---------------
    from types import SimpleNamespace as NS

    def cfg(eff=False, conf=False):
        return NS(track_conf=conf, MODEL=NS(TRACK=NS(efficient_corr=eff)))

    class Abort(Exception):
        pass

    out = {}
    # (name, ctor kwargs, case kwargs, iters, down_ratio, eff, abort_after_first)
    specs = [
        ("coarse_tiny", dict(stride=4, corr_levels=5, corr_radius=2, latent_dim=16, hidden_size=32, depth=1,
                             use_spaceatt=True, fine=False),
         dict(seed=41, B=1, S=4, C=16, H=16, W=16, N=7, stride=4, down_ratio=2), 3, 2, False, False),
        ("coarse_tiny_eff", dict(stride=4, corr_levels=2, corr_radius=3, latent_dim=16, hidden_size=32, depth=1,
                                 use_spaceatt=True, fine=False),
         dict(seed=42, B=1, S=4, C=16, H=16, W=16, N=7, stride=4, down_ratio=2), 2, 2, True, False),
        ("fine_tiny", dict(stride=1, corr_levels=3, corr_radius=3, latent_dim=32, hidden_size=32, depth=1,
                           use_spaceatt=False, fine=True),
         dict(seed=43, B=5, S=3, C=32, H=31, W=31, N=1, stride=1, down_ratio=1), 2, 1, False, False),
        ("coarse_full_it0", dict(stride=4, corr_levels=5, corr_radius=4, latent_dim=128, hidden_size=16, depth=1,
                                 use_spaceatt=True, fine=False),
         dict(seed=44, B=1, S=3, C=128, H=64, W=64, N=16, stride=4, down_ratio=2), 1, 2, False, True),
        ("fine_full_it0", dict(stride=1, corr_levels=3, corr_radius=3, latent_dim=32, hidden_size=16, depth=1,
                               use_spaceatt=False, fine=True),
         dict(seed=45, B=8, S=3, C=32, H=31, W=31, N=1, stride=1, down_ratio=1), 1, 1, False, True),
    ]
    for name, ck, case_kw, iters, dr, eff, abort in specs:
        torch.manual_seed(7)
        m = btp.BaseTrackerPredictor(cfg=cfg(eff), **ck).eval()
        fmaps, q = cases.tracker_case(**case_kw)
        toks = []

        def hook(mod, args):
            toks.append(args[0].detach().clone().numpy())
            if abort:
                raise Abort()

        h = m.updateformer.register_forward_pre_hook(hook)
        try:
            res = m(query_points=t(q), fmaps=t(fmaps), iters=iters, return_feat=True, down_ratio=dr,
                    TRACKorPOSE=False)
        except Abort:
            res = None
        h.remove()
        out[name + "/digest"] = np.frombuffer(digest(fmaps, q).encode(), dtype=np.uint8)
        for i, x in enumerate(toks):
            out[f"{name}/tok{i}"] = x
        if res is not None:
            preds, vis, tf, qf, conf = res
            for i, p in enumerate(preds):
                out[f"{name}/pred{i}"] = p.numpy()
            if vis is not None:
                out[name + "/vis"] = vis.numpy()
            out[name + "/track_feats"] = tf.numpy()
            out[name + "/query_feat"] = qf.numpy()
            for k, v in m.state_dict().items():
                out[f"{name}/sd/{k}"] = v.numpy()
            # the same run in float64: the yard-stick for the per-iteration bars (tests/test_tracker_loop.py).  The
            # reference casts its float64-built sin/cos table to float32 (utils.py:724-755) and F.grid_sample refuses
            # mixed dtypes, so for this run only the table is widened back to float64 (same values).
            toks64 = []
            m64 = m.double()
            h = m64.updateformer.register_forward_pre_hook(lambda mod, args: toks64.append(args[0].detach().clone().numpy()))
            orig = btp.get_2d_sincos_pos_embed
            btp.get_2d_sincos_pos_embed = lambda *a, **k: orig(*a, **k).double()
            try:
                res64 = m64(query_points=t(q).double(), fmaps=t(fmaps).double(), iters=iters, return_feat=True,
                            down_ratio=dr, TRACKorPOSE=False)
            finally:
                btp.get_2d_sincos_pos_embed = orig
                h.remove()
            for i, p64 in enumerate(res64[0]):
                out[f"{name}/pred64_{i}"] = p64.numpy()
                out[f"{name}/tok64_{i}"] = toks64[i]
            out[name + "/track_feats64"] = res64[2].numpy()
            if res64[1] is not None:
                out[name + "/vis64"] = res64[1].numpy()
    save("tracker", **out)

    # ---- refine_track: patch extraction -> ShallowEncoder -> fine tracker -> score ----------------------
    # kornia is not installed here: the two functions refine_track.py:20-21 imports are restated from their
    # documented semantics -- create_meshgrid(h, w, normalized_coordinates=True) -> (1,h,w,2) with
    # [...,0] = x, [...,1] = y in linspace(-1,1); spatial_expectation2d(heat (B,N,h,w), True) -> (B,N,2) = E[x], E[y].
    def create_meshgrid(h, w, normalized_coordinates=True, device=None, dtype=torch.float32):
        xs = torch.linspace(-1, 1, w, device=device, dtype=dtype) if normalized_coordinates else torch.arange(w, device=device, dtype=dtype)
        ys = torch.linspace(-1, 1, h, device=device, dtype=dtype) if normalized_coordinates else torch.arange(h, device=device, dtype=dtype)
        gy, gx = torch.meshgrid(ys, xs, indexing="ij")
        return torch.stack([gx, gy], -1)[None]

    def spatial_expectation2d(heat, normalized_coordinates=True):
        b, n, h, w = heat.shape
        g = create_meshgrid(h, w, normalized_coordinates, heat.device, heat.dtype).reshape(-1, 2)
        flat = heat.reshape(b, n, -1)
        return torch.stack([(flat * g[:, 0]).sum(-1), (flat * g[:, 1]).sum(-1)], -1)

    kornia = types.ModuleType("kornia")
    for name in ("kornia.utils", "kornia.utils.grid", "kornia.geometry", "kornia.geometry.subpix", "kornia.geometry.subpix.dsnt"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["kornia"] = kornia
    sys.modules["kornia.utils.grid"].create_meshgrid = create_meshgrid
    sys.modules["kornia.geometry.subpix.dsnt"].spatial_expectation2d = spatial_expectation2d
    sys.modules["kornia.geometry.subpix"].dsnt = sys.modules["kornia.geometry.subpix.dsnt"]
    import refine_track as rrt

    out = {}
    for name, case_kw, hidden in (("refine_small", dict(seed=51, B=1, S=3, N=5, HW=64), 32),):
        torch.manual_seed(9)
        fnet = blocks.ShallowEncoder(input_dim=3).eval()
        ftr = btp.BaseTrackerPredictor(stride=1, corr_levels=3, corr_radius=3, latent_dim=32, hidden_size=hidden,
                                       depth=1, use_spaceatt=False, fine=True, cfg=cfg()).eval()
        # random-init deltas of a few pixels per iteration make the 6-iteration loop chaotic (sin/cos(flow * 1e3)
        # feeds back into the next delta), so that any two fp32 implementations diverge; damping the output head
        # keeps the refinement contractive and the final tracks / scores comparable at a tight bar
        for prm in ftr.updateformer.flow_head.parameters():
            prm.mul_(0.02)
        images, coarse = cases.refine_case(**case_kw)
        toks = []
        h = ftr.updateformer.register_forward_pre_hook(lambda mod, args: toks.append(args[0].detach().clone().numpy()))
        refined, score = rrt.refine_track(t(images), fnet, ftr, t(coarse), compute_score=True)
        h.remove()
        out[name + "/digest"] = np.frombuffer(digest(images, coarse).encode(), dtype=np.uint8)
        out[name + "/refined"] = refined.numpy()
        out[name + "/score"] = score.numpy()
        for i, x in enumerate(toks):
            out[f"{name}/tok{i}"] = x
        # the encoder alone on a handful of patches (pins ShallowEncoder separately from the loop)
        pin = t(np.random.default_rng(52).random((4, 3, 31, 31)).astype(np.float32))
        out[name + "/enc_in"] = pin.numpy()
        out[name + "/enc_out"] = fnet(pin).numpy()
        for k, v in fnet.state_dict().items():
            out[f"{name}/fnet/{k}"] = v.numpy()
        for k, v in ftr.state_dict().items():
            out[f"{name}/ftr/{k}"] = v.numpy()
    save("refine", **out)

    # ---- BasicEncoder / process_images_to_fmaps (the step before the path) ---------------------------
    out = {}
    enc = blocks.BasicEncoder(input_dim=3, output_dim=128, stride=4).eval()
    sd = cases.seeded_state_dict({k: tuple(v.shape) for k, v in enc.state_dict().items()}, 91)
    enc.load_state_dict({k: t(v) for k, v in sd.items()}, strict=True)
    img = np.random.default_rng(92).random((1, 2, 3, 96, 80)).astype(np.float32)       # H != W, not multiples of 8
    out["basic/images"] = img
    out["basic/keys"] = np.array(sorted(sd), dtype=object).astype(str)
    x = F_interp(torch, t(img).reshape(2, 3, 96, 80))
    out["basic/fmaps"] = enc(x).reshape(1, 2, 128, x.shape[-2] // 4, x.shape[-1] // 4).numpy()
    out["basic/encoder_only"] = enc(t(img).reshape(2, 3, 96, 80)).numpy()
    save("encoders", **out)


def F_interp(torch, x):
    """track_predictor.py:136-145 with down_ratio 2."""
    return torch.nn.functional.interpolate(x, scale_factor=0.5, mode="bilinear", align_corners=True)


if __name__ == "__main__":
    main()
