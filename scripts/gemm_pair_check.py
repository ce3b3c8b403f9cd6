"""CTA-pair (W-tile multicast) GEMM against the single-CTA one: bit-identical results, then timings.
python scripts/gemm_pair_check.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from comet_pose_estimation_b200 import update_former_tc as tc, _lib

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
for np_ in (3, 1):
    run = tc._Run(tc._Weights(), np_, dev)
    for (M, K, N, kw) in ((9216, 384, 1152, {}), (9216, 1536, 384, dict(resid=True)), (8192, 384, 768, {}),
                          (9216, 384, 1536, dict(gelu=True, want_planes=True)), (1024, 384, 384, {}), (9088, 664, 384, {})):
        x = torch.randn(M, K, device=dev, generator=g)
        w = torch.randn(N, K, device=dev, generator=g) / K ** 0.5
        b = torch.randn(N, device=dev, generator=g)
        r = torch.randn(M, N, device=dev, generator=g) if kw.get("resid") else None
        kw2 = {k: v for k, v in kw.items() if k != "resid"}
        xp = run.split(x)
        outs = []
        for pair in (0, 7):
            _lib.check(_lib.lib.comet_set_option(_lib.OPT_GEMM_PAIR, pair))
            o, op = run.linear(xp, w, b, resid=r, **kw2)
            torch.cuda.synchronize()
            outs.append((o.clone(), None if op is None else op.clone()))
        same = torch.equal(outs[0][0], outs[1][0]) and (outs[0][1] is None or torch.equal(outs[0][1], outs[1][1]))
        print(f"np={np_} M={M} K={K} N={N} {kw}: pair result identical: {same}", flush=True)
