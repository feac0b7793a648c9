#!/usr/bin/env python
"""Spread of max|xhat - oracle| over repeated runs of the full-size config-2 forward (the scatter-add order varies)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch, bench
import cdl_oracle as O
import cdlnet_video_b200 as cb
d = torch.device("cuda", 0)
K, M = bench.CFG["K"], bench.CFG["M"]
A, B, u = bench.synthetic_weights(torch, d)
for seed in (0, 1):
    clean, y = bench.synthetic_clip(torch, 1, seed=seed, device=d)
    plan = cb.Plan(3, 1, 1, M, K, bench.CLIP, (7, 7, 7), 2, precision="tf32")
    plan.set_weights(A, B, torch.zeros(K, 2, M, device=d))
    yp, _, _ = plan.preprocess(y)
    z0 = plan.new_code(); plan.analysis_step(0, yp, z0, None, first=True); z0 = plan.export_code(z0)
    q = torch.quantile(z0[0].abs().reshape(M, -1)[:, ::8].float(), 0.85, dim=1)
    t = bench.thresholds_from_quantile(torch, q, u)
    plan.set_weights(A, B, t)
    c = torch.full((1,), bench.SIGMA / 255.0, device=d)
    xr, zr, *_ = O.forward_t(y.cpu(), [a.cpu() for a in A], [b.cpu() for b in B], t.cpu().reshape(K, 2, M, 1, 1, 1), 2, bench.SIGMA, True, 1)
    errs = []
    for rep in range(12):
        xhat, z = plan.denoise(y, None, c)
        errs.append((xhat.cpu() - xr).abs().max().item())
    print(f"seed {seed}: max|xhat - oracle| over 12 runs: min {min(errs):.2e} max {max(errs):.2e}  all {['%.2e' % e for e in errs]}")
