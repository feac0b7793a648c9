#!/usr/bin/env python
"""Where does the tf32 error on xhat come from?  cfg-2 full size, vs the fp32 CUDA-core family (≈ oracle to 1e-6)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import cdlnet_video_b200 as cb
d = torch.device("cuda", 0)
K, M = bench.CFG["K"], bench.CFG["M"]
A, B, u = bench.synthetic_weights(torch, d)
for seed in (0, 1, 2):
    clean, y = bench.synthetic_clip(torch, 1, seed=seed, device=d)
    ptc = cb.Plan(3, 1, 1, M, K, bench.CLIP, (7, 7, 7), 2, precision="tf32")
    p32 = cb.Plan(3, 1, 1, M, K, bench.CLIP, (7, 7, 7), 2, precision="fp32")
    ptc.set_weights(A, B, torch.zeros(K, 2, M, device=d))
    yp, _, mean = ptc.preprocess(y)
    z0 = ptc.new_code(); ptc.analysis_step(0, yp, z0, None, first=True); z0 = ptc.export_code(z0)
    q = torch.quantile(z0[0].abs().reshape(M, -1)[:, ::8].float(), 0.85, dim=1)
    t = bench.thresholds_from_quantile(torch, q, u)
    ptc.set_weights(A, B, t); p32.set_weights(A, B, t)
    c = torch.full((1,), bench.SIGMA / 255.0, device=d)
    x32, z32 = p32.denoise(y, None, c)
    xs = []
    for rep in range(3):
        xtc, ztc = ptc.denoise(y, None, c)
        xs.append((xtc - x32).abs().max().item())
    # tc iterations, fp32 final synthesis
    r = torch.empty_like(yp)
    p32.synthesis_step(0, p32.import_code(ztc), r, residual=False)
    xa = p32.postprocess(r, mean)
    # fp32 iterations, tc final synthesis
    ptc.synthesis_step(0, ptc.import_code(z32), r, residual=False)
    xb = ptc.postprocess(r, mean)
    print(f"seed {seed}: all-tc {['%.2e' % v for v in xs]} | tc iters + fp32 Dz {(xa - x32).abs().max().item():.2e} | fp32 iters + tc Dz {(xb - x32).abs().max().item():.2e} | z err {(ztc - z32).abs().max().item():.2e}")
