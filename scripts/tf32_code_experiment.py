#!/usr/bin/env python
"""Numerics experiment for the next round: what if the sparse code itself is STORED rounded to tf32 (RNE) after every
analysis step (so that the synthesis could read it as a tensor-core operand straight from shared memory, no
producer warps)?  cfg-2 full size, error on xhat vs the fp32 CUDA-core family."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import cdlnet_video_b200 as cb
d = torch.device("cuda", 0)
K, M = bench.CFG["K"], bench.CFG["M"]
A, B, u = bench.synthetic_weights(torch, d)


def rne_tf32_(x):
    b = x.view(torch.int32)
    b.add_(0x1000).bitwise_and_(-8192)          # 0xffffe000
    return x


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
    out = {}
    for mode in ("current", "stored-tf32", "stored-tf32-all"):
        code, r = ptc.new_code(), torch.empty_like(yp)
        ptc.analysis_step(0, yp, code, c, first=True)
        for k in range(1, K):
            if mode != "current":
                rne_tf32_(code)
            ptc.synthesis_step(k, code, r, yp, None, residual=True)
            ptc.analysis_step(k, r, code, c)
        if mode == "stored-tf32-all":
            rne_tf32_(code)
        ptc.synthesis_step(0, code, r, residual=False)
        x = ptc.postprocess(r, mean)
        out[mode] = ((x - x32).abs().max().item(), (ptc.export_code(code) - z32).abs().max().item())
    print(f"seed {seed}: " + " | ".join(f"{m}: xhat {e[0]:.2e} z {e[1]:.2e}" for m, e in out.items()))
