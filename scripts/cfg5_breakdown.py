#!/usr/bin/env python
"""Per-kernel time per Mvoxel at 1080p (one slab of config 5) vs the config-2 clip shape: python scripts/cfg5_breakdown.py [frames]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cdlnet_video_b200 as cb
d = torch.device("cuda", 0)
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 32
K, M = 6, 169
g = torch.Generator().manual_seed(0)
W = (torch.randn(M, 1, 7, 7, 7, generator=g) * 0.004).to(d)
for name, N, dims in (("cfg2 4 clips", 4, (16, 256, 256)), (f"1080p x {frames}", 1, (frames, 1080, 1920))):
    plan = cb.Plan(3, N, 1, M, K, dims, (7, 7, 7), 2, precision="tf32")
    plan.set_weights([W] * K, [W] * K, torch.rand(K, 2, M, device=d) * 0.01)
    y = torch.rand(N, 1, *dims, device=d)
    c = torch.full((N,), 0.1, device=d)
    yp, _, mean = plan.preprocess(y)
    code, r = plan.new_code(), torch.empty_like(yp)
    plan.analysis_step(0, yp, code, c, first=True)
    ta, ts = [], []
    for k in range(1, K):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(); plan.synthesis_step(k, code, r, yp, None, residual=True)
        e[1].record(); plan.analysis_step(k, r, code, c)
        e[2].record(); torch.cuda.synchronize()
        ts.append(e[0].elapsed_time(e[1])); ta.append(e[1].elapsed_time(e[2]))
    mv = N * dims[0] * dims[1] * dims[2] / 1e6
    print(f"{name:16s}: synthesis {min(ts):8.3f} ms = {min(ts) / mv * 1e3:6.1f} us/Mvox | analysis {min(ta):8.3f} ms = {min(ta) / mv * 1e3:6.1f} us/Mvox")
    del plan, code, r, yp, y
    torch.cuda.empty_cache()
