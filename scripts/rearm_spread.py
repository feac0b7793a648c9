#!/usr/bin/env python
"""Run-to-run spread of the tensor-core forward (scatter-add order) vs fused/stepwise/re-armed variants."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np, torch
import cdlnet_video_b200 as cb
from test_tc_gpu import _weights
d = torch.device("cuda", 0)
dims, N, M, K = (8, 32, 64), 2, 169, 4
A, B, g = _weights(M, K, 3, 0.7 / np.sqrt(2.0 * M * 343 / 8))
t = (torch.rand(K, 2, M, generator=g) * 0.01).to(d)
y = torch.rand(N, 1, *dims, generator=g).to(d)
c = torch.tensor([0.1, 0.06], device=d)
plan = cb.Plan(3, N, 1, M, K, dims, (7, 7, 7), 2, precision="tf32")
plan.set_weights([a.to(d) for a in A], [b.to(d) for b in B], t)
x0, z0 = plan.denoise(y, None, c)


def stepwise(rearm):
    plan.set_rearm(rearm)
    yp, _, mean = plan.preprocess(y)
    code, r = plan.new_code(), torch.empty_like(yp)
    plan.analysis_step(0, yp, code, c, first=True)
    for k in range(1, K):
        plan.synthesis_step(k, code, r, yp, None, residual=True)
        plan.analysis_step(k, r, code, c)
    xp = torch.empty_like(yp)
    plan.synthesis_step(0, code, xp, residual=False)
    return plan.postprocess(xp, mean), plan.export_code(code)


for i in range(6):
    x1, z1 = plan.denoise(y, None, c)
    x2, z2 = stepwise(False)
    x3, z3 = stepwise(True)
    print(f"run {i}: fused-fused {(x1 - x0).abs().max().item():.2e}/{(z1 - z0).abs().max().item():.2e}  stepwise {(x2 - x0).abs().max().item():.2e}/{(z2 - z0).abs().max().item():.2e}"
          f"  rearm {(x3 - x0).abs().max().item():.2e}/{(z3 - z0).abs().max().item():.2e}")
