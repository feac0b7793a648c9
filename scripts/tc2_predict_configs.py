#!/usr/bin/env python
"""Predicted parity of the 2-D tensor-core family at the BASELINE configurations, on the CPU: the emulation of
scripts/tc2_emulate_cpu.py (tf32 analysis + tf32 residual synthesis, exact final D z) against the fp32 oracle, on
bench_configs.py's synthetic weights and inputs, one or two samples per configuration (the oracle needs ~25 s per
1024^2 image).  The emulator reproduces GPU-measured errors to 3 %."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "scripts"))

import numpy as np
import torch

import bench_configs as bc
import cdl_oracle as O
from tc2_emulate_cpu import forward

CFG = {"cfg1b": ("cdl", (20, 32, 7, 1, 1), (1, 1, 256, 256), False, 25.0),
       "cfg4": ("gabor", (30, 64, 7, 1, 3), (2, 3, 512, 512), False, 25.0),
       "cfg3": ("cdl", (42, 64, 7, 1, 3), (1, 3, 1024, 1024), True, 10.0)}

for name in sys.argv[1:] or ["cfg1b", "cfg4", "cfg3"]:
    kind, (K, M, P, s, C), shape, use_mask, sigma = CFG[name]
    net = bc.make_net(kind, K, M, P, s, C)
    torch.manual_seed(0)
    y = torch.rand(*shape)
    mask = 1
    if use_mask:
        mask = torch.zeros_like(y)
        mask[:, 0, 0::2, 0::2] = 1; mask[:, 1, 0::2, 1::2] = 1; mask[:, 1, 1::2, 0::2] = 1; mask[:, 2, 1::2, 1::2] = 1
        y = y * mask
    if kind == "gabor":
        A = [m.get_filter(transpose=True).detach() for m in net.A] if hasattr(net.A[0], "get_filter") else None
        A, B = net._filter_banks()
        A, B = [a.detach() for a in A], [b.detach() for b in B]
    else:
        A, B = [m.weight.detach() for m in net.A], [m.weight.detach() for m in net.B]
    t0 = time.time()
    xr, zr, *_ = O.forward_t(y, A, B, net.t.detach(), s, sigma, True, mask)
    d = dict(y=y.numpy(), A=np.stack([a.numpy() for a in A]), B=np.stack([b.numpy() for b in B]), t=net.t.detach().numpy(),
             sigma=sigma, adaptive=True)
    if use_mask:
        d["mask"] = mask.numpy()
    xe, ze = forward(d)
    print(json.dumps({"config": name, "samples": list(shape), "predicted_max_abs_xhat_vs_oracle": (xe - xr).abs().max().item(),
                      "predicted_max_abs_z_vs_oracle": (ze - zr).abs().max().item(), "z_nonzero_frac": (zr != 0).float().mean().item(),
                      "xhat_range": [xr.min().item(), xr.max().item()], "cpu_seconds": round(time.time() - t0, 1)}), flush=True)
