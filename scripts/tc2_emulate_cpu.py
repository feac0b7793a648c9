#!/usr/bin/env python
"""CPU emulation of the 2-D tensor-core family's arithmetic on the golden fixtures: the analysis rounds r and A_k to tf32
(round-to-nearest, ties away), the residual synthesis rounds z and B_k, accumulation in fp32, the final D z exact.
Predicts max|xhat - reference| of the GPU path without a GPU (tests/test_zz_golden_tc2_gpu.py asserts <= 1e-4)."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import cdl_oracle as O  # noqa: E402


def tf32(x):
    b = x.contiguous().view(torch.int32)
    return ((b + 0x1000) & ~0x1fff).view(torch.float32)


def forward(d, analysis="tf32"):
    """analysis: "tf32" = the default kernels; "3term" = the hi/lo analysis of cdl_tc2_analysis_x3.cuh (precision "tf32x3")."""
    y = torch.from_numpy(d["y"])
    A = [torch.from_numpy(a) for a in d["A"]]
    B = [torch.from_numpy(b) for b in d["B"]]
    t = torch.from_numpy(d["t"])
    mask = torch.from_numpy(d["mask"]) if "mask" in d else 1
    sigma = d["sigma"]
    if isinstance(sigma, np.ndarray):
        sigma = torch.from_numpy(sigma)
    yp, mean, pad, mp = O.pre_process_t(y, 1, mask)
    c = 0 if sigma is None or not d["adaptive"] else sigma / 255.0
    def ana(r, w):
        rh, wh = tf32(r), tf32(w)
        u = F.conv2d(rh, wh, padding=3)
        if analysis == "3term":
            u = u + F.conv2d(tf32(r - rh), wh, padding=3) + F.conv2d(rh, tf32(w - wh), padding=3)
        return u
    syn = lambda z, w: F.conv_transpose2d(tf32(z), tf32(w), padding=3)
    z = O.soft_threshold_t(ana(yp, A[0]), t[0, :1] + c * t[0, 1:2])
    for k in range(1, len(A)):
        z = O.soft_threshold_t(z - ana(mp * syn(z, B[k]) - yp, A[k]), t[k, :1] + c * t[k, 1:2])
    xp = F.conv_transpose2d(z, B[0], padding=3)
    return xp + mean, z


if __name__ == "__main__":
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import load_case
    for name in sys.argv[1:] or ["cdlnet2d_nonadaptive", "cdlnet2d_jdd_s1_w4", "gdlnet_s1_c3"]:
        d = load_case(name)
        xhat, z = forward(d)
        x3, _ = forward(d, analysis="3term")
        print(f"{name}: predicted max|xhat - reference| = {np.abs(xhat.numpy() - d['xhat']).max():.3e}   max|z - reference| = {np.abs(z.numpy() - d['z']).max():.3e}"
              f"   with the 3-term analysis: {np.abs(x3.numpy() - d['xhat']).max():.3e}")
