"""Shared helpers for the parity tests: golden-fixture loading and module construction."""
import os

import numpy as np
import torch

import cdlnet_video_b200 as cb

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["cdlnet2d_s2", "cdlnet2d_jdd_mask", "video_s2_p777", "video_s2_p995_odd",
         "video_s1_p775_c2", "gdlnet_s2_c3", "cdlnet2d_nonadaptive"]


def load_case(name):
    d = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    d["s"] = int(d["s"])
    d["adaptive"] = bool(int(d["adaptive"]))
    if "sigma_none" in d:
        d["sigma"] = None
    elif d["sigma"].ndim == 0:
        d["sigma"] = float(d["sigma"])
    return d


def module_from_case(d, name):
    """Build the drop-in module with the fixture's weights (state-dict route, like train.load_ckpt)."""
    A, B, t = d["A"], d["B"], d["t"]
    K, M, C = A.shape[0], A.shape[1], A.shape[2]
    P = A.shape[3:]
    if name.startswith("gdlnet"):
        order = d["A_alpha"].shape[1]
        net = cb.GDLNet(K=K, M=M, P=int(P[0]), s=d["s"], C=C, order=order, adaptive=d["adaptive"], init=False)
        sd = {"t": torch.from_numpy(t)}
        for k in range(K):
            for side in ("A", "B"):
                for nm in ("alpha", "a", "w0", "psi"):
                    sd[f"{side}.{k}.{nm}"] = torch.from_numpy(d[f"{side}_{nm}"][k])
        for nm in ("alpha", "a", "w0", "psi"):
            sd[f"D.{nm}"] = sd[f"B.0.{nm}"]
        net.load_state_dict(sd)
    else:
        if A.ndim == 6:
            net = cb.CDLNetVideo(K=K, M=M, P=list(P), s=d["s"], C=C, adaptive=d["adaptive"], init=False)
        else:
            net = cb.CDLNet(K=K, M=M, P=int(P[0]), s=d["s"], C=C, adaptive=d["adaptive"], init=False)
        sd = {"t": torch.from_numpy(t)}
        for k in range(K):
            sd[f"A.{k}.weight"] = torch.from_numpy(A[k])
            sd[f"B.{k}.weight"] = torch.from_numpy(B[k])
        sd["D.weight"] = sd["B.0.weight"]
        net.load_state_dict(sd)          # CDLNet: no 'g' key, like upstream checkpoints
    return net.eval()


def case_inputs(d, device="cpu"):
    y = torch.from_numpy(d["y"]).to(device)
    sigma = d["sigma"]
    if isinstance(sigma, np.ndarray):
        sigma = torch.from_numpy(sigma).to(device)
    mask = torch.from_numpy(d["mask"]).to(device) if "mask" in d else 1
    return y, sigma, mask
