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


# ------------------------------------------------------------------------------------------------
# SURVEY.md 8(d) synthetic-weight protocol (no trained checkpoints exist)
# ------------------------------------------------------------------------------------------------
def bayer_mask(y):
    """RGGB mask of the reference (utils.py:13-19): R at (0,0), G at (0,1) and (1,0), B at (1,1)."""
    m = torch.zeros_like(y)
    m[:, 0, 0::2, 0::2] = 1
    m[:, 1, 0::2, 1::2] = 1
    m[:, 1, 1::2, 0::2] = 1
    m[:, 2, 1::2, 1::2] = 1
    return m


def protocol_net(kind, K, M, P, s, C, y, sigma_nominal, mask=1, seed=1, gain=1.0, order=1):
    """Module with the survey's synthetic weights: constructor init=True (the reference's power-method normalisation),
    every layer's banks perturbed by 1 + 0.03 randn (generator 7), thresholds from the 85th percentile q_m of |A_0 yp|:
    t[k,0,m] = 0.3 q_m u, t[k,1,m] = 0.7 q_m u / (sigma/255), u ~ U(0.8, 1.2).  `gain` scales the filters after the
    normalisation (gain 2 = the "hot" dictionary whose spectral constant exceeds 1)."""
    import contextlib
    import io
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        if kind == "gabor":
            net = cb.GDLNet(K=K, M=M, P=P, s=s, C=C, order=order, adaptive=True, init=True)
        elif kind == "video":
            net = cb.CDLNetVideo(K=K, M=M, P=P, s=s, C=C, adaptive=True, init=True, depth=8)
        else:
            net = cb.CDLNet(K=K, M=M, P=P, s=s, C=C, adaptive=True, init=True)
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for bank in list(net.A) + list(net.B):
            w = bank.alpha if kind == "gabor" else bank.weight
            w.mul_(gain * (1 + 0.03 * torch.randn(w.shape, generator=g)))
        yp = net._pre(y, mask)[0]
        a0 = net._analysis(0, yp).abs()
        q = torch.quantile(a0.transpose(0, 1).reshape(M, -1)[:, ::4], 0.85, dim=1)
        u = 0.8 + 0.4 * torch.rand(K, M, generator=g)
        shape = (K, M) + (1,) * (net.t.dim() - 3)
        net.t[:, 0] = (0.3 * q[None] * u).reshape(shape)
        net.t[:, 1] = (0.7 * q[None] * u / (sigma_nominal / 255.0)).reshape(shape)
    return net.eval()


def oracle_forward(net, y, sigma, mask=1):
    """The module's forward through the CPU oracle (oracle/cdl_oracle.py forward_t)."""
    import cdl_oracle as O
    A, B = net._filter_banks()
    with torch.no_grad():
        out = O.forward_t(y, [a.detach() for a in A], [b.detach() for b in B], net.t.detach(), net.s, sigma, net.adaptive, mask)
    return out[0], out[1]
