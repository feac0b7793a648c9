"""CPU oracle for the CDLNet ISTA forward path.  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU, the algorithm of the reference hot path
(`/root/reference/model/net.py`, `model/utils.py`, `model/gabor.py`).  It exists
so that the CUDA path in `cdlnet-video_b200/` can be checked against it.  Only
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` may import it; the product never routes through it.

Parity pin: the reference ships NO golden vectors, known-answer tests or
fixtures.  This oracle is therefore pinned against outputs of the
reference itself, produced in the build container by `oracle/gen_golden.py`
(which imports `/root/reference` read-only) and committed under `tests/golden/`.
`tests/test_oracle_golden.py` replays them on every run.

Two arithmetic back-ends are provided for every operator:
  * "numpy": an independent direct-form restatement (loops over taps, no library
    convolution), fp32 or fp64.  Slow; meant for small cases.
  * "torch": the same operator written with torch CPU functional ops.  The
    reference's arithmetic *is* PyTorch (an un-vendored dependency, unpinned in
    `requirements.txt:7`), so this back-end reproduces the reference's own CPU
    arithmetic; it is what the CPU baseline times.

Shapes follow the reference: images `(N,C,H,W)` / clips `(N,C,D,H,W)`, sparse
codes `(N,M,H/s,W/s)` / `(N,M,D/s,H/s,W/s)`, filters `(M,C,P,P)` /
`(M,C,Pd,Ph,Pw)`, thresholds `(K,2,M,1,1[,1])`.
"""
from __future__ import annotations

import math
import random as _random
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

try:  # torch is optional for the numpy back-end
    import torch
    import torch.nn.functional as F
except Exception:  # pragma: no cover
    torch = None
    F = None


# ----------------------------------------------------------------------------
# padding / index layout  (model/utils.py:35-51, 59-68, 100-122)
# ----------------------------------------------------------------------------
def calc_pad_1d(L: int, s: int) -> Tuple[int, int]:
    """model/utils.py:35-44 — pad L up to a multiple of s; the odd unit goes to the far side."""
    if L % s == 0:
        return (0, 0)
    diff = int(math.ceil(L / s)) * s - L
    return (diff // 2, diff - diff // 2)


def calc_pad_2d(H: int, W: int, s: int) -> Tuple[int, int, int, int]:
    """model/utils.py:46-51 — (left, right, top, bottom)."""
    return (*calc_pad_1d(W, s), *calc_pad_1d(H, s))


def calc_pad_3d(D: int, H: int, W: int, s: int) -> Tuple[int, int, int, int, int, int]:
    """model/utils.py:100-108 — (left, right, top, bottom, front, back)."""
    return (*calc_pad_1d(W, s), *calc_pad_1d(H, s), *calc_pad_1d(D, s))


def unpad_2d(x, pad):
    """model/utils.py:59-68 — correct for all four parity cases."""
    l, r, t, b = pad
    H, W = x.shape[-2], x.shape[-1]
    return x[..., t:H - b, l:W - r]


def unpad_3d_reference(x, pad):
    """model/utils.py:110-122, reproduced branch for branch INCLUDING its defect
    (SURVEY F10): it tests only `pad_back`/`pad_right` and slices `top:-bottom`
    even when bottom == 0."""
    l, r, t, b, f, k = pad
    if k == 0 and r > 0:
        return x[..., f:, t:-b, l:-r]          # b == 0 -> t:0, an empty axis
    elif k > 0 and r == 0:
        return x[..., f:-k, t:, l:]
    elif k == 0 and r == 0:
        return x[..., f:, t:, l:]
    else:
        return x[..., f:-k, t:-b, l:-r]


def unpad_3d(x, pad):
    """The evident crop (inverse of the reflect pad) for every parity class.
    Equals `unpad_3d_reference` whenever the latter returns the input shape."""
    l, r, t, b, f, k = pad
    D, H, W = x.shape[-3], x.shape[-2], x.shape[-1]
    return x[..., f:D - k, t:H - b, l:W - r]


def reflect_index(i: int, L: int) -> int:
    """Index map of F.pad(mode='reflect') (no edge repeat), for -L < i < 2L-1."""
    if i < 0:
        return -i
    if i >= L:
        return 2 * (L - 1) - i
    return i


# ----------------------------------------------------------------------------
# soft threshold  (model/net.py:11-14)
# ----------------------------------------------------------------------------
def soft_threshold_np(x: np.ndarray, t: np.ndarray) -> np.ndarray:
    """sign(x) * relu(|x| - t); defined for t < 0 as the reference is."""
    return np.sign(x) * np.maximum(np.abs(x) - t, 0)


# ----------------------------------------------------------------------------
# numpy direct-form operators
# ----------------------------------------------------------------------------
def _as3d_w(W: np.ndarray) -> np.ndarray:
    return W[:, :, None] if W.ndim == 4 else W


def analysis_np(r: np.ndarray, W: np.ndarray, s: int) -> np.ndarray:
    """nn.Conv{2,3}d(C,M,P,stride=s,padding=P//2,bias=False) (model/net.py:32,137-139):
    u[n,m,q] = sum_{c,t} W[m,c,t] * r[n,c, s*q - P//2 + t], zero outside."""
    two_d = r.ndim == 4
    if two_d:
        r = r[:, :, None]
    W3 = _as3d_w(W)
    N, C, D, H, Wd = r.shape
    M, _, Pd, Ph, Pw = W3.shape
    sd = 1 if two_d else s
    pd, ph, pw = Pd // 2, Ph // 2, Pw // 2
    Qd = (D + 2 * pd - Pd) // sd + 1
    Qh = (H + 2 * ph - Ph) // s + 1
    Qw = (Wd + 2 * pw - Pw) // s + 1
    rp = np.zeros((N, C, D + 2 * pd, H + 2 * ph, Wd + 2 * pw), dtype=r.dtype)
    rp[:, :, pd:pd + D, ph:ph + H, pw:pw + Wd] = r
    u = np.zeros((N, M, Qd, Qh, Qw), dtype=r.dtype)
    for c in range(C):
        for td in range(Pd):
            for th in range(Ph):
                for tw in range(Pw):
                    sl = rp[:, c, td:td + sd * (Qd - 1) + 1:sd,
                            th:th + s * (Qh - 1) + 1:s,
                            tw:tw + s * (Qw - 1) + 1:s]
                    u += W3[None, :, c, td, th, tw, None, None, None] * sl[:, None]
    return u[:, :, 0] if two_d else u


def synthesis_np(z: np.ndarray, W: np.ndarray, s: int) -> np.ndarray:
    """nn.ConvTranspose{2,3}d(M,C,P,stride=s,padding=P//2,output_padding=s-1,bias=False)
    (model/net.py:33,140-142); weight is (in=M, out=C, ...):
    x[n,c,v] = sum_{m,t : v = s*q - P//2 + t} W[m,c,t] * z[n,m,q]."""
    two_d = z.ndim == 4
    if two_d:
        z = z[:, :, None]
    W3 = _as3d_w(W)
    N, M, Qd, Qh, Qw = z.shape
    _, C, Pd, Ph, Pw = W3.shape
    sd = 1 if two_d else s
    pd, ph, pw = Pd // 2, Ph // 2, Pw // 2
    D = (Qd - 1) * sd - 2 * pd + Pd + (sd - 1)
    H = (Qh - 1) * s - 2 * ph + Ph + (s - 1)
    Wd = (Qw - 1) * s - 2 * pw + Pw + (s - 1)
    full = np.zeros((N, C, (Qd - 1) * sd + Pd + sd, (Qh - 1) * s + Ph + s, (Qw - 1) * s + Pw + s), dtype=z.dtype)
    for td in range(Pd):
        for th in range(Ph):
            for tw in range(Pw):
                # contribution of tap t from every coarse site, all channels at once
                contrib = np.einsum('nmdhw,mc->ncdhw', z, W3[:, :, td, th, tw])
                full[:, :, td:td + sd * (Qd - 1) + 1:sd,
                     th:th + s * (Qh - 1) + 1:s,
                     tw:tw + s * (Qw - 1) + 1:s] += contrib
    x = full[:, :, pd:pd + D, ph:ph + H, pw:pw + Wd]
    return x[:, :, 0] if two_d else x


def pre_process_np(y: np.ndarray, s: int, mask=1):
    """model/utils.py:5-22 (2D) and :70-87 (3D): per-sample (masked) mean, centre,
    mask, reflect-pad image and mask up to a multiple of s."""
    axes = tuple(range(1, y.ndim))
    has_mask = isinstance(mask, np.ndarray)
    if has_mask:
        mean = y.sum(axis=axes, keepdims=True, dtype=y.dtype) / mask.sum(axis=axes, keepdims=True, dtype=y.dtype)
    else:
        mean = y.mean(axis=axes, keepdims=True, dtype=y.dtype)
    x = mask * (y - mean)
    if y.ndim == 4:
        pad = calc_pad_2d(y.shape[2], y.shape[3], s)
        widths = ((0, 0), (0, 0), (pad[2], pad[3]), (pad[0], pad[1]))
    else:
        pad = calc_pad_3d(y.shape[2], y.shape[3], y.shape[4], s)
        widths = ((0, 0), (0, 0), (pad[4], pad[5]), (pad[2], pad[3]), (pad[0], pad[1]))
    yp = np.pad(x, widths, mode='reflect')
    mp = np.pad(mask, widths, mode='reflect') if has_mask else mask
    return yp.astype(y.dtype), mean, pad, mp


def threshold_np(t: np.ndarray, k: int, c, N: int, ndim: int) -> np.ndarray:
    """model/net.py:85,87 — tau_k = t[k,0] + c * t[k,1]; c scalar or per-sample."""
    M = t.shape[2]
    t0 = t[k, 0].reshape(1, M, *([1] * (ndim - 2)))
    t1 = t[k, 1].reshape(1, M, *([1] * (ndim - 2)))
    if isinstance(c, np.ndarray):
        c = c.reshape(N, *([1] * (ndim - 1)))
    return t0 + c * t1


def forward_np(y: np.ndarray, A: Sequence[np.ndarray], B: Sequence[np.ndarray], t: np.ndarray,
               s: int, sigma=None, adaptive=True, mask=1, fix_unpad=True, trace: Optional[List] = None):
    """CDLNet.forward / CDLNetVideo.forward / GDLNet.forward (model/net.py:76-92,
    192-212, 659-675) with residual=False.  A[k], B[k] are filter arrays; for
    GDLNet pass the synthesised Gabor filters.  Returns (xhat, z, yp, mean, pad)."""
    K = len(A)
    yp, mean, pad, mp = pre_process_np(y, s, mask)
    N = y.shape[0]
    if sigma is None or not adaptive:
        c = 0.0
    elif isinstance(sigma, np.ndarray):
        c = (sigma.reshape(-1) / 255.0).astype(y.dtype)
    else:
        c = y.dtype.type(sigma / 255.0)
    z = soft_threshold_np(analysis_np(yp, A[0], s), threshold_np(t, 0, c, N, y.ndim))
    if trace is not None:
        trace.append(z.copy())
    for k in range(1, K):
        r = mp * synthesis_np(z, B[k], s) - yp
        z = soft_threshold_np(z - analysis_np(r, A[k], s), threshold_np(t, k, c, N, y.ndim))
        if trace is not None:
            trace.append(z.copy())
    xp = synthesis_np(z, B[0], s)
    if y.ndim == 4:
        xhat = unpad_2d(xp, pad) + mean
    else:
        xhat = (unpad_3d(xp, pad) if fix_unpad else unpad_3d_reference(xp, pad)) + mean
    return xhat, z, yp, mean, pad


# ----------------------------------------------------------------------------
# Gabor filter synthesis  (model/gabor.py:7-28, 46-51)
# ----------------------------------------------------------------------------
def gabor_filter_np(alpha: np.ndarray, a: np.ndarray, w0: np.ndarray, psi: np.ndarray, ks: int) -> np.ndarray:
    """(order,M,C,1,1),(order,M,C,2),(order,M,C,2),(order,M,C) -> (M,C,ks,ks):
    sum_order alpha * exp(-|a*(x-x0)|^2) * cos(w0.(x-x0) + psi), x on an ij meshgrid."""
    i = np.arange(ks, dtype=a.dtype)
    gi, gj = np.meshgrid(i, i, indexing='ij')
    x = np.stack([gi, gj], axis=-1) - a.dtype.type((ks - 1) / 2)         # (ks,ks,2)
    ax = a[:, :, :, None, None, :] * x[None, None, None]                 # (o,M,C,ks,ks,2)
    env = np.exp(-np.sum(ax * ax, axis=-1))
    ph = np.sum(w0[:, :, :, None, None, :] * x[None, None, None], axis=-1) + psi[:, :, :, None, None]
    return (alpha * env * np.cos(ph)).sum(axis=0)


# ----------------------------------------------------------------------------
# torch CPU back-end: the reference's own arithmetic library
# ----------------------------------------------------------------------------
def analysis_t(r, W, s):
    """Same operator as `analysis_np`, through torch's CPU convolution."""
    if r.dim() == 4:
        return F.conv2d(r, W, stride=s, padding=(W.shape[2] // 2, W.shape[3] // 2))
    return F.conv3d(r, W, stride=s, padding=(W.shape[2] // 2, W.shape[3] // 2, W.shape[4] // 2))


def synthesis_t(z, W, s):
    """Same operator as `synthesis_np`, through torch's CPU transposed convolution."""
    if z.dim() == 4:
        return F.conv_transpose2d(z, W, stride=s, padding=(W.shape[2] // 2, W.shape[3] // 2), output_padding=s - 1)
    return F.conv_transpose3d(z, W, stride=s, padding=(W.shape[2] // 2, W.shape[3] // 2, W.shape[4] // 2),
                              output_padding=s - 1)


def soft_threshold_t(x, t):
    return x.sign() * F.relu(x.abs() - t)


def pre_process_t(y, s, mask=1):
    dims = tuple(range(1, y.dim()))
    has_mask = torch.is_tensor(mask)
    if has_mask:
        mean = y.sum(dim=dims, keepdim=True) / mask.sum(dim=dims, keepdim=True)
    else:
        mean = y.mean(dim=dims, keepdim=True)
    x = mask * (y - mean)
    pad = calc_pad_2d(*y.shape[2:], s) if y.dim() == 4 else calc_pad_3d(*y.shape[2:], s)
    yp = F.pad(x, pad, mode='reflect')
    mp = F.pad(mask, pad, mode='reflect') if has_mask else mask
    return yp, mean, pad, mp


def forward_t(y, A, B, t, s, sigma=None, adaptive=True, mask=1, fix_unpad=True, trace=None):
    """Torch-CPU twin of `forward_np`.  `t` is (K,2,M,1,1[,1]); sigma None | number |
    tensor broadcastable to (N,1,...)."""
    K = len(A)
    yp, mean, pad, mp = pre_process_t(y, s, mask)
    c = 0 if sigma is None or not adaptive else sigma / 255.0
    z = soft_threshold_t(analysis_t(yp, A[0], s), t[0, :1] + c * t[0, 1:2])
    if trace is not None:
        trace.append(z.clone())
    for k in range(1, K):
        z = soft_threshold_t(z - analysis_t(mp * synthesis_t(z, B[k], s) - yp, A[k], s), t[k, :1] + c * t[k, 1:2])
        if trace is not None:
            trace.append(z.clone())
    xp = synthesis_t(z, B[0], s)
    if y.dim() == 4:
        xhat = unpad_2d(xp, pad) + mean
    else:
        xhat = (unpad_3d(xp, pad) if fix_unpad else unpad_3d_reference(xp, pad)) + mean
    return xhat, z, yp, mean, pad


def gabor_filter_t(alpha, a, w0, psi, ks):
    return torch.from_numpy(gabor_filter_np(alpha.detach().numpy(), a.detach().numpy(),
                                            w0.detach().numpy(), psi.detach().numpy(), ks))


# ----------------------------------------------------------------------------
# synthetic inputs: the reference's own definitions, restated
# ----------------------------------------------------------------------------
def syn_clip(H: int, W: int, D: int, seed: int = 0) -> np.ndarray:
    """syn_data/gen.py:12-31 + the min-max normalisation of syn_data/gen_data_draft.py:33-38.
    Random signed sum of 2-10 sin/cos plane waves on [-pi,pi]^3 -> (D,H,W) float32 in [0,1]."""
    rng = _random.Random(seed)
    x = np.linspace(-np.pi, np.pi, H)
    yv = np.linspace(-np.pi, np.pi, W)
    zv = np.linspace(-np.pi, np.pi, D)
    X, Y, Z = np.meshgrid(x, yv, zv, indexing='ij')
    terms = []
    for _ in range(rng.randint(2, 10)):
        cx, cy, cz = rng.uniform(-5, 5), rng.uniform(-5, 5), rng.uniform(-5, 5)
        f = rng.choice([np.sin, np.cos])
        terms.append(f(cx * X + cy * Y + cz * Z))
    out = terms[0]
    for term in terms[1:]:
        out = out + term if rng.choice([True, False]) else out - term
    lo, hi = out.min(), out.max()
    out = (out - lo) / (hi - lo) if hi > lo else np.zeros_like(out)
    return np.ascontiguousarray(np.transpose(out, (2, 0, 1))).astype(np.float32)


def bayer_mask_np(shape) -> np.ndarray:
    """utils.py:13-19 — RGGB: R (0,0), G (0,1)&(1,0), B (1,1); shape (N,3,H,W)."""
    m = np.zeros(shape, dtype=np.float32)
    m[:, 0, 0::2, 0::2] = 1
    m[:, 1, 0::2, 1::2] = 1
    m[:, 1, 1::2, 0::2] = 1
    m[:, 2, 1::2, 1::2] = 1
    return m


def synthetic_weights(K: int, M: int, C: int, P: Sequence[int], s: int, yp: "torch.Tensor",
                      sigma: float, seed: int = 1, L: Optional[float] = None, power_iters: int = 30):
    """SURVEY.md §8(d) synthetic-weight protocol (the trained checkpoints are missing,
    SURVEY F6).  One randn filter bank normalised by a power-method estimate of the
    spectral constant of D∘A (model/net.py:37-58, model/solvers.py:3-22; fewer
    iterations than the reference's 200 — only magnitudes matter here), perturbed
    per layer by 1+0.03*randn, with per-subband thresholds from the 85th percentile
    of |A_0 yp|.  Returns (A list, B list, t) as torch tensors."""
    g = torch.Generator().manual_seed(seed)
    nd = len(P)
    W = torch.randn(M, C, *P, generator=g)
    if L is None:
        b = torch.rand(1, C, *[max(4 * p, 32) for p in P], generator=g)
        for _ in range(power_iters):
            b = synthesis_t(analysis_t(b, W, s), W, s)
            b = b / b.norm()
        L = float((b * synthesis_t(analysis_t(b, W, s), W, s)).sum())
    W = W / math.sqrt(L)
    g2 = torch.Generator().manual_seed(7)
    A = [W * (1 + 0.03 * torch.randn(W.shape, generator=g2)) for _ in range(K)]
    B = [W * (1 + 0.03 * torch.randn(W.shape, generator=g2)) for _ in range(K)]
    u0 = analysis_t(yp, A[0], s).abs()
    q = torch.quantile(u0.transpose(0, 1).reshape(M, -1)[:, :200000], 0.85, dim=1)
    u = 0.8 + 0.4 * torch.rand(K, M, generator=g2)
    t = torch.zeros(K, 2, M, *([1] * nd))
    t[:, 0] = (0.3 * q[None] * u).reshape(K, M, *([1] * nd))
    t[:, 1] = (0.7 * q[None] * u / max(sigma / 255.0, 1e-6)).reshape(K, M, *([1] * nd))
    return A, B, t, L


def psnr(x, ref) -> float:
    """-10 log10 mean((x-ref)^2), images in [0,1] (analyze.py PSNR convention)."""
    x = np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(-10.0 * np.log10(np.mean((x - ref) ** 2)))


# ------------------------------------------------------------------------------------------------
# blind noise level (reference model/nle.py:17-27 nle_mad, model/wvlt.py:5-42) - SURVEY.md 8f N3
# ------------------------------------------------------------------------------------------------
# pywt.Wavelet('bior4.4').dec_hi: the reference reads it from PyWavelets at run time (model/wvlt.py:9; the dependency is
# unpinned and absent from this image).  Values = sqrt(2) x the published CDF 9/7 7-tap high-pass, zero-padded to the
# 10-tap length pywt uses for bior4.4.  PARITY UNPINNED at the level of these ten numbers (no pywt here to read them
# from); everything downstream of them is pinned by tests/golden/nle_mad.npz, generated by the reference's own
# model/nle.py + model/wvlt.py with a stub `pywt` that serves this table (oracle/gen_golden.py nle).
BIOR44_DEC_HI = (0.0, -0.06453888262869706, 0.04068941760916406, 0.41809227322161724, -0.7884856164055829,
                 0.41809227322161724, 0.04068941760916406, -0.06453888262869706, 0.0, 0.0)
BIOR44_DEC_LO = (0.0, 0.03782845550726404, -0.023849465019556843, -0.11062440441843718, 0.37740285561283066,
                 0.8526986790088938, 0.37740285561283066, -0.11062440441843718, -0.023849465019556843, 0.03782845550726404)
BIOR44_REC_LO = (0.0, -0.06453888262869706, -0.04068941760916406, 0.41809227322161724, 0.7884856164055829,
                 0.41809227322161724, -0.04068941760916406, -0.06453888262869706, 0.0, 0.0)
BIOR44_REC_HI = (0.0, -0.03782845550726404, -0.023849465019556843, 0.11062440441843718, 0.37740285561283066,
                 -0.8526986790088938, 0.37740285561283066, 0.11062440441843718, -0.023849465019556843, -0.03782845550726404)


def nle_mad_np(y: np.ndarray) -> np.ndarray:
    """model/nle.py:17-27 restated in numpy: hh[a][b] = dec_hi[9-a] * dec_hi[9-b] (model/wvlt.py:34-42: outer product,
    both axes flipped; analysis bank index 3 = high/high), cross-correlation with stride 2 and no padding per channel,
    LOWER median of the magnitudes per sample (torch.median), / 0.6745.  y (N,C,H,W) -> (N,) float32."""
    y = np.asarray(y, dtype=np.float32)
    N, C, H, W = y.shape
    g = np.asarray(BIOR44_DEC_HI, dtype=np.float32)[::-1].copy()
    hh = (g[:, None] * g[None, :]).astype(np.float32)
    Ho, Wo = (H - 10) // 2 + 1, (W - 10) // 2 + 1
    acc = np.zeros((N, C, Ho, Wo), dtype=np.float32)
    for a in range(10):
        for b in range(10):
            if hh[a, b] != 0.0:
                acc += y[:, :, a:a + 2 * Ho - 1:2, b:b + 2 * Wo - 1:2] * hh[a, b]
    mag = np.abs(acc).reshape(N, -1)
    k = (mag.shape[1] - 1) // 2
    med = np.partition(mag, k, axis=1)[:, k]
    return (med / np.float32(0.6745)).astype(np.float32)


# ------------------------------------------------------------------------------------------------
# frame-recurrent CSR variants (reference model/net.py:229-262 prox_CSR / prox_CSR_f2; forward loops :426-462, :525-567)
# ------------------------------------------------------------------------------------------------
def prox_csr_t(u, z_prev, lambd, gamma):
    """model/net.py:229-242"""
    return soft_threshold_t(soft_threshold_t(u - z_prev - lambd * torch.sign(z_prev), lambd * gamma) + z_prev + lambd * torch.sign(z_prev), lambd)


def prox_csr_f2_t(u, z_prev, z_after, lambd, gamma1, gamma2):
    """model/net.py:244-262"""
    Ca = z_prev + lambd * torch.sign(z_prev) + lambd * gamma2 * torch.sign(z_prev - z_after)
    Cb = z_after + lambd * torch.sign(z_after) + lambd * gamma1 * torch.sign(z_after - z_prev)
    inner = soft_threshold_t(u - Ca, gamma1 * lambd)
    midder = soft_threshold_t(inner - Cb + lambd * gamma1 * torch.sign(u - Ca), gamma2 * lambd)
    return soft_threshold_t(midder + Cb - lambd * gamma1 * torch.sign(u - Ca), lambd)


def forward_csr_t(y, A, B, t, s, D, sigma=None, adaptive=True, mask=1, z_prev=None, z_after=None, g_prev=None, g_after=None):
    """The CSR forward loops restated on the torch-CPU operators of this oracle.  A, B, t: the operator set the
    iterations use (CDLNet_CSR without z_prev: A2, B2, t2); D: the dictionary of the final synthesis (always B[0] of the
    first set); g_prev / g_after: the (K,2,M,1,1) gamma parameters paired with z_prev / z_after (CDLNet_CSR: g;
    CDLNet_CSRf2: g1 / g2).  Returns (xhat, z)."""
    K = len(A)
    yp, mean, pad, mp = pre_process_t(y, s, mask)
    c = 0 if sigma is None or not adaptive else sigma / 255.0
    thr = lambda p, k: p[k, :1] + c * p[k, 1:2]

    def prox(u, k):
        if z_prev is not None and z_after is not None:
            return prox_csr_f2_t(u, z_prev, z_after, thr(t, k), thr(g_prev, k), thr(g_after, k))
        if z_prev is not None:
            return prox_csr_t(u, z_prev, thr(t, k), thr(g_prev, k))
        if z_after is not None:
            return prox_csr_t(u, z_after, thr(t, k), thr(g_after, k))
        return soft_threshold_t(u, thr(t, k))
    z = prox(analysis_t(yp, A[0], s), 0)
    for k in range(1, K):
        z = prox(z - analysis_t(mp * synthesis_t(z, B[k], s) - yp, A[k], s), k)
    xp = synthesis_t(z, D, s)
    return unpad_2d(xp, pad) + mean, z
