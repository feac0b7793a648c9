"""Drop-in CDLNet / CDLNetVideo / GDLNet modules (reference model/net.py:16-104, 121-227, 569-687).

Constructors, attributes (`A`, `B`, `D`, `t`, `K`, `M`, `P`, `s`, `t0`, `adaptive`), state-dict keys,
`forward(y, sigma=None, mask=1) -> (xhat, z)`, `forward_generator` and `project()` are the
reference's.  What changes is the body of `forward`: under `torch.no_grad()` with fp32 CUDA inputs the
whole pass (mean/pad preprocess, K ISTA iterations, D z, crop) runs in libcdl_b200's hand-written
sm_100a kernels through the C ABI of include/cdl_b200.h.  That route has NO fallback: if the shared
library is missing or a kernel fails, `forward` raises.  With autograd enabled (training), or on CPU
tensors, the modules evaluate the same expression with their own `nn.Conv*` layers so that gradients
reach `A[k].weight`, `B[k].weight` and `t` exactly as in the reference (train.py:83-102).
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .solvers import power_method, uball_project
from .utils import pre_process, post_process, pre_process_3d, post_process_3d
from .gabor import ConvAdjoint2dGabor

try:                                     # imported as cdlnet_video_b200.model.net
    from ..plan import Plan
except ImportError:                      # imported as top-level `model.net` (drop-in inside the reference tree)
    import importlib.util as _ilu
    _pkg_dir = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if "cdlnet_video_b200" not in sys.modules:
        _spec = _ilu.spec_from_file_location("cdlnet_video_b200", os.path.join(_pkg_dir, "__init__.py"),
                                             submodule_search_locations=[_pkg_dir])
        _mod = _ilu.module_from_spec(_spec)
        sys.modules["cdlnet_video_b200"] = _mod
        _spec.loader.exec_module(_mod)
    from cdlnet_video_b200.plan import Plan


def ST(x, t):
    """shrinkage-thresholding: sign(x) * relu(|x| - t)"""
    return x.sign() * F.relu(x.abs() - t)


def _spectral_constant(op, shape):
    print("Running power-method on initial dictionary...")
    with torch.no_grad():
        L = power_method(op, torch.rand(*shape), num_iter=200, verbose=False)[0]
    print(f"Done. L={L:.3e}.")
    if L < 0:
        print("STOP: something is very very wrong...")
        sys.exit()
    return L


class _ISTANet(nn.Module):
    """Shared forward machinery.  Subclasses define `_nsp` (2 or 3 spatial axes), `_analysis(k, x)`,
    `_synthesis(k, z)` (stock torch route) and `_filter_banks()` (tensors handed to the library)."""
    _nsp = 2
    #: "auto" picks the fastest kernel family that meets the parity bar for the geometry;
    #: "fp32" forces the exact CUDA-core kernels, "tf32" requests the tcgen05 path.
    precision = os.environ.get("CDL_PRECISION", "auto")

    # -- plumbing ----------------------------------------------------------------------------------
    def __getstate__(self):
        state = self.__dict__.copy()
        state.pop("_plans", None)        # ctypes handles are per process
        return state

    def _P3(self):
        P = self.P
        n = self._nsp
        return tuple(int(p) for p in P) if isinstance(P, (tuple, list)) else (int(P),) * n

    def _native_ok(self, y, sigma, mask):
        if torch.is_grad_enabled() and (y.requires_grad or any(p.requires_grad for p in self.parameters())):
            return False
        if not (torch.is_tensor(y) and y.is_cuda and y.dtype == torch.float32 and y.dim() == self._nsp + 2):
            return False
        if getattr(self, "residual", False):
            return False                 # ResidualBlock variant: out of scope, stock route
        if torch.is_tensor(mask) and tuple(torch.broadcast_shapes(mask.shape, y.shape)) != tuple(y.shape):
            return False
        if not torch.is_tensor(mask) and mask != 1:
            return False
        if torch.is_tensor(sigma) and sigma.numel() not in (1, y.shape[0]):
            return False                 # per-pixel sigma maps are not part of the reference's callers
        return True

    def _plan_for(self, y, has_mask):
        plans = self.__dict__.setdefault("_plans", {})
        prec = "fp32" if self.precision in ("fp32",) else ("tf32" if self.precision in ("tf32", "auto") else "fp32")
        key = (tuple(y.shape), has_mask, y.device.index, prec)
        plan = plans.get(key)
        if plan is None:
            if len(plans) >= 8:
                plans.pop(next(iter(plans))).close()
            plan = Plan(self._nsp, y.shape[0], y.shape[1], self.M, self.K, tuple(y.shape[2:]), self._P3(), self.s,
                        has_mask=has_mask, precision=prec, device=y.device.index or 0)
            plans[key] = plan
        return plan

    def _weights_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _c_vector(self, sigma, N, device):
        """c = sigma/255 (model/net.py:82,197) as an fp32 vector of N, rounded like the reference:
        a python number is divided in double then cast; a tensor is divided in its own dtype."""
        if sigma is None or not self.adaptive:
            return None
        if torch.is_tensor(sigma):
            c = (sigma / 255.0).to(device=device, dtype=torch.float32).reshape(-1)
            return (c.expand(N) if c.numel() == 1 else c).contiguous()
        return torch.full((N,), float(sigma) / 255.0, dtype=torch.float32, device=device)

    def _prepare(self, y, sigma, mask):
        has_mask = torch.is_tensor(mask)
        y = y.contiguous()
        if has_mask:
            mask = mask.to(device=y.device, dtype=torch.float32).expand_as(y).contiguous()
        else:
            mask = None
        with torch.cuda.device(y.device):
            plan = self._plan_for(y, has_mask)
            A, B = self._filter_banks()
            plan.set_weights(A, B, self.t, key=self._weights_key())
        return plan, y, mask, self._c_vector(sigma, y.shape[0], y.device)

    # -- the hot path ------------------------------------------------------------------------------
    def forward(self, y, sigma=None, mask=1):
        """ LISTA + D w/ noise-adaptive thresholds """
        if not self._native_ok(y, sigma, mask):
            return self._forward_stock(y, sigma, mask)
        if self._embed3d_ok(y, mask):
            return self._forward_embedded3d(y, sigma)
        plan, y, mask, c = self._prepare(y, sigma, mask)
        with torch.cuda.device(y.device):
            return plan.denoise(y, mask, c)

    # -- opt-in: 2-D stride-2 grayscale nets (CDLNet-s2030, BASELINE config 1) on the VIDEO tensor-core kernels ------
    # A 2-D image is a clip of two frames (image, zero) and a 7x7 filter is the td = 3 slice of a 7x7x7 filter: with
    # stride 2 and padding 3 along d the only coarse frame reads fine frames td - 3 = 0 (td = 3) and 1 (td = 4, a zero
    # slice), and the synthesis writes frame 0 from td = 3 and nothing into frame 1 - the 3-D operator restricted to
    # frame 0 IS the 2-D operator.  6/7 of the MMAs multiply zeros, but the problem (one 256x256 image) is launch-bound
    # either way.  NOT yet run on hardware: enabled only by CDL_EMBED3D=1 together with precision "tf32"/"auto".
    def _embed3d_ok(self, y, mask):
        if os.environ.get("CDL_EMBED3D", "0") != "1" or self.precision not in ("tf32", "auto"):
            return False
        if self._nsp != 2 or type(self).__name__ != "CDLNet" or torch.is_tensor(mask):
            return False
        if not (self.s == 2 and self._P3() == (7, 7) and y.shape[1] == 1 and self.M <= 176):
            return False
        return (-(-y.shape[3] // 2) * 2) % 4 == 0            # the video kernels need a padded width that is a multiple of 4

    def _forward_embedded3d(self, y, sigma):
        y = y.contiguous()
        N, dev = y.shape[0], y.device
        plans = self.__dict__.setdefault("_plans", {})
        with torch.cuda.device(dev):
            k2 = (tuple(y.shape), False, dev.index, "fp32")
            p2 = plans.get(k2)
            if p2 is None:                                   # pre/post-processing and the index layout: the 2-D plan
                p2 = plans[k2] = Plan(2, N, 1, self.M, self.K, tuple(y.shape[2:]), (7, 7), 2, precision="fp32", device=dev.index or 0)
            k3 = (tuple(y.shape), "embed3d", dev.index, "tf32")
            p3 = plans.get(k3)
            if p3 is None:                                   # two frames of the PADDED image: no further stride padding
                p3 = plans[k3] = Plan(3, N, 1, self.M, self.K, (2, *p2.fine[1:]), (7, 7, 7), 2, precision="tf32", device=dev.index or 0)
            if p3.precision != "tf32":
                raise RuntimeError("CDL_EMBED3D=1: the video tensor-core kernels do not cover this geometry")
            key = self._weights_key()
            if p3._weights_key != key:
                def lift(w):                                 # (M,1,7,7) -> (M,1,7,7,7), the filter sits in the td = 3 slice
                    w3 = torch.zeros(w.shape[0], 1, 7, 7, 7, dtype=torch.float32, device=dev)
                    w3[:, :, 3] = w.detach().to(dev, torch.float32)
                    return w3
                p3.set_weights([lift(m.weight) for m in self.A], [lift(m.weight) for m in self.B], self.t, key=key)
            yp, _, mean = p2.preprocess(y)
            yp3 = torch.zeros(N, 1, 2, *yp.shape[2:], dtype=torch.float32, device=dev)
            yp3[:, :, 0] = yp
            z3, xp3 = p3.forward(yp3, None, self._c_vector(sigma, N, dev))
            xhat = p2.postprocess(xp3[:, :, 0].contiguous(), mean)
            return xhat, z3[:, :, 0]

    def forward_generator(self, y, sigma=None, mask=1):
        """ same as forward but yields intermediate sparse codes z_0..z_{K-1}, then xhat """
        if not self._native_ok(y, sigma, mask):
            yield from self._generator_stock(y, sigma, mask)
            return
        plan, y, mask, c = self._prepare(y, sigma, mask)
        with torch.cuda.device(y.device):
            yp, mp, mean = plan.preprocess(y, mask)
            z = plan.new_code()                  # the code stays in the plan's internal layout between steps
            r = torch.empty_like(yp)
            plan.analysis_step(0, yp, z, c, first=True)
            yield plan.export_code(z)
            for k in range(1, self.K):
                plan.synthesis_step(k, z, r, yp, mp, residual=True)
                plan.analysis_step(k, r, z, c)
                yield plan.export_code(z)
            plan.synthesis_step(0, z, r, residual=False)
            yield plan.postprocess(r, mean)

    # -- stock torch route (autograd / CPU) ------------------------------------------------------
    def _pre(self, y, mask):
        return (pre_process if self._nsp == 2 else pre_process_3d)(y, self.s, mask=mask)

    def _post(self, x, params):
        return (post_process if self._nsp == 2 else post_process_3d)(x, params)

    def _tau(self, k, c):
        return self.t[k, :1] + c * self.t[k, 1:2]

    def _generator_stock(self, y, sigma, mask):
        yp, params, mask = self._pre(y, mask)
        c = 0 if sigma is None or not self.adaptive else sigma / 255.0
        z = ST(self._analysis(0, yp), self._tau(0, c))
        z = self._after(0, z)
        yield z
        for k in range(1, self.K):
            z = ST(z - self._analysis(k, mask * self._synthesis(k, z) - yp), self._tau(k, c))
            z = self._after(k, z)
            yield z
        yield self._post(self.D(z), params)

    def _forward_stock(self, y, sigma, mask):
        z = None
        for item in self._generator_stock(y, sigma, mask):
            z, prev = item, z
        return z, prev                   # last item is xhat, the one before is z_{K-1}

    def _after(self, k, z):
        return z

    def _analysis(self, k, x):
        return self.A[k](x)

    def _synthesis(self, k, z):
        return self.B[k](z)

    def _filter_banks(self):
        return [m.weight for m in self.A], [m.weight for m in self.B]


class CDLNet(_ISTANet):
    """ Convolutional Dictionary Learning Network:
    Interpretable denoising DNN with adaptive thresholds for robustness.
    (Called with a Bayer `mask` tensor and C=3 it is the joint demosaic+denoise model)
    """
    _nsp = 2

    def __init__(self,
                 K=3,              # num. unrollings
                 M=64,             # num. filters in each filter bank operation
                 P=7,              # square filter side length
                 s=1,              # stride of convolutions
                 C=1,              # num. input channels
                 t0=0,             # initial threshold
                 adaptive=False,   # noise-adaptive thresholds
                 init=True):       # False -> skip the power-method weight init (loading a state-dict)
        super().__init__()
        pad = (P - 1) // 2
        self.A = nn.ModuleList([nn.Conv2d(C, M, P, stride=s, padding=pad, bias=False) for _ in range(K)])
        self.B = nn.ModuleList([nn.ConvTranspose2d(M, C, P, stride=s, padding=pad, output_padding=s - 1, bias=False)
                                for _ in range(K)])
        self.D = self.B[0]                                        # alias; keeps the 'D.weight' state-dict key
        self.t = nn.Parameter(t0 * torch.ones(K, 2, M, 1, 1))     # (layer, [t_0, t_1*sigma], subband, 1, 1)
        self.g = nn.Parameter(t0 * torch.ones(K, 2, M, 1, 1))     # present in the reference, unused in forward
        W = torch.randn(M, C, P, P)
        for k in range(K):
            self.A[k].weight.data = W.clone()
            self.B[k].weight.data = W.clone()
        if init:
            L = _spectral_constant(lambda x: self.D(self.A[0](x)), (1, C, 128, 128))
            for k in range(K):
                self.A[k].weight.data /= np.sqrt(L)
                self.B[k].weight.data /= np.sqrt(L)
        self.K, self.M, self.P, self.s, self.t0, self.adaptive = K, M, P, s, t0, adaptive

    def load_state_dict(self, state_dict, *args, **kwargs):
        if "g" not in state_dict:                                 # checkpoints that predate `g`
            state_dict = dict(state_dict)
            state_dict["g"] = self.g.detach().clone()
        return super().load_state_dict(state_dict, *args, **kwargs)

    @torch.no_grad()
    def project(self):
        """ l2-ball projection for filters, R_+ projection for thresholds """
        self.t.clamp_(0.0)
        for k in range(self.K):
            self.A[k].weight.data = uball_project(self.A[k].weight.data)
            self.B[k].weight.data = uball_project(self.B[k].weight.data)


class ResidualBlock(nn.Module):
    """Two 3x3x3 convolutions with a skip connection (reference model/net.py:105-120).  Only used when
    CDLNetVideo(residual=True), which no shipped config enables; it stays on the stock torch route."""

    def __init__(self, in_channels, out_channels, kernel_size=(3, 3, 3), stride=1, padding=1):
        super().__init__()
        self.conv1 = nn.Conv3d(in_channels, out_channels, kernel_size, stride=stride, padding=padding, bias=False)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv3d(out_channels, out_channels, kernel_size, stride=stride, padding=padding, bias=False)

    def forward(self, x):
        out = self.conv2(self.relu(self.conv1(x)))
        out += x
        return self.relu(out)


class CDLNetVideo(_ISTANet):
    """ Convolutional Dictionary Learning Network for video denoising.
    P is (frames, rows, cols) — the tuple goes unchanged to nn.Conv3d — or an int for a
    cubic filter (the form args3d.json uses)."""
    _nsp = 3

    def __init__(self,
                 K=3,
                 M=64,
                 P=(7, 7, 5),
                 s=1,
                 C=1,
                 t0=0,
                 adaptive=False,
                 depth=3,
                 init=True,
                 residual=False):
        super().__init__()
        P3 = tuple(int(p) for p in P) if isinstance(P, (tuple, list)) else (int(P),) * 3
        pad = tuple(p // 2 for p in P3)
        self.A = nn.ModuleList([nn.Conv3d(C, M, P3, stride=s, padding=pad, bias=False) for _ in range(K)])
        self.B = nn.ModuleList([nn.ConvTranspose3d(M, C, P3, stride=s, padding=pad, output_padding=s - 1, bias=False)
                                for _ in range(K)])
        self.D = self.B[0]
        self.t = nn.Parameter(t0 * torch.ones(K, 2, M, 1, 1, 1))
        self.residual = residual
        if self.residual:
            self.residual_blocks = nn.ModuleList([ResidualBlock(M, M) for _ in range(K)])
        W = torch.randn(M, C, *P3)
        for k in range(K):
            self.A[k].weight.data = W.clone()
            self.B[k].weight.data = W.clone()
        if init:
            L = _spectral_constant(lambda x: self.D(self.A[0](x)), (1, C, depth, 128, 128))
            for k in range(K):
                self.A[k].weight.data /= np.sqrt(L)
                self.B[k].weight.data /= np.sqrt(L)
        self.K, self.M, self.P, self.s, self.t0, self.adaptive = K, M, P, s, t0, adaptive

    def _after(self, k, z):
        return self.residual_blocks[k](z) if self.residual else z

    @torch.no_grad()
    def project(self):
        self.t.clamp_(0.0)
        for k in range(self.K):
            self.A[k].weight.data = uball_project(self.A[k].weight.data, dim=(2, 3, 4))
            self.B[k].weight.data = uball_project(self.B[k].weight.data, dim=(2, 3, 4))


class GDLNet(_ISTANet):
    """ Gabor Dictionary Learning Network: the same loop with Gabor-parameterised filter banks;
    analysis is `A[k].T`, synthesis `B[k]` (reference model/net.py:659-675)."""
    _nsp = 2

    def __init__(self,
                 K=3,
                 M=64,
                 P=7,
                 s=1,
                 C=1,
                 t0=0,
                 order=1,          # mixture-of-Gabor order
                 adaptive=False,
                 shared="",        # which Gabor parameters are shared across layers, e.g. "a_psi_w0_alpha"
                 init=True):
        super().__init__()
        self.A = nn.ModuleList([ConvAdjoint2dGabor(M, C, P, stride=s, order=order) for _ in range(K)])
        self.B = nn.ModuleList([ConvAdjoint2dGabor(M, C, P, stride=s, order=order) for _ in range(K)])
        self.D = self.B[0]
        self.t = nn.Parameter(t0 * torch.ones(K, 2, M, 1, 1))
        draw = dict(alpha=torch.randn(order, M, C, 1, 1), a=torch.randn(order, M, C, 2),
                    w0=torch.randn(order, M, C, 2), psi=torch.randn(order, M, C))
        for k in range(K):
            for bank in (self.A[k], self.B[k]):
                for name, value in draw.items():
                    getattr(bank, name).data = value.clone()
            if k > 0:                                             # parameter tying (reference :607-622)
                if "alpha" in shared:
                    self.A[k].alpha = self.A[0].alpha
                    if k > 1:                                     # never tie alpha with the final dictionary B[0]
                        self.B[k].alpha = self.B[1].alpha
                for token, name in (("a_", "a"), ("w0", "w0"), ("psi", "psi")):
                    if token in shared:
                        setattr(self.A[k], name, getattr(self.A[0], name))
                        setattr(self.B[k], name, getattr(self.B[0], name))
        if init:
            L = _spectral_constant(lambda x: self.D(self.A[0].T(x)), (1, C, 128, 128))
            for k in range(K):
                self.A[k].alpha.data /= np.sqrt(L)
                self.B[k].alpha.data /= np.sqrt(L)
                if "alpha" in shared:
                    self.B[1].alpha.data /= np.sqrt(L)
                    break
        self.K, self.M, self.P, self.s, self.t0, self.order, self.adaptive = K, M, P, s, t0, order, adaptive

    @torch.no_grad()
    def project(self):
        self.t.clamp_(0.0)

    def _analysis(self, k, x):
        return self.A[k].T(x)

    def _filter_banks(self):
        with torch.no_grad():
            return [m.get_filter(transpose=True) for m in self.A], [m.get_filter() for m in self.B]
