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
    #: "auto" picks the fastest kernel family that meets the parity bar (max|xhat - fp32| <= 1e-4) for the geometry AND
    #: the weights: the tcgen05 (tf32 operand) kernels where they exist, after a one-time calibration per set of weights
    #: (`_calibrate`) has shown that they stay inside the bar - otherwise the exact fp32 kernels;
    #: "fp32" forces the exact CUDA-core kernels, "tf32" / "tf32x3" force the tcgen05 path (single-pass operands / 3-term
    #: split analysis, 2-D stride-1 geometries) without calibration.
    precision = os.environ.get("CDL_PRECISION", "auto")
    #: `auto` keeps the tensor-core kernels only if, on a calibration crop of the first input, they agree with the exact
    #: fp32 kernels to this max-abs deviation on xhat (the parity bar is 1e-4; the margin covers crop-vs-full variation)
    auto_tolerance = 6.5e-5
    #: replay the forward from a CUDA graph (Plan.denoise_graphed): worth it for launch-bound inputs (one small image);
    #: opt-in because the outputs are clones of static buffers and the first call per shape pays the capture
    use_cuda_graph = os.environ.get("CDL_CUDA_GRAPH", "0") == "1"

    # -- plumbing ----------------------------------------------------------------------------------
    def __getstate__(self):
        state = self.__dict__.copy()
        for k in ("_plans", "_last_plan", "_auto_choice", "_embed_choice", "_last_calibration"):
            state.pop(k, None)           # ctypes handles are per process
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
        if y.shape[1] != self._in_channels():
            return False                 # the stock route raises the reference's conv shape error
        if torch.is_tensor(mask) and not (tuple(mask.shape[1:]) == tuple(y.shape[1:]) and mask.shape[0] in (1, y.shape[0])):
            return False                 # pre_process divides by mask.sum() over the mask's OWN shape (model/utils.py:11,76):
                                         # only a batch-broadcast mask has the same sum after expansion
        if not torch.is_tensor(mask) and mask != 1:
            return False
        if torch.is_tensor(sigma) and sigma.numel() not in (1, y.shape[0]):
            return False                 # per-pixel sigma maps are not part of the reference's callers
        return True

    def _in_channels(self):
        w = getattr(self.A[0], "weight", None)
        return int(w.shape[1]) if w is not None else int(self.A[0].alpha.shape[2])

    def _plan_for(self, shape, has_mask, device_index, prec):
        plans = self.__dict__.setdefault("_plans", {})
        key = (tuple(shape), has_mask, device_index, prec)
        plan = plans.get(key)
        if plan is None:
            if len(plans) >= 8:
                plans.pop(next(iter(plans))).close()
            plan = Plan(self._nsp, shape[0], shape[1], self.M, self.K, tuple(shape[2:]), self._plan_P(prec), self.s,
                        has_mask=has_mask, precision=prec, device=device_index or 0)
            plans[key] = plan
        return plan

    def _plan_P(self, prec):
        """Filter extents the plan is created with.  The video tensor-core kernels are written for 7x7x7 (stride 2, C = 1,
        M <= 176); any smaller odd extents (the constructor's default (7,7,5), (5,5,5), ...) are the same operator with the
        filters zero-embedded in the centre of a 7x7x7 box - padding P//2 and the output extents are unchanged - so the
        tensor-core families get the embedded geometry (2/7 of the MMAs multiply zeros for (7,7,5)); the exact fp32 family
        keeps the native extents."""
        P = self._P3()
        if (prec != "fp32" and self._nsp == 3 and self.s == 2 and self._in_channels() == 1 and self.M <= 176
                and P != (7, 7, 7) and all(q % 2 == 1 and q <= 7 for q in P)):
            return (7, 7, 7)
        return P

    def _banks_for(self, plan, A=None, B=None):
        """filter banks in the extents `plan` was created with (zero-embedded when the plan is wider than the filters)"""
        if A is None:
            A, B = self._filter_banks()
        want = tuple(plan.Pfull[-self._nsp:])
        have = tuple(A[0].shape[2:])
        if want == have:
            return A, B
        pad = []
        for w_, h_ in zip(reversed(want), reversed(have)):         # F.pad order: innermost axis first
            pad += [(w_ - h_) // 2, (w_ - h_) // 2]
        return [F.pad(a.detach(), pad) for a in A], [F.pad(b.detach(), pad) for b in B]

    def _weights_key(self):
        """Identity of the current weights: (storage, version counter) of every parameter plus an epoch that
        `refresh_weights()` bumps.  In-place edits through `.data` (p.data.copy_(ema), net.t.data.clamp_(0)) bump
        neither the pointer nor `p._version` - after such an edit call `refresh_weights()`; a module in training mode
        is repacked on every call."""
        return (self.__dict__.get("_weights_epoch", 0),) + tuple((p.data_ptr(), p._version) for p in self.parameters())

    def refresh_weights(self):
        """Forget the packed filters / thresholds and the `auto` calibration: the next forward repacks from the
        parameters.  Needed only after in-place edits that bypass autograd's version counter (`.data` writes)."""
        self.__dict__["_weights_epoch"] = self.__dict__.get("_weights_epoch", 0) + 1
        self.__dict__.pop("_auto_choice", None)
        self.__dict__.pop("_embed_choice", None)
    invalidate = refresh_weights

    def _set_plan_weights(self, plan, key):
        if self.training:
            key = None                   # parameters are expected to move: never trust a cached pack
        if key is not None and plan._weights_key == key:
            return                       # (GDLNet: the Gabor banks are not even synthesised)
        A, B = self._banks_for(plan)
        plan.set_weights(A, B, self.t, key=key)

    def _precision_for(self, y, mask, c, key):
        """Kernel family for this input: the module's `precision`, with "auto" resolved per set of weights."""
        if self.precision in ("fp32", "tf32", "tf32x3"):
            return self.precision
        if self.precision != "auto":
            raise ValueError(f"unknown precision {self.precision!r}")
        choice = self.__dict__.setdefault("_auto_choice", {})
        ck = (key, tuple(y.shape[1:]), mask is not None)
        if ck not in choice or self.training:
            choice[ck] = self._calibrate(y, mask, c)
        self.__dict__["_last_calibration"] = choice[ck]
        return choice[ck][0]

    def _calibrate(self, y, mask, c):
        """One-time check of the tensor-core family against the exact fp32 kernels ON THE GPU, on a crop of the first
        input (first sample; at most 16 frames x 128 x 128, 2-D: 256 x 256): single-pass tf32 operands cost 2-6e-5 on
        xhat for reference-like weights but grow with the amplitude and density of the code, so `auto`
        measures instead of assuming.  Returns (family, measured max-abs deviation or None)."""
        lim = (16, 128, 128) if self._nsp == 3 else (256, 256)
        sl = (slice(0, 1), slice(None)) + tuple(slice(0, min(int(n), m)) for n, m in zip(y.shape[2:], lim))
        yc = y[sl].contiguous()
        mc = None if mask is None else mask[(slice(0, 1),) + sl[1:]].expand_as(yc).contiguous()
        cc = None if c is None else c[:1].contiguous()
        A, B = self._filter_banks()
        p_ex = self._plan_for(yc.shape, mc is not None, y.device.index, "fp32")
        ref, dev = None, None
        for family in ("tf32", "tf32x3"):
            p_tc = self._plan_for(yc.shape, mc is not None, y.device.index, family)
            if p_tc.precision != family:
                if family == "tf32":
                    return ("tf32", None)        # no tensor-core kernel for this geometry: the request resolves to the fp32 family anyway
                continue                          # no 3-term kernel for this geometry (video): next stop is fp32
            if ref is None:
                p_ex.set_weights(*self._banks_for(p_ex, A, B), self.t)
                ref = p_ex.denoise(yc, mc, cc, want_z=False)[0]
            p_tc.set_weights(*self._banks_for(p_tc, A, B), self.t)
            dev = float((p_tc.denoise(yc, mc, cc, want_z=False)[0] - ref).abs().max())
            if dev <= self.auto_tolerance:
                return (family, dev)
        return ("fp32", dev)

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
        c = self._c_vector(sigma, y.shape[0], y.device)
        with torch.cuda.device(y.device):
            key = self._weights_key()
            prec = self._precision_for(y, mask, c, key)
            plan = self._plan_for(y.shape, has_mask, y.device.index, prec)
            self._set_plan_weights(plan, key)
        self.__dict__["_last_plan"] = plan
        return plan, y, mask, c

    # -- the hot path ------------------------------------------------------------------------------
    def forward(self, y, sigma=None, mask=1):
        """ LISTA + D w/ noise-adaptive thresholds """
        if not self._native_ok(y, sigma, mask):
            return self._forward_stock(y, sigma, mask)
        if self._embed3d_ok(y, mask, sigma):
            return self._forward_embedded3d(y, sigma)
        plan, y, mask, c = self._prepare(y, sigma, mask)
        with torch.cuda.device(y.device):
            if self.use_cuda_graph and not torch.cuda.is_current_stream_capturing():
                return plan.denoise_graphed(y, mask, c)
            return plan.denoise(y, mask, c)

    # -- 2-D stride-2 grayscale nets (CDLNet-s2030, BASELINE config 1) on the VIDEO tensor-core kernels ------------------
    # A 2-D image is a clip of two frames (image, zero) and a 7x7 filter is the td = 3 slice of a 7x7x7 filter: with
    # stride 2 and padding 3 along d the only coarse frame reads fine frames td - 3 = 0 (td = 3) and 1 (td = 4, a zero
    # slice), and the synthesis writes frame 0 from td = 3 and nothing into frame 1 - the 3-D operator restricted to
    # frame 0 IS the 2-D operator.  6/7 of the MMAs multiply zeros, but one 256x256 image is launch-bound either way:
    # measured 1.01 ms against 5.68 ms on the fp32 CUDA-core kernels, max|xhat - oracle| 2.3e-5
    # (profiles/r02q_cfg1_embed3d_timing.json, r02o_embed3d_tests.log).  Taken for precision "tf32", and for "auto" after
    # the same one-time calibration against the exact kernels as `_calibrate`; CDL_EMBED3D=0 disables the route.
    def _embed3d_ok(self, y, mask, sigma=None):
        if os.environ.get("CDL_EMBED3D", "1") == "0" or self.precision not in ("tf32", "auto"):
            return False
        if self._nsp != 2 or type(self).__name__ != "CDLNet" or torch.is_tensor(mask):
            return False
        if not (self.s == 2 and self._P3() == (7, 7) and y.shape[1] == 1 and self.M <= 176):
            return False
        if (-(-y.shape[3] // 2) * 2) % 4 != 0:               # the video kernels need a padded width that is a multiple of 4
            return False
        if self.precision == "tf32":
            return True
        choice = self.__dict__.setdefault("_embed_choice", {})
        ck = (self._weights_key(), tuple(y.shape[1:]))
        if ck not in choice or self.training:
            choice[ck] = self._calibrate_embed3d(y, sigma)
        self.__dict__["_last_calibration"] = choice[ck]
        return choice[ck][0] == "tf32"

    def _calibrate_embed3d(self, y, sigma):
        """embedded route vs the exact fp32 kernels on (a crop of) the first sample -> ("tf32" | "fp32", deviation)"""
        yc = y[:1, :, :min(int(y.shape[2]), 256), :min(int(y.shape[3]), 256)].contiguous()
        sc = sigma.reshape(-1)[:1] if torch.is_tensor(sigma) else sigma
        if (-(-yc.shape[3] // 2) * 2) % 4 != 0:
            return ("fp32", None)
        with torch.cuda.device(y.device):
            p_ex = self._plan_for(yc.shape, False, y.device.index, "fp32")
            A, B = self._filter_banks()
            p_ex.set_weights(A, B, self.t)
            ref = p_ex.denoise(yc, None, self._c_vector(sc, 1, y.device), want_z=False)[0]
            dev = float((self._forward_embedded3d(yc, sc)[0] - ref).abs().max())
        return ("tf32" if dev <= self.auto_tolerance else "fp32", dev)

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
            self.__dict__["_last_plan"] = p3
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


# ------------------------------------------------------------------------------------------------
# frame-recurrent CSR variants (reference model/net.py:229-262, 363-567)
# ------------------------------------------------------------------------------------------------
def prox_CSR(u, z_prev, lambd, gamma):
    """ST(ST(u - z_prev - lambd*sign(z_prev), lambd*gamma) + z_prev + lambd*sign(z_prev), lambd)"""
    shift = lambd * torch.sign(z_prev)
    return ST(ST(u - z_prev - shift, lambd * gamma) + z_prev + shift, lambd)


def prox_CSR_f2(u, z_prev, z_after, lambd, gamma1, gamma2):
    """two-neighbour CSR proximal operator (previous and next frame)"""
    Ca = z_prev + lambd * torch.sign(z_prev) + lambd * gamma2 * torch.sign(z_prev - z_after)
    Cb = z_after + lambd * torch.sign(z_after) + lambd * gamma1 * torch.sign(z_after - z_prev)
    kick = lambd * gamma1 * torch.sign(u - Ca)
    inner = ST(u - Ca, gamma1 * lambd)
    midder = ST(inner - Cb + kick, gamma2 * lambd)
    return ST(midder + Cb - kick, lambd)


class _CSRNet(_ISTANet):
    """2-D CDLNet whose proximal step couples a frame's code to the neighbouring frames' codes.  The convolutions are the
    ordinary analysis / synthesis steps; on CUDA (no grad) the loop runs on the exact fp32 kernels of libcdl_b200 with
    the CSR proximal operator fused into the analysis epilogue (cdl_analysis_step_csr) - the shipped hyper-parameters
    (argscsr.json: P = 9, s = 2) have no tensor-core kernel, and the prox chain is kept in fp32 throughout."""
    _nsp = 2

    def _build(self, K, M, P, s, C, t0, adaptive, init, second_bank):
        pad = (P - 1) // 2
        conv = lambda: nn.ModuleList([nn.Conv2d(C, M, P, stride=s, padding=pad, bias=False) for _ in range(K)])
        convT = lambda: nn.ModuleList([nn.ConvTranspose2d(M, C, P, stride=s, padding=pad, output_padding=s - 1, bias=False) for _ in range(K)])
        self.A, self.B = conv(), convT()
        if second_bank:                                           # CDLNet_CSR: the frame without a neighbour has its own operators
            self.A2, self.B2 = conv(), convT()
        self.D = self.B[0]
        self.t = nn.Parameter(t0 * torch.ones(K, 2, M, 1, 1))
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

    @torch.no_grad()
    def project(self):
        """ l2-ball projection for filters, R_+ projection for thresholds """
        self.t.clamp_(0.0)
        for k in range(self.K):
            self.A[k].weight.data = uball_project(self.A[k].weight.data)
            self.B[k].weight.data = uball_project(self.B[k].weight.data)

    # banks[name] = (A modules, B modules, thresholds) of one operator set
    def _bank(self, second):
        return (self.A2, self.B2, self.t2) if second else (self.A, self.B, self.t)

    def _csr_forward(self, y, sigma, mask, z_prev, z_after, g_prev, g_after, second_bank=False):
        """the shared loop: `second_bank` selects (A2, B2, t2) for the iterations (D stays B[0])"""
        A, B, t = self._bank(second_bank)
        if not self._native_ok(y, sigma, mask) or any(n is not None and not (n.is_cuda and n.dtype == torch.float32) for n in (z_prev, z_after)):
            return self._csr_stock(y, sigma, mask, z_prev, z_after, g_prev, g_after, A, B, t)
        has_mask = torch.is_tensor(mask)
        y = y.contiguous()
        mask = mask.to(device=y.device, dtype=torch.float32).expand_as(y).contiguous() if has_mask else None
        c = self._c_vector(sigma, y.shape[0], y.device)
        with torch.cuda.device(y.device):
            key = self._weights_key()
            plan = self._plan_for(y.shape, has_mask, y.device.index, "fp32")
            tag = (key, bool(second_bank))
            if self.training or plan._weights_key != tag:
                plan.set_weights([m.weight for m in A], [m.weight for m in B], t, key=tag)
            plan_d = plan
            if second_bank:                                       # the final D z uses B[0] of the FIRST bank (model/net.py:459)
                plan_d = self._plan_for(y.shape, has_mask, y.device.index, "fp32-d")
                if self.training or plan_d._weights_key != (key, "d"):
                    plan_d.set_weights([m.weight for m in self.A], [m.weight for m in self.B], self.t, key=(key, "d"))
            self.__dict__["_last_plan"] = plan
            zp = None if z_prev is None else z_prev.contiguous()
            za = None if z_after is None else z_after.contiguous()
            gp = None if zp is None else g_prev.detach().reshape(self.K, 2, self.M).contiguous()
            ga = None if za is None else g_after.detach().reshape(self.K, 2, self.M).contiguous()
            yp, mp, mean = plan.preprocess(y, mask)
            z = plan.new_code()
            r = torch.empty_like(yp)
            plan.analysis_step_csr(0, yp, z, c, first=True, z_prev=zp, z_after=za, g1=gp, g2=ga)
            for k in range(1, self.K):
                plan.synthesis_step(k, z, r, yp, mp, residual=True)
                plan.analysis_step_csr(k, r, z, c, z_prev=zp, z_after=za, g1=gp, g2=ga)
            plan_d.synthesis_step(0, z, r, residual=False)
            return plan.postprocess(r, mean), plan.export_code(z)

    def _plan_for(self, shape, has_mask, device_index, prec):
        if prec == "fp32-d":                                      # a second exact plan holding the first bank (for D)
            plans = self.__dict__.setdefault("_plans", {})
            key = (tuple(shape), has_mask, device_index, prec)
            if key not in plans:
                plans[key] = Plan(2, shape[0], shape[1], self.M, self.K, tuple(shape[2:]), self._P3(), self.s,
                                  has_mask=has_mask, precision="fp32", device=device_index or 0)
            return plans[key]
        return super()._plan_for(shape, has_mask, device_index, prec)

    def _csr_stock(self, y, sigma, mask, z_prev, z_after, g_prev, g_after, A, B, t):
        yp, params, mask = pre_process(y, self.s, mask=mask)
        c = 0 if sigma is None or not self.adaptive else sigma / 255.0
        thr = lambda p, k: p[k, :1] + c * p[k, 1:2]

        def prox(u, k):
            if z_prev is not None and z_after is not None:
                return prox_CSR_f2(u, z_prev, z_after, thr(t, k), thr(g_prev, k), thr(g_after, k))
            if z_prev is not None:
                return prox_CSR(u, z_prev, thr(t, k), thr(g_prev, k))
            if z_after is not None:
                return prox_CSR(u, z_after, thr(t, k), thr(g_after, k))
            return ST(u, thr(t, k))
        z = prox(A[0](yp), 0)
        for k in range(1, self.K):
            z = prox(z - A[k](mask * B[k](z) - yp), k)
        return post_process(self.D(z), params), z


class CDLNet_CSR(_CSRNet):
    """ CDLNet with the one-neighbour CSR proximal step (reference model/net.py:363-462): with `z_prev` the iterations use
    (A, B, t, g) and prox_CSR; without it they are the plain ISTA loop on a second operator set (A2, B2, t2). """

    def __init__(self, K=3, M=64, P=7, s=1, C=1, t0=0, adaptive=False, init=True):
        super().__init__()
        self._build(K, M, P, s, C, t0, adaptive, init, second_bank=True)
        self.t2 = nn.Parameter(t0 * torch.ones(K, 2, M, 1, 1))
        self.g = nn.Parameter(t0 * torch.ones(K, 2, M, 1, 1))

    def forward(self, y, z_prev=None, sigma=None, mask=1):
        """ LISTA + D; `z_prev` = sparse code of the previous frame (None for the first frame) """
        if z_prev is None:
            return self._csr_forward(y, sigma, mask, None, None, None, None, second_bank=True)
        return self._csr_forward(y, sigma, mask, z_prev, None, self.g, None)


class CDLNet_CSRf2(_CSRNet):
    """ CDLNet with the two-neighbour CSR proximal step (reference model/net.py:464-567; argscsr.json): prox_CSR_f2 with both
    neighbours, prox_CSR with one (g1 pairs with z_prev, g2 with z_after), plain ST with none. """

    def __init__(self, K=3, M=64, P=7, s=1, C=1, t0=0, adaptive=False, init=True):
        super().__init__()
        self._build(K, M, P, s, C, t0, adaptive, init, second_bank=False)
        self.g1 = nn.Parameter(t0 * torch.ones(K, 2, M, 1, 1))
        self.g2 = nn.Parameter(t0 * torch.ones(K, 2, M, 1, 1))

    def forward(self, y, z_prev=None, z_after=None, sigma=None, mask=1):
        return self._csr_forward(y, sigma, mask, z_prev, z_after, self.g1, self.g2)
