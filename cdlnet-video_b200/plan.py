"""Host-side handle on a `cdl_plan_t` (include/cdl_b200.h): geometry, packed filters, workspace.

PyTorch is used only for device memory and streams: every tensor handed to the library is a
contiguous fp32 CUDA tensor whose `data_ptr()` is passed through ctypes.
"""
import ctypes

import torch

from . import _lib

PREC = {"fp32": 0, "tf32": 1, "tf32x3": 2}


def _ptr(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class Plan:
    """One geometry: (ndim, N, C, M, K, dims, P, s, mask?, precision[, temporal halos]) on one device."""

    def __init__(self, ndim, N, C, M, K, dims, P, s, has_mask=False, precision="fp32", device=0,
                 halo_front=0, halo_back=0):
        self.lib = _lib.load()
        d = _lib.CdlDesc()
        d.ndim, d.N, d.C, d.M, d.K, d.s = ndim, N, C, M, K, s
        dims3 = (1, *dims) if ndim == 2 else tuple(dims)
        P3 = (1, *P) if ndim == 2 else tuple(P)
        for i in range(3):
            d.dims[i] = int(dims3[i])
            d.P[i] = int(P3[i])
        d.has_mask = int(bool(has_mask))
        d.precision = PREC[precision] if isinstance(precision, str) else int(precision)
        d.halo_front, d.halo_back, d.device = int(halo_front), int(halo_back), int(device)
        self.desc = d
        self.device = torch.device("cuda", int(device))
        handle = ctypes.c_void_p()
        _lib.check(self.lib.cdl_plan_create(ctypes.byref(handle), ctypes.byref(d)), "cdl_plan_create")
        self.handle = handle
        lay = _lib.CdlLayout()
        _lib.check(self.lib.cdl_plan_layout(handle, ctypes.byref(lay)), "cdl_plan_layout")
        self.pad = tuple(lay.pad)                    # (l, r, t, b, f, k)
        self.fine = tuple(lay.fine)                  # padded (D,H,W)
        self.coarse = tuple(lay.coarse)              # z (D,H,W)
        self.ndim, self.N, self.C, self.M, self.K, self.s = ndim, N, C, M, K, s
        self.dims = tuple(int(v) for v in dims)
        self.Pfull = tuple(int(v) for v in P3)
        self.has_mask = bool(has_mask)
        n = ctypes.c_size_t()
        # the workspace regions are prefixes of one another: reduce < step < forward < host (< host + z staging)
        self._ws_bytes = {}
        for kind, fn in (("reduce", self.lib.cdl_plan_reduce_workspace_bytes), ("step", self.lib.cdl_plan_step_workspace_bytes),
                         ("forward", self.lib.cdl_plan_workspace_bytes), ("host", self.lib.cdl_plan_host_workspace_bytes_noz),
                         ("host_z", self.lib.cdl_plan_host_workspace_bytes)):
            _lib.check(fn(handle, ctypes.byref(n)))
            self._ws_bytes[kind] = n.value
        self.workspace_bytes = self._ws_bytes["forward"]
        self.host_workspace_bytes = self._ws_bytes["host_z"]
        _lib.check(self.lib.cdl_plan_code_bytes(handle, ctypes.byref(n)))
        self.code_bytes = n.value
        self._ws = None
        self._weights_key = None
        self._keep = None

    # shapes -------------------------------------------------------------------------------------
    @property
    def fine_shape(self):
        return (self.N, self.C, *(self.fine if self.ndim == 3 else self.fine[1:]))

    @property
    def z_shape(self):
        return (self.N, self.M, *(self.coarse if self.ndim == 3 else self.coarse[1:]))

    @property
    def in_shape(self):
        return (self.N, self.C, *self.dims)

    @property
    def precision(self):
        return {0: "fp32", 1: "tf32", 2: "tf32x3"}[self.lib.cdl_plan_precision(self.handle)]

    def set_rearm(self, enable=True):
        """Stepwise drivers only: let analysis_step overwrite its consumed input r with -yp for the next residual synthesis
        into the same buffer (see cdl_plan_set_rearm in include/cdl_b200.h)."""
        _lib.check(self.lib.cdl_plan_set_rearm(self.handle, 1 if enable else 0), "cdl_plan_set_rearm")

    def launch_count(self):
        n = ctypes.c_uint64()
        _lib.check(self.lib.cdl_plan_launch_count(self.handle, ctypes.byref(n)))
        return n.value

    def workspace(self, kind="forward", host=False):
        """Device workspace large enough for `kind`: "reduce" (cdl_reduce_sums only), "step" (stepwise entry points),
        "forward" (cdl_forward / cdl_denoise), "host" / "host_z" (cdl_denoise_host without / with z).  Grows on demand."""
        if host:
            kind = "host_z"
        need = self._ws_bytes[kind]
        if self._ws is None or self._ws.numel() < need:
            self._ws = None                          # release before growing (the forward workspace of a long clip is tens of GB)
            self.__dict__.pop("_graphs", None)       # captured forwards point into the old workspace
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def close(self):
        if getattr(self, "handle", None):
            self.lib.cdl_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # weights ------------------------------------------------------------------------------------
    def set_weights(self, A, B, t, key=None):
        """A, B: K tensors (M,C,P...) fp32 CUDA; t: (K,2,M,...) fp32 CUDA."""
        if key is not None and key == self._weights_key:
            return
        self.__dict__.pop("_graphs", None)           # captured forwards hold the old packed filters' launches: re-capture
        K = self.K
        P = self.Pfull[-self.ndim:]
        want = (self.M, self.C, *P)
        if len(A) != K or len(B) != K or t.numel() != K * 2 * self.M:
            raise ValueError(f"set_weights: expected {K} analysis and {K} synthesis banks and {K}x2x{self.M} thresholds")
        for name, bank in (("A", A), ("B", B)):
            for k, w in enumerate(bank):
                if tuple(w.shape) != want:          # the library only sees raw pointers: a mismatch would read out of bounds
                    raise ValueError(f"set_weights: {name}[{k}] has shape {tuple(w.shape)}, the plan needs {want}")
        A = [a.detach().to(self.device, torch.float32).contiguous() for a in A]
        B = [b.detach().to(self.device, torch.float32).contiguous() for b in B]
        t = t.detach().to(self.device, torch.float32).reshape(K, 2, self.M).contiguous()
        arrA = (ctypes.c_void_p * K)(*[a.data_ptr() for a in A])
        arrB = (ctypes.c_void_p * K)(*[b.data_ptr() for b in B])
        _lib.check(self.lib.cdl_set_weights(self.handle, arrA, arrB, _ptr(t), _stream()), "cdl_set_weights")
        self._keep = (A, B, t)      # keep sources alive until the async repack has run
        self._weights_key = key

    # whole forward ------------------------------------------------------------------------------
    def denoise(self, y, mask=None, c=None, z_out=None, want_z=True):
        """y (N,C,dims) -> (xhat like y, z).  c: (N,) fp32 sigma/255 or None.  want_z=False: z is not converted to
        (N,M,coarse) and None is returned for it (the code of a long clip is tens of GB)."""
        if tuple(y.shape) != self.in_shape or (mask is not None and tuple(mask.shape) != self.in_shape):
            raise ValueError(f"denoise: input shape {tuple(y.shape)} does not match the plan's {self.in_shape}")
        xhat = torch.empty(self.in_shape, dtype=torch.float32, device=self.device)
        z = None
        if want_z:
            z = z_out if z_out is not None else torch.empty(self.z_shape, dtype=torch.float32, device=self.device)
        ws = self.workspace()
        _lib.check(self.lib.cdl_denoise(self.handle, _ptr(y), _ptr(mask), _ptr(c), _ptr(xhat), _ptr(z), _ptr(ws), _stream()),
                   "cdl_denoise")
        return xhat, z

    def denoise_graphed(self, y, mask=None, c=None, want_z=True):
        """`denoise` replayed from a CUDA graph (launch-bound shapes: a 256x256 image is ~100 kernels of a few microseconds).
        The first call with a given (mask?, c?, want_z) signature captures `cdl_denoise` on static buffers; later calls copy
        the inputs in, replay, and return clones of the static outputs.  The graph is dropped when the weights change
        (`set_weights`).  The library enqueues only kernels / memsets on the given stream, so the capture needs no special
        entry point."""
        key = (mask is not None, c is not None, bool(want_z))
        graphs = self.__dict__.setdefault("_graphs", {})
        g = graphs.get(key)
        if g is None:
            st = {"y": torch.empty_like(y), "mask": None if mask is None else torch.empty_like(mask), "c": None if c is None else torch.empty_like(c)}
            st["y"].copy_(y)
            if mask is not None:
                st["mask"].copy_(mask)
            if c is not None:
                st["c"].copy_(c)
            self.workspace()                                          # allocate outside the capture
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                             # warm-up (tensor maps encoded, lazy module loading done)
                self.denoise(st["y"], st["mask"], st["c"], want_z=want_z)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                st["xhat"], st["z"] = self.denoise(st["y"], st["mask"], st["c"], want_z=want_z)
            st["graph"] = graph
            g = graphs[key] = st
        g["y"].copy_(y)
        if mask is not None:
            g["mask"].copy_(mask)
        if c is not None:
            g["c"].copy_(c)
        g["graph"].replay()
        return g["xhat"].clone(), (g["z"].clone() if want_z else None)

    def denoise_host(self, y_host, xhat_host, mask_host=None, c_host=None, z_host=None):
        """Pinned host buffers in, pinned host buffers out; asynchronous on the current stream."""
        ws = self.workspace("host" if z_host is None else "host_z")
        _lib.check(self.lib.cdl_denoise_host(self.handle, _ptr(y_host), _ptr(mask_host), _ptr(c_host), _ptr(xhat_host),
                                             _ptr(z_host), _ptr(ws), _stream()), "cdl_denoise_host")

    # stepwise API (forward_generator, temporal slabs) ----------------------------------------------
    # The sparse code lives in the plan's internal layout between steps (see include/cdl_b200.h).
    def new_code(self):
        return torch.empty(self.code_bytes // 4, dtype=torch.float32, device=self.device)   # fully written by the first analysis step

    def export_code(self, code):
        z = torch.empty(self.z_shape, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.cdl_code_export(self.handle, _ptr(code), _ptr(z), _stream()), "cdl_code_export")
        return z

    def import_code(self, z):
        code = self.new_code()
        _lib.check(self.lib.cdl_code_import(self.handle, _ptr(z.contiguous()), _ptr(code), _stream()), "cdl_code_import")
        return code

    def preprocess(self, y, mask=None):
        yp = torch.empty(self.fine_shape, dtype=torch.float32, device=self.device)
        mp = torch.empty(self.fine_shape, dtype=torch.float32, device=self.device) if self.has_mask else None
        mean = torch.empty(self.N, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.cdl_preprocess(self.handle, _ptr(y), _ptr(mask), _ptr(yp), _ptr(mp), _ptr(mean),
                                           _ptr(self.workspace("step")), _stream()), "cdl_preprocess")
        return yp, mp, mean

    def preprocess_noisy(self, x, noise=None, c=None, mask=None, bayer=False, want_y=False):
        """awgn + mask + pre_process fused (cdl_preprocess_noisy; reference utils.py:13-55 + model/utils.py:5-22,70-87):
        x clean (device), noise = the caller's randn_like draw, c = sigma/255 per sample -> (yp, mask_p, mean[, noisy])"""
        yp = torch.empty(self.fine_shape, dtype=torch.float32, device=self.device)
        mp = torch.empty(self.fine_shape, dtype=torch.float32, device=self.device) if self.has_mask else None
        mean = torch.empty(self.N, dtype=torch.float32, device=self.device)
        y = torch.empty_like(x) if want_y else None
        _lib.check(self.lib.cdl_preprocess_noisy(self.handle, _ptr(x), _ptr(noise), _ptr(c), _ptr(mask), 1 if bayer else 0, _ptr(y),
                                                 _ptr(yp), _ptr(mp), _ptr(mean), _ptr(self.workspace("step")), _stream()),
                   "cdl_preprocess_noisy")
        return (yp, mp, mean, y) if want_y else (yp, mp, mean)

    def reduce_sums(self, y, mask=None):
        sums = torch.empty(2 * self.N, dtype=torch.float64, device=self.device)
        _lib.check(self.lib.cdl_reduce_sums(self.handle, _ptr(y), _ptr(mask), _ptr(sums), _ptr(self.workspace("reduce")), _stream()),
                   "cdl_reduce_sums")
        return sums

    def mean_from_sums(self, sums):
        mean = torch.empty(self.N, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.cdl_mean_from_sums(self.handle, _ptr(sums), _ptr(mean), _stream()), "cdl_mean_from_sums")
        return mean

    def center_pad(self, y, mean, mask=None):
        yp = torch.empty(self.fine_shape, dtype=torch.float32, device=self.device)
        mp = torch.empty(self.fine_shape, dtype=torch.float32, device=self.device) if self.has_mask else None
        _lib.check(self.lib.cdl_center_pad(self.handle, _ptr(y), _ptr(mask), _ptr(mean), _ptr(yp), _ptr(mp), _stream()),
                   "cdl_center_pad")
        return yp, mp

    def analysis_step(self, k, r, z, c=None, first=False):
        _lib.check(self.lib.cdl_analysis_step(self.handle, k, int(first), _ptr(r), _ptr(c), _ptr(z),
                                              _ptr(self.workspace("step")), _stream()), "cdl_analysis_step")
        return z

    def analysis_step_csr(self, k, r, z, c=None, first=False, z_prev=None, z_after=None, g1=None, g2=None):
        """analysis step with the CSR proximal operators (cdl_analysis_step_csr; reference model/net.py:229-262)"""
        _lib.check(self.lib.cdl_analysis_step_csr(self.handle, k, int(first), _ptr(r), _ptr(c), _ptr(z), _ptr(z_prev), _ptr(z_after),
                                                  _ptr(g1), _ptr(g2), _ptr(self.workspace("step")), _stream()), "cdl_analysis_step_csr")
        return z

    def synthesis_step(self, k, z, out, yp=None, mask_p=None, residual=True):
        _lib.check(self.lib.cdl_synthesis_step(self.handle, k, int(residual), _ptr(z), _ptr(yp), _ptr(mask_p), _ptr(out),
                                               _ptr(self.workspace("step")), _stream()), "cdl_synthesis_step")
        return out

    # temporal slabs across GPUs ----------------------------------------------------------------------
    @property
    def halo_bytes(self):
        n = ctypes.c_size_t()
        _lib.check(self.lib.cdl_halo_bytes(self.handle, ctypes.byref(n)))
        return n.value

    def analysis_step_halo(self, k, r, z, c, recv_prev, recv_next, yp):
        _lib.check(self.lib.cdl_analysis_step_halo(self.handle, k, _ptr(r), _ptr(c), _ptr(z), _ptr(recv_prev), _ptr(recv_next),
                                                   _ptr(yp), _ptr(self.workspace("step")), _stream()), "cdl_analysis_step_halo")
        return z

    def halo_add(self, r, recv_prev, recv_next, yp=None):
        _lib.check(self.lib.cdl_halo_add(self.handle, _ptr(r), _ptr(recv_prev), _ptr(recv_next), _ptr(yp), _stream()), "cdl_halo_add")

    def comm_allreduce(self, comm, sums):
        _lib.check(self.lib.cdl_comm_allreduce_f64(comm.handle, _ptr(sums), sums.numel(), _stream()), "cdl_comm_allreduce_f64")

    def forward_sharded(self, comm, yp, c, code, r, halo_ws):
        """cdl_forward_sharded: all K iterations + D z of this rank's slab; r returns xphat on the resident frames."""
        _lib.check(self.lib.cdl_forward_sharded(self.handle, comm.handle if comm is not None else None, _ptr(yp), _ptr(c),
                                                _ptr(code), _ptr(r), _ptr(halo_ws), _ptr(self.workspace("step")), _stream()),
                   "cdl_forward_sharded")

    def forward(self, yp, mask_p=None, c=None):
        z = torch.empty(self.z_shape, dtype=torch.float32, device=self.device)
        xp = torch.empty(self.fine_shape, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.cdl_forward(self.handle, _ptr(yp), _ptr(mask_p), _ptr(c), _ptr(z), _ptr(xp),
                                        _ptr(self.workspace()), _stream()), "cdl_forward")
        return z, xp

    def postprocess(self, xp, mean):
        xhat = torch.empty(self.in_shape, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.cdl_postprocess(self.handle, _ptr(xp), _ptr(mean), _ptr(xhat), _stream()), "cdl_postprocess")
        return xhat
