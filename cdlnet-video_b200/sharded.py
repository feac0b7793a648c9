"""Temporal sharding of ONE long clip across ranks (BASELINE config 5).

The reference never does this (it chops videos into independent 16-frame windows, analyze3d.py:62,105);
the module's forward, however, is defined for any clip length, and this driver computes exactly that
forward with the coarse time axis split into contiguous slabs, one per rank:

  * rank r owns coarse frames [q0, q1) and keeps the fine frames [s*q0 - hf, s*q1 + hb) resident, where
    hf = Pd//2 at a seam (0 at the clip start) and hb = Pd//2 - s + 1 at a seam (0 at the clip end);
  * its uncropped synthesis  B z_local  covers exactly that range; on the Pd - s frames it shares with a
    neighbour both ranks hold partial sums, so once per iteration (and once for D z) the two exchange those
    frames (P2P send/recv, ring neighbours only) and add them - image-domain halos, 25x smaller than a
    z-domain halo;
  * the per-sample mean of pre_process_3d needs one all-reduce of per-rank fp64 sums.

`SlabRank` holds one rank's state and is written against a small `ops` interface so that the same driver
logic runs on the CUDA plan (`PlanOps`) and, in the CPU tests, on the oracle (`tests/test_sharded_cpu.py`).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def slab_bounds(Qd_total: int, world: int, rank: int):
    """Contiguous, near-equal split of the coarse frames."""
    base, rem = divmod(Qd_total, world)
    q0 = rank * base + min(rank, rem)
    q1 = q0 + base + (1 if rank < rem else 0)
    return q0, q1


def slab_geometry(D: int, Pd: int, s: int, world: int, rank: int):
    """-> dict(q0, q1, hf, hb, f0, f1): owned coarse frames, halos, resident fine range [f0, f1)."""
    if D % s:
        raise ValueError("temporal sharding needs D divisible by the stride (no temporal stride padding at seams)")
    Qd = D // s
    if Qd < world:
        raise ValueError("fewer coarse frames than ranks")
    q0, q1 = slab_bounds(Qd, world, rank)
    h = Pd // 2
    if world > 1 and Pd - s > 0 and h - s + 1 <= 0:
        # e.g. Pd = 3, s = 2: the seam overlap (Pd - s = 1 frame) lies entirely in the NEXT rank's front halo, so the
        # symmetric exchange below would have one side with nothing to send or receive - not implemented
        raise ValueError(f"temporal sharding needs Pd//2 - s + 1 > 0 (got Pd={Pd}, s={s})")
    hf = h if rank > 0 else 0
    hb = (h - s + 1) if rank < world - 1 else 0
    if min(b - a for a, b in (slab_bounds(Qd, world, r) for r in range(world))) * s < Pd:
        raise ValueError("slabs are thinner than the temporal filter extent")
    return dict(q0=q0, q1=q1, hf=hf, hb=hb, f0=s * q0 - hf, f1=s * q1 + hb, overlap=Pd - s)


class PlanOps:
    """The slab operators on one GPU: a libcdl_b200 plan with temporal halos (no fallback)."""

    def __init__(self, plan, sum_plan):
        self.plan, self.sum_plan = plan, sum_plan
        # SlabRank alternates synthesis(out = r) / [halo add] / analysis(r) on one buffer and never reads r after the
        # analysis: let the analysis step re-arm it with -yp (saves the initialisation pass of every residual synthesis)
        plan.set_rearm(True)

    def owned_sums(self, y_owned):                    # (2N,) float64: sum(y), count
        return self.sum_plan.reduce_sums(y_owned.contiguous())

    def mean_from_sums(self, sums):
        return self.plan.mean_from_sums(sums)

    def center_pad(self, y_loc, mean):
        return self.plan.center_pad(y_loc.contiguous(), mean)[0]

    def new_code(self):
        return self.plan.new_code()

    def new_fine(self):
        return torch.empty(self.plan.fine_shape, dtype=torch.float32, device=self.plan.device)

    def analysis(self, k, r, code, c, first):
        self.plan.analysis_step(k, r, code, c, first=first)

    def analysis_halo(self, k, r, code, c, recv_prev, recv_next, yp):
        """analysis step on r + the neighbours' seam partials (+ yp back): the sum is fused into the step's rounding pass"""
        self.plan.analysis_step_halo(k, r, code, c, recv_prev, recv_next, yp)

    def halo_add(self, r, recv_prev, recv_next, yp=None):
        self.plan.halo_add(r, recv_prev, recv_next, yp)

    def synthesis(self, k, code, out, yp, residual):
        self.plan.synthesis_step(k, code, out, yp, None, residual=residual)

    def postprocess(self, xp, mean):
        return self.plan.postprocess(xp, mean)

    def export_code(self, code):
        return self.plan.export_code(code)


class SlabRank:
    """One rank's share of the forward pass.  Phases are separate methods so that a driver can interleave the
    neighbour exchange (distributed) or run several ranks in lock step inside one process (tests)."""

    def __init__(self, ops, geo, K, s):
        self.ops, self.geo, self.K, self.s = ops, geo, K, s
        self.ov = geo["overlap"]

    # -- preprocess --------------------------------------------------------------------------------
    def local_sums(self, y_loc):
        g = self.geo
        owned = y_loc[:, :, g["hf"]:y_loc.shape[2] - g["hb"]]
        self.y_loc = y_loc
        return self.ops.owned_sums(owned)

    def set_global_sums(self, sums, c):
        self.mean = self.ops.mean_from_sums(sums)
        self.yp = self.ops.center_pad(self.y_loc, self.mean)
        self.c = c
        if getattr(self, "code", None) is None:          # buffers are allocated once and reused across calls
            self.code = self.ops.new_code()
            self.r = self.ops.new_fine()

    # -- iterations ----------------------------------------------------------------------------------
    def first(self):
        self.ops.analysis(0, self.yp, self.code, self.c, True)

    def synth(self, k, residual=True):
        """Local partial B_k z (minus yp when residual); returns the (head, tail) overlap slices to send."""
        self.ops.synthesis(k, self.code, self.r, self.yp if residual else None, residual)
        self._residual = residual
        return self._head(self.r), self._tail(self.r)

    def _head(self, t):
        return t[:, :, :self.ov] if self.geo["hf"] else None

    def _tail(self, t):
        return t[:, :, t.shape[2] - self.ov:] if self.geo["hb"] else None

    def add_halo(self, recv_prev, recv_next):
        """r_overlap = mine + theirs (+ yp: both partials already carry -yp in residual mode)."""
        if recv_prev is not None:
            h = self._head(self.r)
            h.add_(recv_prev)
            if self._residual:
                h.add_(self._head(self.yp))
        if recv_next is not None:
            t = self._tail(self.r)
            t.add_(recv_next)
            if self._residual:
                t.add_(self._tail(self.yp))

    def ana(self, k):
        self.ops.analysis(k, self.r, self.code, self.c, False)

    def halo_ana(self, k, recv_prev, recv_next):
        """add_halo + ana; one fused pass where the operators offer it (the CUDA plan), two steps otherwise (oracle ops)"""
        if hasattr(self.ops, "analysis_halo") and self._residual:
            rp = recv_prev.contiguous() if recv_prev is not None else None
            rn = recv_next.contiguous() if recv_next is not None else None
            self.ops.analysis_halo(k, self.r, self.code, self.c, rp, rn, self.yp)
        else:
            self.add_halo(recv_prev, recv_next)
            self.ana(k)

    def finish(self):
        """After the final synth(0, residual=False) + add_halo: crop to the owned frames."""
        x = self.ops.postprocess(self.r, self.mean)
        g = self.geo
        return x[:, :, g["hf"]:x.shape[2] - g["hb"]], self.ops.export_code(self.code)


# ------------------------------------------------------------------------------------------------
# drivers
# ------------------------------------------------------------------------------------------------
def run_lockstep(ranks, y_slabs, c):
    """All ranks inside one process (single-GPU emulation / CPU tests): the exchange is a tensor hand-over."""
    sums = [r.local_sums(y) for r, y in zip(ranks, y_slabs)]
    total = sum(s.double() for s in sums)
    for r in ranks:
        r.set_global_sums(total.clone(), c)

    def exchange(pairs):
        heads = [p[0] for p in pairs]
        tails = [p[1] for p in pairs]
        snap_h = [h.clone() if h is not None else None for h in heads]
        snap_t = [t.clone() if t is not None else None for t in tails]
        for i, r in enumerate(ranks):
            r.add_halo(snap_t[i - 1] if i > 0 else None, snap_h[i + 1] if i + 1 < len(ranks) else None)

    for r in ranks:
        r.first()
    K = ranks[0].K
    for k in range(1, K):
        pairs = [r.synth(k, True) for r in ranks]
        snap_h = [p[0].clone() if p[0] is not None else None for p in pairs]
        snap_t = [p[1].clone() if p[1] is not None else None for p in pairs]
        for i, r in enumerate(ranks):
            r.halo_ana(k, snap_t[i - 1] if i > 0 else None, snap_h[i + 1] if i + 1 < len(ranks) else None)
    exchange([r.synth(0, False) for r in ranks])
    outs = [r.finish() for r in ranks]
    return torch.cat([o[0] for o in outs], dim=2), torch.cat([o[1] for o in outs], dim=2)


class DistExchange:
    """Neighbour exchange over torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""

    def __init__(self, group=None):
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def all_reduce_sums(self, sums):
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=self.group)
        return sums

    def halo(self, head, tail):
        ops, recv_prev, recv_next = [], None, None
        if head is not None:
            recv_prev = torch.empty_like(head)
            send = head.contiguous()
            ops += [dist.P2POp(dist.isend, send, self.rank - 1, self.group), dist.P2POp(dist.irecv, recv_prev, self.rank - 1, self.group)]
        if tail is not None:
            recv_next = torch.empty_like(tail)
            send = tail.contiguous()
            ops += [dist.P2POp(dist.isend, send, self.rank + 1, self.group), dist.P2POp(dist.irecv, recv_next, self.rank + 1, self.group)]
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return recv_prev, recv_next


def run_distributed(rank_state, y_slab, c, xch: DistExchange):
    """One rank of the distributed forward.  Returns (xhat of the owned frames, z of the owned coarse frames)."""
    sums = xch.all_reduce_sums(rank_state.local_sums(y_slab).double())
    rank_state.set_global_sums(sums, c)
    rank_state.first()
    for k in range(1, rank_state.K):
        head, tail = rank_state.synth(k, True)
        rank_state.halo_ana(k, *xch.halo(head, tail))
    head, tail = rank_state.synth(0, False)
    rank_state.add_halo(*xch.halo(head, tail))
    return rank_state.finish()


class NativeComm:
    """libcdl_b200's own NCCL communicator (cdl_comm_create): rank 0 draws the ncclUniqueId and the process group that
    torchrun set up carries its 128 bytes to the other ranks."""

    def __init__(self, rank, world, device, group=None):
        import ctypes
        from . import _lib
        self.lib = _lib.load()
        dev = torch.device(device)
        idt = torch.zeros(128, dtype=torch.uint8, device=dev if dist.get_backend(group) == "nccl" else "cpu")
        if rank == 0:
            buf = ctypes.create_string_buffer(128)
            _lib.check(self.lib.cdl_comm_unique_id(buf), "cdl_comm_unique_id")
            idt.copy_(torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8))
        dist.broadcast(idt, src=0, group=group)
        raw = bytes(idt.cpu().numpy().tobytes())
        self.handle = ctypes.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(self.lib.cdl_comm_create(ctypes.byref(self.handle), raw, rank, world, dev.index or 0), "cdl_comm_create")
        self.rank, self.world = rank, world

    def close(self):
        if getattr(self, "handle", None):
            self.lib.cdl_comm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ShardedVideoDenoiser:
    """One long clip (N,1,D,H,W) over `world` GPUs, one process per GPU (BASELINE config 5).

        den = ShardedVideoDenoiser(net, clip_shape, rank, world, device)
        xhat_owned, z_owned = den(y_slab, sigma)          # y_slab = clip frames den.geo["f0"]:den.geo["f1"]

    The whole forward of a rank runs inside the library (cdl_forward_sharded): K iterations, per-iteration NCCL
    neighbour exchange of the Pd - s seam frames, seam sums fused into the analysis step.  world == 1 runs the same
    entry point without a communicator (the one-GPU point of a strong-scaling curve)."""

    def __init__(self, net, clip_shape, rank, world, device, precision="tf32", group=None):
        from .plan import Plan
        N, C, D, H, W = clip_shape
        P3 = net._plan_P(precision)                  # (7,7,7) when smaller odd filters ride the tensor-core kernels zero-embedded:
        self.geo = slab_geometry(D, P3[0], net.s, world, rank)      # the halos then follow the embedded temporal extent
        g = self.geo
        dev_index = torch.device(device).index or 0
        plan = Plan(3, N, C, net.M, net.K, (g["f1"] - g["f0"], H, W), P3, net.s, precision=precision, device=dev_index,
                    halo_front=g["hf"], halo_back=g["hb"])
        sum_plan = Plan(3, N, C, net.M, 1, (net.s * (g["q1"] - g["q0"]), H, W), P3, net.s, precision="fp32", device=dev_index)
        if plan.precision == "fp32" and P3 != net._P3():     # the tensor-core kernels declined (odd width, ...): native extents
            plan.close()
            P3 = net._P3()
            self.geo = g = slab_geometry(D, P3[0], net.s, world, rank)
            plan = Plan(3, N, C, net.M, net.K, (g["f1"] - g["f0"], H, W), P3, net.s, precision="fp32", device=dev_index,
                        halo_front=g["hf"], halo_back=g["hb"])
        A, B = net._banks_for(plan)
        plan.set_weights(A, B, net.t)
        self.net, self.plan, self.sum_plan = net, plan, sum_plan
        self.state = SlabRank(PlanOps(plan, sum_plan), g, net.K, net.s)
        self.group, self.xch = group, None          # the exchange is created on first use (needs an initialised process group)
        self.rank, self.world = rank, world
        self.comm = None
        self.code = self.r = self.halo = None

    # -- native route ---------------------------------------------------------------------------------
    def _buffers(self):
        if self.code is None:
            self.code = self.plan.new_code()
            self.r = torch.empty(self.plan.fine_shape, dtype=torch.float32, device=self.plan.device)
            self.halo = torch.empty(2 * self.plan.halo_bytes // 4, dtype=torch.float32, device=self.plan.device)
        if self.world > 1 and self.comm is None:
            self.comm = NativeComm(self.rank, self.world, self.plan.device, self.group)

    def forward_resident(self, y_slab, sigma=None, want_z=False):
        """y_slab on the device -> (xhat of the owned frames, z of the owned coarse frames or None)."""
        self._buffers()
        plan, g = self.plan, self.geo
        c = self.net._c_vector(sigma, y_slab.shape[0], y_slab.device)
        owned = y_slab[:, :, g["hf"]:y_slab.shape[2] - g["hb"]]
        sums = self.sum_plan.reduce_sums(owned.contiguous())
        if self.world > 1:
            plan.comm_allreduce(self.comm, sums)
        mean = plan.mean_from_sums(sums)
        yp = plan.center_pad(y_slab.contiguous(), mean)[0]
        plan.forward_sharded(self.comm, yp, c, self.code, self.r, self.halo)
        x = plan.postprocess(self.r, mean)
        xhat = x[:, :, g["hf"]:x.shape[2] - g["hb"]]
        return xhat, (plan.export_code(self.code) if want_z else None)

    def denoise_host(self, y_host_slab, xhat_host_owned, sigma=None):
        """Pinned host slab in, owned frames of xhat out into a pinned host buffer.  Asynchronous: the upload runs on a
        copy stream into one of two device buffers, the forward on the current stream, the download on a second copy stream,
        so that in a sequence of calls (a stream of clips) the copies of neighbouring steps overlap the compute of this one
        - every step still performs its own H2D and D2H.  The output buffer is complete after `den.wait()` (or a device
        synchronisation)."""
        dev = self.plan.device
        cur = torch.cuda.current_stream(dev)
        st = self.__dict__.setdefault("_io", None)
        if st is None or st["shape"] != tuple(y_host_slab.shape):
            st = self.__dict__["_io"] = {
                "shape": tuple(y_host_slab.shape), "h2d": torch.cuda.Stream(device=dev), "d2h": torch.cuda.Stream(device=dev),
                "ybuf": [torch.empty(y_host_slab.shape, dtype=torch.float32, device=dev) for _ in range(2)],
                "free": [None, None], "flip": 0}
        i = st["flip"]
        st["flip"] ^= 1
        with torch.cuda.stream(st["h2d"]):
            if st["free"][i] is not None:
                st["h2d"].wait_event(st["free"][i])          # the forward that last read this buffer (two calls ago) is done
            st["ybuf"][i].copy_(y_host_slab, non_blocking=True)
            up = st["h2d"].record_event()
        cur.wait_event(up)
        xhat, _ = self.forward_resident(st["ybuf"][i], sigma)
        st["free"][i] = cur.record_event()
        st["d2h"].wait_event(st["free"][i])
        with torch.cuda.stream(st["d2h"]):
            xhat_host_owned.copy_(xhat, non_blocking=True)
        xhat.record_stream(st["d2h"])

    def join(self):
        """Make the current stream wait for the downloads of every enqueued denoise_host call (no host blocking)."""
        st = self.__dict__.get("_io")
        if st is not None:
            torch.cuda.current_stream(self.plan.device).wait_stream(st["d2h"])

    def wait(self):
        """Block the host until every enqueued denoise_host call has delivered its output."""
        st = self.__dict__.get("_io")
        if st is not None:
            st["d2h"].synchronize()
        torch.cuda.current_stream(self.plan.device).synchronize()

    def __call__(self, y_slab, sigma=None):
        if y_slab.is_cuda:
            return self.forward_resident(y_slab, sigma, want_z=True)
        c = self.net._c_vector(sigma, y_slab.shape[0], y_slab.device)
        if self.world == 1:
            st = self.state
            st.set_global_sums(st.local_sums(y_slab).double(), c)
            st.first()
            for k in range(1, st.K):
                st.synth(k, True)
                st.ana(k)
            st.synth(0, False)
            return st.finish()
        if self.xch is None:
            self.xch = DistExchange(self.group)
        return run_distributed(self.state, y_slab, c, self.xch)
