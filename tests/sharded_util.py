"""Oracle-backed slab operators for the CPU tests of the temporal-sharding driver (test infrastructure)."""
import torch
import torch.nn.functional as F

import cdl_oracle as O


class OracleOps:
    """Same interface as cdlnet_video_b200.sharded.PlanOps, computed with the torch-CPU oracle operators."""

    def __init__(self, A, B, t, s, geo, adaptive=True):
        self.A, self.B, self.t, self.s, self.geo = A, B, t, s, geo
        self.P = tuple(A[0].shape[2:])
        self.h = self.P[0] // 2

    def owned_sums(self, y):
        N = y.shape[0]
        out = torch.zeros(2 * N, dtype=torch.float64)
        out[0::2] = y.double().sum(dim=(1, 2, 3, 4))
        out[1::2] = float(y[0].numel())
        return out

    def mean_from_sums(self, sums):
        return (sums[0::2].float() / sums[1::2].float())

    def center_pad(self, y_loc, mean):
        x = y_loc - mean.reshape(-1, 1, 1, 1, 1)
        pad = O.calc_pad_2d(y_loc.shape[3], y_loc.shape[4], self.s)
        self.pad_hw = pad
        return F.pad(x, (*pad, 0, 0), mode="reflect") if any(pad) else x

    def new_code(self):
        return None

    def new_fine(self):
        return None

    def _tau(self, k, c):
        cc = 0 if c is None else c.reshape(-1, 1, 1, 1, 1)
        return self.t[k, :1] + cc * self.t[k, 1:2]

    def analysis(self, k, r, code, c, first):
        g, h, P = self.geo, self.h, self.P
        rp = F.pad(r, (0, 0, 0, 0, h - g["hf"], h - (g["hb"] if g["hb"] else 0) if g["hb"] else h))
        u = F.conv3d(rp, self.A[k], stride=self.s, padding=(0, P[1] // 2, P[2] // 2))
        nq = g["q1"] - g["q0"]
        u = u[:, :, :nq]
        self.code = O.soft_threshold_t(u if first else self.code - u, self._tau(k, c))

    def synthesis(self, k, code, out, yp, residual):
        g, h, P, s = self.geo, self.h, self.P, self.s
        x = F.conv_transpose3d(self.code, self.B[k], stride=s, padding=(0, P[1] // 2, P[2] // 2), output_padding=(0, s - 1, s - 1))
        lo = h - g["hf"]
        D_loc = g["f1"] - g["f0"]
        x = x[:, :, lo:lo + D_loc]
        self.r_out = x - yp if residual else x

    def postprocess(self, xp, mean):
        return O.unpad_2d(xp, self.pad_hw) + mean.reshape(-1, 1, 1, 1, 1)

    def export_code(self, code):
        return self.code


class OracleSlabRank:
    """SlabRank variant whose buffers are produced by the (functional) oracle ops."""

    def __new__(cls, ops, geo, K, s):
        from cdlnet_video_b200.sharded import SlabRank

        class _R(SlabRank):
            def set_global_sums(self, sums, c):
                self.mean = self.ops.mean_from_sums(sums)
                self.yp = self.ops.center_pad(self.y_loc, self.mean)
                self.c = c
                self.code = None
                self.r = None

            def synth(self, k, residual=True):
                self.ops.synthesis(k, None, None, self.yp if residual else None, residual)
                self.r = self.ops.r_out.clone()
                self._residual = residual
                return self._head(self.r), self._tail(self.r)
        return _R(ops, geo, K, s)


def make_problem(seed=0, N=1, M=10, K=3, D=24, H=14, W=18, P=(7, 7, 7), s=2):
    g = torch.Generator().manual_seed(seed)
    y = torch.rand(N, 1, D, H, W, generator=g)
    Wt = torch.randn(M, 1, *P, generator=g) * (0.7 / (2.0 * M * P[0] * P[1] * P[2] / s ** 3) ** 0.5)
    A = [Wt * (1 + 0.1 * torch.randn(Wt.shape, generator=g)) for _ in range(K)]
    B = [Wt * (1 + 0.1 * torch.randn(Wt.shape, generator=g)) for _ in range(K)]
    t = torch.rand(K, 2, M, 1, 1, 1, generator=g) * 0.01
    return y, A, B, t
