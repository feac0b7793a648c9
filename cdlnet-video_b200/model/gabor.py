"""Gabor-parameterised filter bank (reference model/gabor.py:7-67).

`get_filter()` synthesises the (M, C, ks, ks) bank from (alpha, a, w0, psi); `.T(x)` is the analysis
(strided cross-correlation) and `forward(x)` the synthesis (its adjoint).  Under no-grad CUDA
inference GDLNet feeds the synthesised banks to libcdl_b200; these torch ops are the autograd route.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def gabor_kernel(a, w0, psi, ks):
    """a, w0: (B, O, I, 2); psi: (B, O, I)  ->  (B, O, I, ks, ks):
    exp(-|a*(x-x0)|^2) * cos(w0.(x-x0) + psi) on an ij-indexed ks x ks grid centred at (ks-1)/2."""
    idx = torch.arange(ks, device=a.device)
    grid = torch.stack(torch.meshgrid(idx, idx, indexing='ij'), dim=-1) - (ks - 1) / 2      # (ks,ks,2)
    grid = grid.to(a.dtype)
    ax = a[:, :, :, None, None, :] * grid
    envelope = torch.exp(-(ax * ax).sum(dim=-1))
    phase = (w0[:, :, :, None, None, :] * grid).sum(dim=-1) + psi[:, :, :, None, None]
    return envelope * torch.cos(phase)


class ConvAdjoint2dGabor(nn.Module):
    """2D convolution pair with a mixture-of-Gabor kernel; nic = subbands (M), noc = image channels (C)."""

    def __init__(self, nic, noc, ks, stride=2, order=1):
        super().__init__()
        self.alpha = nn.Parameter(torch.randn(order, nic, noc, 1, 1))
        self.a = nn.Parameter(torch.randn(order, nic, noc, 2))
        self.w0 = nn.Parameter(torch.randn(order, nic, noc, 2))
        self.psi = nn.Parameter(torch.randn(order, nic, noc))
        self.order = order
        self.stride = stride
        self.ks = ks
        p = (ks - 1) // 2
        self._pad = (p, p, p, p)

    def get_filter(self, transpose=False):
        # negating (w0, psi) is a numerical no-op (cos is even); kept for interface parity
        w0, psi = (-self.w0, -self.psi) if transpose else (self.w0, self.psi)
        return (self.alpha * gabor_kernel(self.a, w0, psi, self.ks)).sum(dim=0)

    @property
    def weight(self):
        """The synthesised (M, C, ks, ks) bank, so callers that read `.weight` keep working."""
        return self.get_filter()

    def T(self, x):
        return F.conv2d(F.pad(x, self._pad), self.get_filter(transpose=True), stride=self.stride)

    def forward(self, x):
        return F.conv_transpose2d(x, self.get_filter(), padding=self._pad[0], stride=self.stride,
                                  output_padding=self.stride - 1)
