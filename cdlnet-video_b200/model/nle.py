"""Blind noise-level estimation with the reference's names (reference model/nle.py:9-27).

    s = 255 * model.nle.noise_level(noisy, method="MAD")     # analyze.py / analyze3d.py:118-121
    xhat, _ = net(noisy, s, mask=mask)

On a CUDA fp32 tensor `nle_mad` runs inside libcdl_b200 (csrc/cdl_nle.cuh: one coefficient pass + an exact radix select)
and returns a DEVICE tensor, so sigma never round-trips to the host; there is no fallback on that route.  CPU tensors
evaluate the reference's torch expression.  The "PCA" method of the reference is not on the path and is not provided."""
import ctypes

import torch
import torch.nn.functional as F

from . import wvlt


def noise_level(y, method="MAD", **kwargs):
    if method in [True, "MAD", "wvlt"]:
        return nle_mad(y)
    raise NotImplementedError(f"noise_level: method {method!r} (only the MAD estimator is on the hot path)")


def _frames_as_channels(y):
    """(N,C,D,H,W) -> (N,C*D,H,W).  The reference call site hands 5-D clips to F.conv2d, which raises; the evident intent
    (the HH subband of every frame, one median per clip) is this view."""
    return y.reshape(y.shape[0], -1, *y.shape[-2:]) if y.dim() == 5 else y


def nle_mad(y):
    """Median absolute deviation of the diagonal bior4.4 detail band / 0.6745 -> (N,1,1,1) (5-D input: (N,1,1,1,1))"""
    out_shape = (y.shape[0],) + (1,) * (y.dim() - 1)
    y4 = _frames_as_channels(y)
    if y4.is_cuda and y4.dtype == torch.float32:
        return _nle_mad_native(y4.contiguous()).reshape(out_shape)
    hh = wvlt.filter_bank_2D('bior4.4')[0][3:4].to(y4.device)
    C = y4.shape[1]
    HHy = F.conv2d(y4, torch.cat([hh] * C), stride=2, groups=C)
    return (torch.median(HHy.abs().reshape(y4.shape[0], -1), dim=1)[0] / 0.6745).reshape(out_shape)


def _nle_mad_native(y):
    try:                                     # imported as cdlnet_video_b200.model.nle
        from .. import _lib
    except ImportError:                      # imported as top-level `model.nle`: model.net has registered the package
        from . import net as _net            # noqa: F401
        from cdlnet_video_b200 import _lib
    lib = _lib.load()
    N, C, H, W = y.shape
    need = ctypes.c_size_t()
    _lib.check(lib.cdl_nle_mad_workspace_bytes(N, C, H, W, ctypes.byref(need)), "cdl_nle_mad_workspace_bytes")
    with torch.cuda.device(y.device):
        ws = torch.empty(need.value, dtype=torch.uint8, device=y.device)
        out = torch.empty(N, dtype=torch.float32, device=y.device)
        _lib.check(lib.cdl_nle_mad(y.data_ptr(), N, C, H, W, out.data_ptr(), ws.data_ptr(),
                                   torch.cuda.current_stream().cuda_stream), "cdl_nle_mad")
        ws.record_stream(torch.cuda.current_stream())
    return out
