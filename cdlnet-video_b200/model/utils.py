"""Pre/post-processing helpers with the reference's names and argument meaning
(reference model/utils.py:5-122).  These torch versions serve the autograd / CPU route of the
modules and `power_method`; under `torch.no_grad()` on CUDA the same arithmetic runs inside
libcdl_b200 (csrc/cdl_prepost.cuh).
"""
import torch
import torch.nn.functional as F


def calc_pad_1D(L, M):
    """[lo, hi] so that L+lo+hi is a multiple of M; the odd unit goes to the far side."""
    rem = (-L) % M
    return [rem // 2, rem - rem // 2]


def calc_pad_2D(H, W, M):
    """(left, right, top, bottom)"""
    return (*calc_pad_1D(W, M), *calc_pad_1D(H, M))


def calc_pad_3D(D, H, W, M):
    """(left, right, top, bottom, front, back)"""
    return (*calc_pad_1D(W, M), *calc_pad_1D(H, M), *calc_pad_1D(D, M))


def _crop(x, pad):
    """Inverse of F.pad for a pad tuple ordered innermost axis first."""
    nax = len(pad) // 2
    index = [slice(None)] * (x.dim() - nax)
    for ax in range(nax):                      # ax = 0 is the innermost (W) axis
        lo, hi = pad[2 * ax], pad[2 * ax + 1]
        size = x.shape[x.dim() - 1 - ax]
        index.insert(x.dim() - nax, slice(lo, size - hi))
    return x[tuple(index)]


def unpad(I, pad):
    """Remove 2D stride padding."""
    return _crop(I, tuple(pad)[:4])


def unpad_3d(I, pad):
    """Remove 3D stride padding.  NOTE: the reference's unpad_3d (model/utils.py:110-122) returns an
    empty or uncropped tensor for 4 of the 8 pad-parity classes; this is the evident
    crop, identical to the reference wherever the reference returns the input shape."""
    return _crop(I, tuple(pad)[:6])


def _pre(x, stride, mask, nsp):
    dims = tuple(range(1, x.dim()))
    if torch.is_tensor(mask):
        xmean = x.sum(dim=dims, keepdim=True) / mask.sum(dim=dims, keepdim=True)
    else:
        xmean = x.mean(dim=dims, keepdim=True)
    x = mask * (x - xmean)
    pad = calc_pad_2D(*x.shape[2:], stride) if nsp == 2 else calc_pad_3D(*x.shape[2:], stride)
    if any(pad):
        x = F.pad(x, pad, mode='reflect')
        if torch.is_tensor(mask):
            mask = F.pad(mask, pad, mode='reflect')
    return x, [xmean, pad], mask


def pre_process(x, stride, mask=1):
    """mean-subtract (masked mean if mask is a tensor), mask, reflect-pad to a multiple of stride."""
    return _pre(x, stride, mask, 2)


def pre_process_3d(x, stride, mask=1):
    return _pre(x, stride, mask, 3)


def post_process(x, params):
    pad = params.pop()
    xmean = params.pop()
    return unpad(x, pad) + xmean


def post_process_3d(x, params):
    pad = params.pop()
    xmean = params.pop()
    return unpad_3d(x, pad) + xmean
