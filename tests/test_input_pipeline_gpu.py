"""awgn + Bayer mask + pre_process fused (cdl_preprocess_noisy) against the reference's step-by-step
expression (utils.py:13-55, model/utils.py:5-22,70-87) restated by the oracle - bit-exact for yp / mask_p / the noisy clip
(same roundings), <= 1 ulp for the mean (fp64 sum vs torch's pairwise fp32 sum, as for cdl_preprocess); then the whole
evaluation step (windows.noisy_forward) against oracle.forward_t on the materialised noisy input, and the 16-frame-window
driver on the GPU."""
import numpy as np
import pytest
import torch

import cdl_oracle as O

pytestmark = pytest.mark.gpu


def _bayer(x):
    m = torch.zeros_like(x)
    m[:, 0, 0::2, 0::2] = 1; m[:, 1, 0::2, 1::2] = 1; m[:, 1, 1::2, 0::2] = 1; m[:, 2, 1::2, 1::2] = 1
    return m


@pytest.mark.parametrize("ndim,shape,s,mode", [(2, (2, 3, 33, 46), 2, "bayer"), (2, (1, 3, 40, 64), 1, "bayer"), (2, (2, 1, 31, 30), 2, "none"),
                                               (3, (2, 1, 9, 22, 26), 2, "none"), (3, (1, 1, 8, 20, 24), 2, "tensor"), (2, (1, 3, 24, 28), 1, "nonoise")])
def test_fused_input_pipeline_equals_stepwise(ndim, shape, s, mode):
    import cdlnet_video_b200 as cb
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(11)
    x = torch.rand(*shape, generator=g)
    noise = None if mode == "nonoise" else torch.randn(*shape, generator=g)
    sig = torch.tensor([25.0, 10.0][:shape[0]])
    c = sig / 255.0
    mask = _bayer(x) if mode in ("bayer", "nonoise") else ((torch.rand(*shape, generator=g) > 0.3).float() if mode == "tensor" else None)
    noisy = x if noise is None else x + noise * c.reshape(-1, *([1] * (len(shape) - 1)))       # utils.awgn
    if mask is not None:
        noisy = mask * noisy
    P = (7, 7) if ndim == 2 else (7, 7, 7)
    plan = cb.Plan(ndim, shape[0], shape[1], 8, 2, shape[2:], P, s, has_mask=mask is not None, precision="fp32")
    yp0, mp0, mean0 = plan.preprocess(noisy.to(dev), None if mask is None else mask.to(dev))
    yp, mp, mean, y = plan.preprocess_noisy(x.to(dev), None if noise is None else noise.to(dev), c.to(dev),
                                           mask.to(dev) if mode == "tensor" else None, bayer=mode in ("bayer", "nonoise"), want_y=True)
    assert torch.equal(y.cpu(), noisy)
    assert torch.equal(mean, mean0) and torch.equal(yp, yp0)
    if mask is not None:
        assert torch.equal(mp, mp0)
    ypo, _, _, mo = O.pre_process_t(noisy, s, 1 if mask is None else mask)
    assert (yp.cpu() - ypo).abs().max().item() <= 2.4e-7           # mean within 1 ulp of torch's
    if mask is not None:
        assert torch.equal(mp.cpu(), mo)


def test_noisy_forward_video_and_windows():
    import cdlnet_video_b200 as cb
    from cdlnet_video_b200 import windows
    dev = torch.device("cuda", 0)
    torch.manual_seed(4)
    K, M = 3, 169
    net = cb.CDLNetVideo(K=K, M=M, P=7, s=2, C=1, t0=0.0, adaptive=True, init=False)
    with torch.no_grad():
        for k in range(K):
            net.A[k].weight.mul_(0.006)
            net.B[k].weight.copy_(net.A[k].weight * (1 + 0.05 * torch.randn_like(net.A[k].weight)))
        net.t.copy_(torch.rand_like(net.t) * 0.01)
    x = torch.rand(1, 1, 32, 24, 40)
    noise = torch.randn_like(x)
    noisy = x + noise * (25.0 / 255)
    A, B = [m.weight.detach() for m in net.A], [m.weight.detach() for m in net.B]
    net = net.cuda().eval()
    # one 16-frame window through the fused input pipeline, exact family
    net.precision = "fp32"
    xr, zr, *_ = O.forward_t(noisy[:, :, :16], A, B, net.t.detach().cpu(), 2, 25.0, True, 1)
    xhat, z, y = windows.noisy_forward(net, x[:, :, :16].to(dev), 25.0, noise[:, :, :16].to(dev), want_noisy=True)
    assert torch.equal(y.cpu(), noisy[:, :, :16])
    assert (xhat.cpu() - xr).abs().max().item() <= 2e-5 and (z.cpu() - zr).abs().max().item() <= 2e-5
    # the window driver (analyze3d.py's evaluation): two independent 16-frame windows, default (tensor-core) family
    net.precision = "auto"
    out = windows.denoise_windows(net, noisy.to(dev), 25.0, window=16, batch=2)
    for a in (0, 16):
        xr, *_ = O.forward_t(noisy[:, :, a:a + 16], A, B, net.t.detach().cpu(), 2, 25.0, True, 1)
        assert (out[:, :, a:a + 16].cpu() - xr).abs().max().item() <= 1e-4
