"""16-frame-window evaluation (cdlnet-video_b200/windows.py; reference analyze3d.py:100-128) on the stock CPU route:
batching windows along the batch axis gives exactly what denoising each window on its own gives."""
import pytest
import torch

import cdlnet_video_b200 as cb
from cdlnet_video_b200.windows import denoise_windows, split_windows


def test_split_windows():
    assert split_windows(32, 16) == [(0, 16), (16, 32)]
    assert split_windows(40, 16) == [(0, 16), (16, 32), (32, 40)]
    assert split_windows(7, 16) == [(0, 7)]
    with pytest.raises(ValueError):
        split_windows(8, 0)


@pytest.mark.parametrize("D,use_mask,per_sample", [(8, False, False), (10, True, True)])
def test_windows_equal_independent_calls(D, use_mask, per_sample):
    torch.manual_seed(D)
    net = cb.CDLNetVideo(K=2, M=4, P=3, s=2, C=1, adaptive=True, init=False).eval()
    with torch.no_grad():
        for k in range(2):
            net.A[k].weight.mul_(0.1); net.B[k].weight.mul_(0.1)
        net.t.copy_(torch.rand_like(net.t) * 0.01)
    N, W = 2, 4
    clip = torch.rand(N, 1, D, 6, 8)
    mask = (torch.rand_like(clip) < 0.7).float() if use_mask else 1
    sigma = torch.tensor([10.0, 30.0]) if per_sample else 25.0
    out = denoise_windows(net, clip, sigma, mask=mask, window=W, batch=3)
    assert out.shape == clip.shape
    for a, b in split_windows(D, W):
        for n in range(N):
            m = mask[n:n + 1, :, a:b] if use_mask else 1
            s = sigma[n:n + 1].reshape(1, 1, 1, 1, 1) if per_sample else sigma
            with torch.no_grad():
                ref, _ = net(clip[n:n + 1, :, a:b], s, mask=m)
            assert torch.allclose(out[n:n + 1, :, a:b], ref, atol=1e-6), (a, b, n)


def test_noisy_forward_has_no_cpu_route():
    """windows.noisy_forward is the fused input pipeline of the CUDA path (cdl_preprocess_noisy): a CPU tensor must fail
    loudly, not fall back (the modules' own forward keeps the stock torch route for CPU tensors)."""
    import pytest
    import torch
    import cdlnet_video_b200 as cb
    from cdlnet_video_b200 import windows
    net = cb.CDLNetVideo(K=2, M=4, P=7, s=2, C=1, adaptive=True, init=False).eval()
    with pytest.raises(RuntimeError, match="CUDA fp32"):
        windows.noisy_forward(net, torch.rand(1, 1, 4, 8, 8), 25.0, torch.randn(1, 1, 4, 8, 8))
