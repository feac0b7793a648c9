"""Blind noise level: reference model/nle.py:17-27 `nle_mad`.

CPU: the oracle restatement (oracle/cdl_oracle.py nle_mad_np) and the package's torch route against the fixture that the
reference's own model/nle.py + model/wvlt.py produced (oracle/gen_golden.py nle; pywt stubbed with the baked bior4.4 table).
GPU: cdl_nle_mad (csrc/cdl_nle.cuh) against the oracle - the median is an exact selection, the coefficients differ from
the reference's conv2d only by fp32 summation order (tolerance 2e-6 relative + 2e-7 absolute on sigma_hat: on a smooth image the band is a cancellation)."""
import os

import numpy as np
import pytest
import torch

import cdl_oracle as O


def _cases(golden_dir):
    d = np.load(os.path.join(golden_dir, "nle_mad.npz"))
    return [(d[f"y{i}"], d[f"sigma_hat{i}"], d[f"sigma_true{i}"]) for i in range(int(d["n"]))]


def test_oracle_matches_reference_fixture(golden_dir):
    for y, sh, st in _cases(golden_dir):
        got = O.nle_mad_np(y)
        assert got.shape == sh.shape
        np.testing.assert_allclose(got, sh, rtol=2e-6, atol=2e-7)
        if y.shape[-1] >= 64:                                     # the estimator itself: smooth clean image, large enough -> within 15 %
            assert np.all(np.abs(sh - st) <= 0.15 * st)


def test_torch_route_matches_reference_fixture(golden_dir):
    import cdlnet_video_b200.model.nle as nle
    for y, sh, _ in _cases(golden_dir):
        out = nle.noise_level(torch.from_numpy(y), method="MAD")
        assert tuple(out.shape) == (y.shape[0], 1, 1, 1)
        np.testing.assert_allclose(out.reshape(-1).numpy(), sh, rtol=2e-6, atol=2e-7)


def test_wavelet_bank_is_a_perfect_reconstruction_pair():
    """sanity of the baked bior4.4 table (PyWavelets is not installed here): the CDF 9/7 pair satisfies the biorthogonality
    conditions sum(dec_lo) = sum(rec_lo) = sqrt(2), sum(dec_hi) = 0, <dec_lo, rec_lo shifted by 2k> = delta_k"""
    lo, hi, rlo, rhi = (np.array(v) for v in (O.BIOR44_DEC_LO, O.BIOR44_DEC_HI, O.BIOR44_REC_LO, O.BIOR44_REC_HI))
    assert abs(lo.sum() - np.sqrt(2)) < 1e-10 and abs(rlo.sum() - np.sqrt(2)) < 1e-10 and abs(hi.sum()) < 1e-10 and abs(rhi.sum()) < 1e-10
    L = len(lo)
    for m in range(-4, 5):                                         # sum_k dec_lo[k] rec_lo[L-1-k+2m] = delta_m, likewise for the high-pass pair
        for d, r in ((lo, rlo), (hi, rhi)):
            acc = sum(d[k] * r[L - 1 - k + 2 * m] for k in range(L) if 0 <= L - 1 - k + 2 * m < L)
            assert abs(acc - (1.0 if m == 0 else 0.0)) < 1e-10
        cross = sum(lo[k] * rhi[L - 1 - k + 2 * m] for k in range(L) if 0 <= L - 1 - k + 2 * m < L)
        assert abs(cross) < 1e-10                                  # low-pass analysis is orthogonal to high-pass synthesis
    # the high-pass filters are the alternating-sign mirrors of the opposite low-pass filters (up to a shift)
    assert np.allclose(np.abs(np.trim_zeros(hi)), np.abs(np.trim_zeros(rlo))) and np.allclose(np.abs(np.trim_zeros(rhi)), np.abs(np.trim_zeros(lo)))


def test_short_input_is_rejected():
    import ctypes
    import cdlnet_video_b200 as cb
    lib = cb.load_library()
    need = ctypes.c_size_t()
    assert lib.cdl_nle_mad_workspace_bytes(1, 1, 9, 64, ctypes.byref(need)) == -2       # CDL_ERR_SHAPE: H < 10 (conv2d would raise)
    assert lib.cdl_nle_mad_workspace_bytes(2, 3, 10, 10, ctypes.byref(need)) == 0 and need.value > 0


@pytest.mark.gpu
def test_native_matches_oracle_and_fixture(golden_dir):
    import cdlnet_video_b200.model.nle as nle
    dev = torch.device("cuda", 0)
    for y, sh, _ in _cases(golden_dir):
        out = nle.noise_level(torch.from_numpy(y).to(dev), method="MAD")
        assert out.is_cuda and tuple(out.shape) == (y.shape[0], 1, 1, 1)
        np.testing.assert_allclose(out.reshape(-1).cpu().numpy(), sh, rtol=2e-6, atol=2e-7)
    g = torch.Generator().manual_seed(3)
    for shape in [(2, 3, 257, 300), (1, 1, 10, 10), (4, 1, 31, 1024), (1, 2, 16, 64, 48)]:   # ragged, minimal, wide, video (frames as channels)
        y = torch.rand(*shape, generator=g) * 0.6 + 0.05 * torch.randn(*shape, generator=g)
        out = nle.noise_level(y.to(dev), method="MAD")
        y4 = y.reshape(shape[0], -1, *shape[-2:])
        ref = O.nle_mad_np(y4.numpy())
        np.testing.assert_allclose(out.reshape(-1).cpu().numpy(), ref, rtol=2e-6, atol=2e-7)
        assert tuple(out.shape) == (shape[0],) + (1,) * (len(shape) - 1)


@pytest.mark.gpu
def test_blind_sigma_feeds_the_network_without_a_host_round_trip():
    """analyze.py's blind path: s = 255 * noise_level(noisy); net(noisy, s) - sigma stays a device tensor end to end"""
    import cdlnet_video_b200 as cb
    import cdlnet_video_b200.model.nle as nle
    torch.manual_seed(0)
    net = cb.CDLNet(K=4, M=16, P=7, s=1, C=1, adaptive=True, init=False)
    with torch.no_grad():
        for k in range(4):
            net.A[k].weight.mul_(0.02); net.B[k].weight.copy_(net.A[k].weight)
        net.t.copy_(torch.rand_like(net.t) * 0.01)
    y = torch.rand(2, 1, 64, 64) + (25 / 255.0) * torch.randn(2, 1, 64, 64)
    s_cpu = 255 * nle.noise_level(y)
    xr, _, *_ = O.forward_t(y, [m.weight.detach() for m in net.A], [m.weight.detach() for m in net.B], net.t.detach(), 1, s_cpu, True, 1)
    net = net.cuda().eval(); net.precision = "fp32"
    with torch.no_grad():
        s = 255 * nle.noise_level(y.cuda())
        assert s.is_cuda
        xhat, _ = net(y.cuda(), s)
    assert (xhat.cpu() - xr).abs().max().item() <= 2e-5
