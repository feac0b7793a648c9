"""Pin the oracle (oracle/cdl_oracle.py) against outputs of the unmodified reference.

The fixtures in tests/golden/ were written by oracle/gen_golden.py from
/root/reference (model/net.py forward / forward_generator, model/utils.py
pre_process*).  Tolerances: the torch back-end must reproduce the reference's own
fp32 arithmetic to 1e-6 (same library, same op order); the independent numpy
direct-form back-end to 2e-5 in fp32 and the fp64 run brackets both.
Index layout (pad, shapes, reflect padding) is bit-exact.
"""
import glob
import os

import numpy as np
import pytest
import torch

import cdl_oracle as O

CASES = ["cdlnet2d_s2", "cdlnet2d_jdd_mask", "video_s2_p777", "video_s2_p995_odd",
         "video_s1_p775_c2", "gdlnet_s2_c3", "cdlnet2d_nonadaptive",
         "cdlnet2d_jdd_s1_w4", "gdlnet_s1_c3"]        # the last two: geometries of the 2-D tensor-core kernels


def load(golden_dir, name):
    d = dict(np.load(os.path.join(golden_dir, name + ".npz")))
    d["s"] = int(d["s"])
    d["adaptive"] = bool(int(d["adaptive"]))
    if "sigma_none" in d:
        d["sigma"] = None
    elif d["sigma"].ndim == 0:
        d["sigma"] = float(d["sigma"])
    return d


def test_all_fixtures_present(golden_dir):
    have = {os.path.basename(p)[:-4] for p in glob.glob(os.path.join(golden_dir, "*.npz"))}
    assert set(CASES) | {"unpad3d_table"} <= have


@pytest.mark.parametrize("name", CASES)
def test_torch_backend_matches_reference(golden_dir, name):
    d = load(golden_dir, name)
    y = torch.from_numpy(d["y"])
    A = [torch.from_numpy(a) for a in d["A"]]
    B = [torch.from_numpy(b) for b in d["B"]]
    t = torch.from_numpy(d["t"])
    sigma = d["sigma"]
    if isinstance(sigma, np.ndarray):
        sigma = torch.from_numpy(sigma)
    mask = torch.from_numpy(d["mask"]) if "mask" in d else 1
    trace = []
    xhat, z, yp, mean, pad = O.forward_t(y, A, B, t, d["s"], sigma, d["adaptive"], mask, trace=trace)
    # index layout: bit-exact
    assert tuple(pad) == tuple(int(v) for v in d["pad"])
    assert xhat.shape == d["xhat"].shape and z.shape == d["z"].shape
    assert np.array_equal(yp.numpy(), d["yp"])
    assert np.array_equal(mean.numpy(), d["mean"])
    # arithmetic: same library => tight
    assert np.abs(xhat.numpy() - d["xhat"]).max() <= 1e-6
    assert np.abs(z.numpy() - d["z"]).max() <= 1e-6
    if "trace" in d:
        for k, zk in enumerate(trace):
            assert np.abs(zk.numpy() - d["trace"][k]).max() <= 1e-6, k


@pytest.mark.parametrize("name", CASES)
def test_numpy_backend_matches_reference(golden_dir, name):
    d = load(golden_dir, name)
    mask = d["mask"] if "mask" in d else 1
    xhat, z, yp, mean, pad = O.forward_np(d["y"], list(d["A"]), list(d["B"]), d["t"], d["s"],
                                          d["sigma"], d["adaptive"], mask)
    assert tuple(pad) == tuple(int(v) for v in d["pad"])
    assert np.array_equal(yp, d["yp"]) or np.abs(yp - d["yp"]).max() <= 1e-6   # mean reduction order differs
    assert xhat.shape == d["xhat"].shape and z.shape == d["z"].shape
    assert np.abs(xhat - d["xhat"]).max() <= 2e-5
    assert np.abs(z - d["z"]).max() <= 2e-5
    # fp64 run of the same restatement brackets the fp32 reference (its own rounding ~1e-7)
    to64 = lambda a: a.astype(np.float64) if isinstance(a, np.ndarray) else a
    sig = to64(d["sigma"]) if isinstance(d["sigma"], np.ndarray) else d["sigma"]
    x64, z64, *_ = O.forward_np(to64(d["y"]), [to64(a) for a in d["A"]], [to64(b) for b in d["B"]],
                                to64(d["t"]), d["s"], sig, d["adaptive"], to64(mask))
    assert np.abs(x64 - d["xhat"]).max() <= 5e-6
    assert np.abs(z64 - d["z"]).max() <= 5e-6


def test_gabor_filter_matches_reference(golden_dir):
    d = load(golden_dir, "gdlnet_s2_c3")
    for k in range(d["A"].shape[0]):
        for side in ("A", "B"):
            f = O.gabor_filter_np(d[side + "_alpha"][k], d[side + "_a"][k], d[side + "_w0"][k], d[side + "_psi"][k], 7)
            assert np.abs(f - d[side][k]).max() <= 2e-6


def test_unpad3d_table(golden_dir):
    """the reference's unpad_3d is wrong in 4 of 8 parity classes; the oracle
    keeps both behaviours and the product follows the evident crop."""
    tab = np.load(os.path.join(golden_dir, "unpad3d_table.npz"))["table"]
    n_bad = 0
    for D, H, W, l, r, t, b, f, k, oD, oH, oW in tab:
        pad = (int(l), int(r), int(t), int(b), int(f), int(k))
        assert O.calc_pad_3d(int(D), int(H), int(W), 2) == pad
        x = np.zeros((1, 1, D + f + k, H + t + b, W + l + r), dtype=np.float32)
        assert O.unpad_3d_reference(x, pad).shape[2:] == (oD, oH, oW)
        assert O.unpad_3d(x, pad).shape[2:] == (D, H, W)
        n_bad += (oD, oH, oW) != (D, H, W)
    assert n_bad == 4


def test_calc_pad_matches_numpy_floor_ceil():
    # model/utils.py:41-43 uses numpy float ceil/floor; integers must agree for all sizes
    for s in (1, 2, 3, 4):
        for L in range(1, 70):
            lo, hi = O.calc_pad_1d(L, s)
            if L % s == 0:
                assert (lo, hi) == (0, 0)
            else:
                diff = np.ceil(L / s) * s - L
                assert (lo, hi) == (int(np.floor(diff / 2)), int(np.ceil(diff / 2)))


def test_adjoint_pair_and_linearity():
    """SURVEY §4 invariants 1-2: <A x, z> = <x, B z> for equal weights; t == 0 -> linear."""
    rng = np.random.default_rng(0)
    for shape, P, s in (((1, 2, 8, 10), (5, 5), 2), ((1, 1, 4, 6, 8), (3, 5, 3), 2), ((1, 1, 5, 7), (7, 7), 1)):
        C = shape[1]
        W = rng.standard_normal((4, C, *P))
        x = rng.standard_normal(shape)
        u = O.analysis_np(x, W, s)
        zz = rng.standard_normal(u.shape)
        assert abs((u * zz).sum() - (x * O.synthesis_np(zz, W, s)).sum()) < 1e-9
