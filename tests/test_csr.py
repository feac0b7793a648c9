"""the frame-recurrent CSR variants (reference model/net.py:229-262 prox_CSR / prox_CSR_f2; CDLNet_CSR.forward
:426-462, CDLNet_CSRf2.forward :525-567).  tests/golden/csr.npz holds outputs of the unmodified reference classes
(oracle/gen_golden.py csr): three consecutive frames through CDLNet_CSR (argscsr.json's geometry family: P = 9, s = 2), and
every neighbour combination through CDLNet_CSRf2 (colour, masked).
CPU: the oracle restatement and the drop-in modules' torch route reproduce the fixture; GPU: the native route (exact fp32
kernels with the CSR proximal operator fused into the analysis epilogue, cdl_analysis_step_csr) against the fixture."""
import os

import numpy as np
import pytest
import torch

import cdl_oracle as O


@pytest.fixture(scope="module")
def csr(golden_dir):
    return {k: v for k, v in np.load(os.path.join(golden_dir, "csr.npz")).items()}


def _T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def _csr_net(d, device="cpu"):
    import cdlnet_video_b200 as cb
    K, M, C, P = d["csr_A"].shape[0], d["csr_A"].shape[1], d["csr_A"].shape[2], d["csr_A"].shape[3]
    net = cb.CDLNet_CSR(K=K, M=M, P=P, s=int(d["csr_s"]), C=C, t0=0, adaptive=True, init=False)
    with torch.no_grad():
        for k in range(K):
            net.A[k].weight.copy_(_T(d["csr_A"][k])); net.B[k].weight.copy_(_T(d["csr_B"][k]))
            net.A2[k].weight.copy_(_T(d["csr_A2"][k])); net.B2[k].weight.copy_(_T(d["csr_B2"][k]))
        net.t.copy_(_T(d["csr_t"])); net.t2.copy_(_T(d["csr_t2"])); net.g.copy_(_T(d["csr_g"]))
    return net.to(device).eval()


def _f2_net(d, device="cpu"):
    import cdlnet_video_b200 as cb
    K, M, C, P = d["f2_A"].shape[0], d["f2_A"].shape[1], d["f2_A"].shape[2], d["f2_A"].shape[3]
    net = cb.CDLNet_CSRf2(K=K, M=M, P=P, s=int(d["f2_s"]), C=C, t0=0, adaptive=True, init=False)
    with torch.no_grad():
        for k in range(K):
            net.A[k].weight.copy_(_T(d["f2_A"][k])); net.B[k].weight.copy_(_T(d["f2_B"][k]))
        net.t.copy_(_T(d["f2_t"])); net.g1.copy_(_T(d["f2_g1"])); net.g2.copy_(_T(d["f2_g2"]))
    return net.to(device).eval()


def test_state_dict_keys_match_the_reference():
    import cdlnet_video_b200 as cb
    net = cb.CDLNet_CSR(K=2, M=4, P=7, s=1, C=1, init=False)
    keys = set(net.state_dict().keys())
    assert {"t", "t2", "g", "D.weight", "A.0.weight", "B.1.weight", "A2.0.weight", "B2.1.weight"} <= keys
    net = cb.CDLNet_CSRf2(K=2, M=4, P=7, s=1, C=1, init=False)
    assert {"t", "g1", "g2", "D.weight", "A.1.weight", "B.0.weight"} <= set(net.state_dict().keys())


def test_oracle_reproduces_reference_csr(csr):
    d = csr
    A, B = [_T(a) for a in d["csr_A"]], [_T(b) for b in d["csr_B"]]
    A2, B2 = [_T(a) for a in d["csr_A2"]], [_T(b) for b in d["csr_B2"]]
    t, t2, g, sig, s = _T(d["csr_t"]), _T(d["csr_t2"]), _T(d["csr_g"]), _T(d["csr_sigma"]), int(d["csr_s"])
    frames = _T(d["csr_frames"])
    x0, z0 = O.forward_csr_t(frames[0], A2, B2, t2, s, B[0], sig)
    x1, z1 = O.forward_csr_t(frames[1], A, B, t, s, B[0], sig, z_prev=z0, g_prev=g)
    x2, z2 = O.forward_csr_t(frames[2], A, B, t, s, B[0], sig, z_prev=z1, g_prev=g)
    for got, want in ((x0, d["csr_x"][0]), (x1, d["csr_x"][1]), (x2, d["csr_x"][2]), (z0, d["csr_z"][0]), (z2, d["csr_z"][2])):
        assert np.abs(got.numpy() - want).max() <= 1e-6


def test_oracle_reproduces_reference_csr_f2(csr):
    d = csr
    A, B = [_T(a) for a in d["f2_A"]], [_T(b) for b in d["f2_B"]]
    t, g1, g2, s = _T(d["f2_t"]), _T(d["f2_g1"]), _T(d["f2_g2"]), int(d["f2_s"])
    y, mask, zp, za = _T(d["f2_y"]), _T(d["f2_mask"]), _T(d["f2_zprev"]), _T(d["f2_zafter"])
    combos = [(None, None), (zp, None), (None, za), (zp, za)]
    for i, (a, b) in enumerate(combos):
        x, z = O.forward_csr_t(y, A, B, t, s, B[0], float(d["f2_sigma"]), mask=mask, z_prev=a, z_after=b, g_prev=g1, g_after=g2)
        assert np.abs(x.numpy() - d["f2_x"][i]).max() <= 1e-6 and np.abs(z.numpy() - d["f2_z"][i]).max() <= 1e-6


def test_modules_torch_route_reproduces_reference(csr):
    d = csr
    net = _csr_net(d)
    frames, sig = _T(d["csr_frames"]), _T(d["csr_sigma"])
    with torch.no_grad():
        x0, z0 = net(frames[0], None, sig)
        x1, z1 = net(frames[1], z0, sig)
    assert np.abs(x0.numpy() - d["csr_x"][0]).max() <= 1e-6 and np.abs(x1.numpy() - d["csr_x"][1]).max() <= 1e-6
    net = _f2_net(d)
    y, mask, zp, za = _T(d["f2_y"]), _T(d["f2_mask"]), _T(d["f2_zprev"]), _T(d["f2_zafter"])
    with torch.no_grad():
        for i, (a, b) in enumerate([(None, None), (zp, None), (None, za), (zp, za)]):
            x, z = net(y, a, b, float(d["f2_sigma"]), mask=mask)
            assert np.abs(x.numpy() - d["f2_x"][i]).max() <= 1e-6


@pytest.mark.gpu
def test_native_csr_matches_reference_fixture(csr):
    d = csr
    dev = torch.device("cuda", 0)
    net = _csr_net(d, dev)
    frames, sig = _T(d["csr_frames"]).to(dev), _T(d["csr_sigma"]).to(dev)
    with torch.no_grad():
        x0, z0 = net(frames[0], None, sig)
        n0 = net._last_plan.launch_count()
        x1, z1 = net(frames[1], z0, sig)
        x2, z2 = net(frames[2], z1, sig)
    assert net._last_plan.precision == "fp32" and net._last_plan.launch_count() > n0          # the native route ran
    for got, want in ((x0, d["csr_x"][0]), (x1, d["csr_x"][1]), (x2, d["csr_x"][2]), (z0, d["csr_z"][0]), (z1, d["csr_z"][1]), (z2, d["csr_z"][2])):
        assert np.abs(got.cpu().numpy() - want).max() <= 2e-5


@pytest.mark.gpu
def test_native_csr_f2_matches_reference_fixture(csr):
    d = csr
    dev = torch.device("cuda", 0)
    net = _f2_net(d, dev)
    y, mask, zp, za = (_T(d[k]).to(dev) for k in ("f2_y", "f2_mask", "f2_zprev", "f2_zafter"))
    with torch.no_grad():
        for i, (a, b) in enumerate([(None, None), (zp, None), (None, za), (zp, za)]):
            x, z = net(y, a, b, float(d["f2_sigma"]), mask=mask)
            assert np.abs(x.cpu().numpy() - d["f2_x"][i]).max() <= 2e-5, i
            assert np.abs(z.cpu().numpy() - d["f2_z"][i]).max() <= 2e-5, i
