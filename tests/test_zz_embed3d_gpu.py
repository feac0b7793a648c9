"""The 2-D stride-2 grayscale network (CDLNet-s2030 hyper-parameters, BASELINE config 1) on the video tensor-core kernels
through the two-frame embedding (model/net.py::_forward_embedded3d; default for precision "tf32" / "auto", CDL_EMBED3D=0
disables).  Validated on hardware in round 2 (profiles/r02o_embed3d_tests.log: 2.3e-5; r02q_cfg1_embed3d_timing.json:
1.01 ms vs 5.68 ms on the fp32 kernels).  Bar: max|xhat - oracle| <= 1e-4 (north_star), and the reference's golden vector."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(net, y, sigma):
    os.environ["CDL_EMBED3D"] = "1"
    try:
        with torch.no_grad():
            xhat, z = net(y, sigma)
        torch.cuda.synchronize()
    finally:
        os.environ.pop("CDL_EMBED3D", None)
    assert any(k[1] == "embed3d" for k in net._plans), "embedded route not taken"
    return xhat, z


def test_config1_like_vs_oracle():
    import cdl_oracle as O
    import cdlnet_video_b200 as cb
    torch.manual_seed(5)
    K, M = 30, 169
    net = cb.CDLNet(K=K, M=M, P=7, s=2, C=1, adaptive=True, init=False)
    with torch.no_grad():
        for k in range(K):
            net.A[k].weight.mul_(0.7 / (2.0 * M * 49 / 4) ** 0.5)
            net.B[k].weight.copy_(net.A[k].weight * (1 + 0.05 * torch.randn_like(net.A[k].weight)))
        net.t.copy_(torch.rand_like(net.t) * 0.01)
    y = torch.rand(1, 1, 96, 128)
    xr, zr, *_ = O.forward_t(y, [m.weight.detach() for m in net.A], [m.weight.detach() for m in net.B], net.t.detach(), 2, 25.0, True, 1)
    net = net.cuda().eval()
    net.precision = "tf32"
    xhat, z = _run(net, y.cuda(), 25.0)
    assert tuple(z.shape) == tuple(zr.shape) and tuple(xhat.shape) == tuple(xr.shape)
    ex = (xhat.cpu() - xr).abs().max().item()
    print(f"embed3d: max|xhat-oracle|={ex:.3e} max|z-oracle|={(z.cpu() - zr).abs().max().item():.3e}")
    assert ex <= 1e-4, ex


def test_golden_vector_odd_size():
    """cdlnet2d_s2: 33 x 30 image (both axes need stride padding; padded width 30 is not a multiple of 4 -> the route must
    decline and the ordinary 2-D path must answer)."""
    from util import case_inputs, load_case, module_from_case
    d = load_case("cdlnet2d_s2")
    net = module_from_case(d, "cdlnet2d_s2").cuda()
    net.precision = "tf32"
    y, sigma, mask = case_inputs(d, torch.device("cuda", 0))
    os.environ["CDL_EMBED3D"] = "1"
    try:
        with torch.no_grad():
            xhat, z = net(y, sigma)
    finally:
        os.environ.pop("CDL_EMBED3D", None)
    assert not any(k[1] == "embed3d" for k in net._plans)
    assert np.abs(xhat.cpu().numpy() - d["xhat"]).max() <= 2e-5
