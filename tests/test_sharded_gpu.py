"""Temporal slabs on the GPU: the libcdl_b200 kernels with temporal halos (cdl_desc_t.halo_front/back) must
reproduce the unsharded forward.  Single-GPU: all slabs run in lock step in one process (the exchange is a
tensor hand-over).  With >= 2 GPUs: a real NCCL P2P run, one process per GPU."""
import os
import subprocess
import sys

import pytest
import torch

import cdlnet_video_b200 as cb
from cdlnet_video_b200 import sharded
from sharded_util import make_problem

pytestmark = pytest.mark.gpu


def _net(A, B, t, s, P):
    K, M = len(A), A[0].shape[0]
    net = cb.CDLNetVideo(K=K, M=M, P=list(P), s=s, C=1, adaptive=True, init=False)
    with torch.no_grad():
        for k in range(K):
            net.A[k].weight.copy_(A[k]); net.B[k].weight.copy_(B[k])
        net.t.copy_(t)
    return net.cuda().eval()


@pytest.mark.parametrize("prec,world,tol", [("fp32", 2, 2e-5), ("fp32", 3, 2e-5), ("tf32", 2, 1e-4), ("tf32", 3, 1e-4)])
def test_lockstep_slabs_equal_unsharded(prec, world, tol):
    y, A, B, t = make_problem(seed=3, N=1, M=24, K=4, D=36, H=24, W=40)
    net = _net(A, B, t, 2, (7, 7, 7))
    net.precision = prec
    d = torch.device("cuda", 0)
    with torch.no_grad():
        xr, zr = net(y.to(d), 25.0)
    ranks, slabs = [], []
    for r in range(world):
        den = sharded.ShardedVideoDenoiser(net, tuple(y.shape), r, world, d, precision=prec)
        assert den.plan.precision == prec
        ranks.append(den.state)
        slabs.append(y[:, :, den.geo["f0"]:den.geo["f1"]].contiguous().to(d))
    c = torch.full((1,), 25.0 / 255.0, device=d)
    xhat, z = sharded.run_lockstep(ranks, slabs, c)
    assert xhat.shape == xr.shape and z.shape == zr.shape
    ex, ez = (xhat - xr).abs().max().item(), (z - zr).abs().max().item()
    assert ex <= tol and ez <= 2 * tol, (ex, ez)


def test_nccl_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(root, "tests", "sharded_nccl_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "SHARDED_OK" in out.stdout


def test_lockstep_slabs_cfg5_hyperparameters_vs_oracle():
    """BASELINE config 5's network (CDLNetVideo K=30, M=169, 7x7x7, s=2; SURVEY 8(d) weights and thresholds) on a
    64 x 96 x 128 clip split into 4 temporal slabs (16 fine frames each + halos, lock step on one GPU): the sharded
    tcgen05 forward must agree with the unsharded one AND with the CPU oracle within the 1e-4 bar."""
    import bench
    import cdl_oracle as O
    d = torch.device("cuda", 0)
    K, M = bench.CFG["K"], bench.CFG["M"]
    D, H, W = 64, 96, 128
    A, B, u = bench.synthetic_weights(torch, d)
    clean, y = bench.synthetic_clip(torch, 1, seed=5, device=d, shape=(D, H, W))
    net = cb.CDLNetVideo(K=K, M=M, P=7, s=2, C=1, adaptive=True, init=False)
    with torch.no_grad():
        for k in range(K):
            net.A[k].weight.copy_(A[k]); net.B[k].weight.copy_(B[k])
    net = net.cuda().eval()
    net.precision = "tf32"
    t = bench.calibrate_thresholds(torch, A, B, u, y[:, :, :16], d)
    with torch.no_grad():
        net.t.copy_(t.reshape(net.t.shape))
        xu, zu = net(y, bench.SIGMA)
    assert net._last_plan.precision == "tf32"
    world = 4
    ranks, slabs = [], []
    for r in range(world):
        den = sharded.ShardedVideoDenoiser(net, tuple(y.shape), r, world, d, precision="tf32")
        ranks.append(den.state)
        slabs.append(y[:, :, den.geo["f0"]:den.geo["f1"]].contiguous())
    c = torch.full((1,), bench.SIGMA / 255.0, device=d)
    xs, zs = sharded.run_lockstep(ranks, slabs, c)
    xr, zr, *_ = O.forward_t(y.cpu(), [a.cpu() for a in A], [b.cpu() for b in B], t.cpu().reshape(K, 2, M, 1, 1, 1), 2, bench.SIGMA, True, 1)
    e_su = (xs - xu).abs().max().item()
    e_so = (xs.cpu() - xr).abs().max().item()
    e_uo = (xu.cpu() - xr).abs().max().item()
    print(f"cfg5-like {D}x{H}x{W}, 4 slabs: sharded vs unsharded {e_su:.3e}; sharded vs oracle {e_so:.3e}; unsharded vs oracle {e_uo:.3e}; "
          f"nnz {float((zr != 0).float().mean()):.3f}")
    assert e_so <= 1e-4 and e_uo <= 1e-4 and e_su <= 1e-4, (e_su, e_so, e_uo)


def test_denoise_host_pipelined_calls_deliver_each_clip():
    """ShardedVideoDenoiser.denoise_host (the entry bench.py's e2e leg times): uploads / downloads run on copy streams and
    overlap the forward of the neighbouring calls; three different clips enqueued back to back must each come back as their
    own forward (exact fp32 family: bit-identical to forward_resident)."""
    y, A, B, t = make_problem(seed=5, N=1, M=16, K=3, D=12, H=24, W=40)
    net = _net(A, B, t, 2, (7, 7, 7))
    d = torch.device("cuda", 0)
    den = sharded.ShardedVideoDenoiser(net, tuple(y.shape), 0, 1, d, precision="fp32")
    g = torch.Generator().manual_seed(1)
    clips = [y] + [torch.rand(y.shape, generator=g) for _ in range(2)]
    want = [den.forward_resident(c.to(d), 25.0)[0].cpu() for c in clips]
    ins = [c.pin_memory() for c in clips]
    outs = [torch.empty_like(c).pin_memory() for c in clips]
    for a, b in zip(ins, outs):
        den.denoise_host(a, b, 25.0)
    den.wait()
    for got, ref in zip(outs, want):
        assert torch.equal(got, ref)


def test_lockstep_slabs_with_zero_embedded_filters():
    """A video net with filters (5,7,7) on the tensor-core kernels (zero-embedded in 7x7x7, model/net.py::_plan_P): the slab
    geometry follows the embedded temporal extent; 2 slabs in lock step equal the unsharded forward and the oracle."""
    import torch.nn.functional as F
    import cdl_oracle as O
    y, A7, B7, t = make_problem(seed=9, N=1, M=24, K=3, D=24, H=24, W=40)
    A = [a[:, :, 1:6].contiguous() for a in A7]                  # crop the temporal extent 7 -> 5
    B = [b[:, :, 1:6].contiguous() for b in B7]
    net = _net(A, B, t, 2, (5, 7, 7))
    net.precision = "tf32"
    d = torch.device("cuda", 0)
    with torch.no_grad():
        xr, _ = net(y.to(d), 25.0)
    assert tuple(net._last_plan.Pfull) == (7, 7, 7) and net._last_plan.precision == "tf32"
    xo, *_ = O.forward_t(y, A, B, t, 2, 25.0, True, 1)
    assert (xr.cpu() - xo).abs().max().item() <= 1e-4
    ranks, slabs = [], []
    for r in range(2):
        den = sharded.ShardedVideoDenoiser(net, tuple(y.shape), r, 2, d, precision="tf32")
        assert den.plan.precision == "tf32" and tuple(den.plan.Pfull) == (7, 7, 7)
        ranks.append(den.state)
        slabs.append(y[:, :, den.geo["f0"]:den.geo["f1"]].contiguous().to(d))
    xhat, _ = sharded.run_lockstep(ranks, slabs, torch.full((1,), 25.0 / 255.0, device=d))
    assert (xhat - xr).abs().max().item() <= 1e-4
