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
