"""Temporal sharding driver (cdlnet_video_b200/sharded.py) on CPU: slab geometry, lock-step emulation and a
world_size-2 gloo run, all against the unsharded oracle forward.  The compute is the oracle (test
infrastructure); what is under test is the host logic: slab bounds, halo sizes, exchange, global mean."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cdl_oracle as O
from cdlnet_video_b200 import sharded
from sharded_util import OracleOps, OracleSlabRank, make_problem


def test_slab_geometry():
    for D, Pd, s, world in ((240, 7, 2, 8), (32, 7, 2, 2), (48, 9, 2, 3), (24, 7, 1, 2)):
        prev_f1 = None
        for r in range(world):
            g = sharded.slab_geometry(D, Pd, s, world, r)
            assert g["hf"] == (Pd // 2 if r else 0) and g["hb"] == ((Pd // 2 - s + 1) if r < world - 1 else 0)
            assert g["f0"] == s * g["q0"] - g["hf"] and g["f1"] == s * g["q1"] + g["hb"]
            if prev_f1 is not None:
                assert prev_f1 - g["f0"] == Pd - s == g["overlap"]          # Pd - s shared frames per seam
            prev_f1 = g["f1"]
        assert g["q1"] == D // s
    with pytest.raises(ValueError):
        sharded.slab_geometry(15, 7, 2, 2, 0)


@pytest.mark.parametrize("world,s,P", [(2, 2, (7, 7, 7)), (3, 2, (7, 7, 7)), (2, 1, (7, 7, 5)), (2, 2, (9, 9, 5))])
def test_lockstep_equals_unsharded(world, s, P):
    y, A, B, t = make_problem(seed=world, D=24 if s == 2 else 20, P=P, s=s)
    sigma = 25.0
    xr, zr, *_ = O.forward_t(y, A, B, t, s, sigma, True, 1)
    K = len(A)
    ranks, slabs = [], []
    for r in range(world):
        g = sharded.slab_geometry(y.shape[2], P[0], s, world, r)
        ranks.append(OracleSlabRank(OracleOps(A, B, t, s, g), g, K, s))
        slabs.append(y[:, :, g["f0"]:g["f1"]].contiguous())
    c = torch.full((y.shape[0],), sigma / 255.0)
    xhat, z = sharded.run_lockstep(ranks, slabs, c)
    assert xhat.shape == xr.shape and z.shape == zr.shape
    assert (xhat - xr).abs().max().item() <= 2e-6
    assert (z - zr).abs().max().item() <= 2e-6


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    y, A, B, t = make_problem(seed=5, D=28)
    g = sharded.slab_geometry(y.shape[2], 7, 2, world, rank)
    st = OracleSlabRank(OracleOps(A, B, t, 2, g), g, len(A), 2)
    c = torch.full((y.shape[0],), 20.0 / 255.0)
    xhat, z = sharded.run_distributed(st, y[:, :, g["f0"]:g["f1"]].contiguous(), c, sharded.DistExchange())
    torch.save((xhat, z), os.path.join(out, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_equals_unsharded(tmp_path):
    world, port = 2, 29517 + (os.getpid() % 200)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    y, A, B, t = make_problem(seed=5, D=28)
    xr, zr, *_ = O.forward_t(y, A, B, t, 2, 20.0, True, 1)
    parts = [torch.load(os.path.join(str(tmp_path), f"r{r}.pt")) for r in range(world)]
    xhat = torch.cat([p[0] for p in parts], dim=2)
    z = torch.cat([p[1] for p in parts], dim=2)
    assert (xhat - xr).abs().max().item() <= 2e-6
    assert (z - zr).abs().max().item() <= 2e-6
