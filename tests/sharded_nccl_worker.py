"""torchrun worker for tests/test_sharded_gpu.py::test_nccl_two_gpus (one process per GPU, NCCL P2P halos)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import cdlnet_video_b200 as cb                      # noqa: E402
from cdlnet_video_b200 import sharded              # noqa: E402
from sharded_util import make_problem              # noqa: E402
from test_sharded_gpu import _net                   # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
d = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=d)
y, A, B, t = make_problem(seed=9, N=1, M=24, K=4, D=40, H=24, W=40)
net = _net(A, B, t, 2, (7, 7, 7)).to(d)
for prec, tol in (("fp32", 2e-5), ("tf32", 1e-4)):
    net.precision = prec
    with torch.no_grad():
        xr, zr = net(y.to(d), 25.0)
    den = sharded.ShardedVideoDenoiser(net, tuple(y.shape), rank, world, d, precision=prec)
    g = den.geo
    xhat, z = den(y[:, :, g["f0"]:g["f1"]].contiguous().to(d), 25.0)
    ex = (xhat - xr[:, :, 2 * g["q0"]:2 * g["q1"]]).abs().max()
    ez = (z - zr[:, :, g["q0"]:g["q1"]]).abs().max()
    worst = torch.stack([ex, ez])
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    assert worst[0].item() <= tol and worst[1].item() <= 2 * tol, (prec, worst.tolist())
    if rank == 0:
        print(f"{prec}: max|dxhat|={worst[0].item():.2e} max|dz|={worst[1].item():.2e}")
dist.barrier()
if rank == 0:
    print("SHARDED_OK")
dist.destroy_process_group()
