#!/usr/bin/env python
"""BASELINE config 1 (CDLNet-s2030: 2-D, K=30, M=169, P=7, s=2, C=1, one 256x256 image): forward time of the fp32 CUDA-core
family vs the CDL_EMBED3D route (the same operator on the video tcgen05 kernels through a two-frame embedding)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cdlnet_video_b200 as cb

torch.manual_seed(0)
K, M = 30, 169
net = cb.CDLNet(K=K, M=M, P=7, s=2, C=1, adaptive=True, init=False)
with torch.no_grad():
    for k in range(K):
        net.A[k].weight.mul_(0.7 / (2.0 * M * 49 / 4) ** 0.5)
        net.B[k].weight.copy_(net.A[k].weight * (1 + 0.05 * torch.randn_like(net.A[k].weight)))
    net.t.copy_(torch.rand_like(net.t) * 0.01)
net = net.cuda().eval()
y = torch.rand(1, 1, 256, 256, device="cuda")
out = {}
for name, env, prec in (("fp32", "0", "fp32"), ("embed3d_tf32", "1", "tf32")):
    os.environ["CDL_EMBED3D"] = env
    net.precision = prec
    with torch.no_grad():
        for _ in range(3):
            x, z = net(y, 25.0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            x, z = net(y, 25.0)
        e1.record(); torch.cuda.synchronize()
    out[name] = {"ms_per_forward": e0.elapsed_time(e1) / 10, "xhat": x.clone()}
d = (out["fp32"]["xhat"] - out["embed3d_tf32"]["xhat"]).abs().max().item()
print(json.dumps({"config": "cfg1 CDLNet(K=30,M=169,P=7,s=2,C=1) 1x1x256x256", "fp32_ms": out["fp32"]["ms_per_forward"],
                  "embed3d_tf32_ms": out["embed3d_tf32"]["ms_per_forward"], "max_abs_xhat_diff": d}))
