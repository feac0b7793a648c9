#!/usr/bin/env python
"""Secondary measurements: the other BASELINE.json configurations (bench.py keeps the driver contract on config 2).

  python bench_configs.py --config cfg1            # CDLNet-s2030, one 256x256 image           (fp32 CUDA-core kernels: s = 2)
  python bench_configs.py --config cfg1b           # CDLNet root args.json, one 256x256 image   (2-D tcgen05 kernels)
  python bench_configs.py --config cfg3            # JDD CDLNet, 32 x 3 x 1024^2 + Bayer mask   (2-D tcgen05 kernels)
  python bench_configs.py --config cfg4            # GDLNet colour, 64 x 3 x 512^2              (2-D tcgen05 kernels)
  (--precision fp32 or CDL_TC2D=0 selects the exact fp32 CUDA-core kernels for the 2-D configurations)
  python bench_configs.py --config cfg5 --frames 240                      # one 1080p clip on 1 GPU (tcgen05 kernels)
  torchrun --nproc-per-node N ... bench_configs.py --config cfg5 --frames 240   # temporally sharded, halo exchange

One JSON line per run: value in Mvoxels/s (pixels x frames of xhat per second), CUDA-event timed, weights random
(spectral scale ~ reference init), inputs synthetic.  cfg5 is STRONG scaling: the clip is fixed, ranks split it.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

import cdlnet_video_b200 as cb
from cdlnet_video_b200 import sharded


def make_net(kind, K, M, P, s, C, seed=1):
    torch.manual_seed(seed)
    if kind == "video":
        net = cb.CDLNetVideo(K=K, M=M, P=P, s=s, C=C, adaptive=True, init=False)
        T, nsp = P ** 3, 3
    elif kind == "gabor":
        net = cb.GDLNet(K=K, M=M, P=P, s=s, C=C, order=1, adaptive=True, init=False)
        T, nsp = P * P, 2
    else:
        net = cb.CDLNet(K=K, M=M, P=P, s=s, C=C, adaptive=True, init=False)
        T, nsp = P * P, 2
    scale = 0.7 / (2.0 * M * T * C / s ** nsp) ** 0.5          # spectral constant of D∘A ~ 1.4-2 x M*T*C/s^d for randn banks
    with torch.no_grad():
        if kind == "gabor":
            for m in list(net.A) + list(net.B):
                m.alpha.mul_(scale * 3)
        else:
            for k in range(K):
                net.A[k].weight.mul_(scale)
                net.B[k].weight.copy_(net.A[k].weight * (1 + 0.03 * torch.randn_like(net.A[k].weight)))
        net.t.copy_(torch.rand_like(net.t) * 0.01)
    return net.eval()


def time_steps(fn, steps, warmup, world):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item() / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg5")
    ap.add_argument("--frames", type=int, default=240)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--precision", default="auto")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    out = {"config": args.config, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "unit": "Mvoxels/s", "data": "synthetic"}

    if args.config == "cfg5":
        K, M, P, s = 30, 169, 7, 2
        net = make_net("video", K, M, P, s, 1).to(dev)
        D, H, W = args.frames, args.height, args.width
        prec = "fp32" if args.precision == "fp32" else "tf32"
        den = sharded.ShardedVideoDenoiser(net, (1, 1, D, H, W), rank, world, dev, precision=prec)
        g = den.geo
        gen = torch.Generator(device=dev).manual_seed(0)      # same clip on every rank; each keeps its slab (+ halos)
        y = torch.rand(1, 1, g["f1"] - g["f0"], H, W, generator=gen, device=dev)
        state = den.state

        def fn():
            xhat, code = run_no_export(den, y)
            return xhat
        ms = time_steps(fn, args.steps, args.warmup, world)
        vox = D * H * W
        flops = 2 * K * 2.0 * (vox / 8) * M * 343
        # per-iteration roofline (north_star): analysis = 2 passes of the 176-subband code over HBM, synthesis = tensor pipe
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            peaks = {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}
        zbytes = (vox / 8) * 176 * 4.0
        kflops = 2.0 * (vox / 8) * M * 343
        tf32 = peaks.get("bf16_tflops_sustained", 1400.0) / 2 * 1e12
        floor_s = K * (max(2 * zbytes / (peaks["hbm_gbs"] * 1e9), kflops / tf32) + max(zbytes / (peaks["hbm_gbs"] * 1e9), kflops / tf32)) / world
        out.update(workload=f"cfg5: CDLNetVideo K=30 M=169 7x7x7 s=2, one {D}x{H}x{W} clip, temporal slabs over {world} GPU(s), "
                            f"halo {g['overlap']} frames per seam per iteration", precision=den.plan.precision, scaling="strong",
                   value=vox / (ms * 1e-3) / 1e6, ms_per_step=ms, tflops=flops / (ms * 1e-3) / 1e12,
                   frac_of_tf32_roofline=flops / (ms * 1e-3) / 1e12 / 688.1 / world,
                   frac_of_per_iteration_roofline=floor_s / (ms * 1e-3))
    else:
        if args.config in ("cfg1", "cfg1b"):
            kind, (K, M, P, s, C), shape, use_mask = "cdl", ((30, 169, 7, 2, 1) if args.config == "cfg1" else (20, 32, 7, 1, 1)), (1, 1, 256, 256), False
        elif args.config == "cfg3":
            kind, (K, M, P, s, C), shape, use_mask = "cdl", (42, 64, 7, 1, 3), (32 // world, 3, 1024, 1024), True
        elif args.config == "cfg4":
            kind, (K, M, P, s, C), shape, use_mask = "gabor", (30, 64, 7, 1, 3), (64 // world, 3, 512, 512), False
        else:
            raise SystemExit("unknown config")
        net = make_net(kind, K, M, P, s, C).to(dev)
        net.precision = args.precision
        y = torch.rand(*shape, device=dev)
        mask = 1
        if use_mask:                                           # RGGB Bayer mask (reference utils.py:13-19)
            mask = torch.zeros_like(y)
            mask[:, 0, 0::2, 0::2] = 1; mask[:, 1, 0::2, 1::2] = 1; mask[:, 1, 1::2, 0::2] = 1; mask[:, 2, 1::2, 1::2] = 1
            y = y * mask
        sigma = 10.0 if use_mask else 25.0

        def fn():
            with torch.no_grad():
                return net(y, sigma, mask=mask)
        ms = time_steps(fn, args.steps, args.warmup, world)
        plan = net._last_plan
        vox = world * shape[0] * shape[2] * shape[3]
        flops = 2 * K * 2.0 * (vox / s ** 2) * M * C * P * P
        # z makes one HBM round trip per iteration; images: preprocess 2 reads, K reads of yp (+ K-1 of the
        # mask), one write of xhat.  2-D configurations are HBM-bound on tensor cores (AI 24-72 FLOP/B).
        try:
            hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception:
            hbm = 6650.0
        bytes_fwd = 8.0 * K * (vox / s ** 2) * M + 4.0 * vox * C * (K + 2 + (K - 1 if use_mask else 0))
        out.update(algorithmic_bytes=bytes_fwd, hbm_floor_ms=bytes_fwd / (hbm * 1e9) * 1e3 / world,
                   frac_of_hbm_roofline=bytes_fwd / (hbm * 1e9) * 1e3 / world / ms)
        out.update(workload=f"{args.config}: {type(net).__name__}(K={K},M={M},P={P},s={s},C={C}) on {world}x{tuple(shape)}"
                            f"{' + Bayer mask' if use_mask else ''}", precision=plan.precision, scaling="weak",
                   value=vox / (ms * 1e-3) / 1e6, ms_per_step=ms, tflops=flops / (ms * 1e-3) / 1e12)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_no_export(den, y_slab):
    """ShardedVideoDenoiser.__call__ without the final code export (z of a 240-frame 1080p clip is 42 GB)."""
    st = den.state
    c = den.net._c_vector(25.0, 1, y_slab.device)
    if den.world > 1 and den.xch is None:
        den.xch = sharded.DistExchange(den.group)
    sums = st.local_sums(y_slab).double()
    if den.world > 1:
        sums = den.xch.all_reduce_sums(sums)
    st.set_global_sums(sums, c)
    st.first()
    for k in range(1, st.K):
        head, tail = st.synth(k, True)
        if den.world > 1:
            st.add_halo(*den.xch.halo(head, tail))
        st.ana(k)
    head, tail = st.synth(0, False)
    if den.world > 1:
        st.add_halo(*den.xch.halo(head, tail))
    x = st.ops.postprocess(st.r, st.mean)
    g = st.geo
    return x[:, :, g["hf"]:x.shape[2] - g["hb"]], st.code


if __name__ == "__main__":
    main()
