#!/usr/bin/env python
"""Experimental 2-D tcgen05 analysis kernel (CDL_TC2D=1) against the exact fp32 CUDA-core kernels on the 2-D BASELINE
configurations: whole forward and the analysis step alone, CUDA events; max|xhat(tc2) - xhat(fp32)| on the same inputs.
One JSON line per configuration into gpurun_out/tc2_bench.jsonl."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

import bench_configs as bc

CFG = {"cfg1b": ("cdl", (20, 32, 7, 1, 1), (1, 1, 256, 256), False),
       "cfg3": ("cdl", (42, 64, 7, 1, 3), (32, 3, 1024, 1024), True),
       "cfg4": ("gabor", (30, 64, 7, 1, 3), (64, 3, 512, 512), False)}


def timed(fn, steps=2, warmup=1):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, out


def main():
    names = sys.argv[1:] or ["cfg1b", "cfg4", "cfg3"]
    dev = torch.device("cuda", 0)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "tc2_bench.jsonl"), "a")
    for name in names:
        kind, (K, M, P, s, C), shape, use_mask = CFG[name]
        net = bc.make_net(kind, K, M, P, s, C).to(dev)
        y = torch.rand(*shape, device=dev)
        mask = 1
        if use_mask:
            mask = torch.zeros_like(y)
            mask[:, 0, 0::2, 0::2] = 1; mask[:, 1, 0::2, 1::2] = 1; mask[:, 1, 1::2, 0::2] = 1; mask[:, 2, 1::2, 1::2] = 1
            y = y * mask
        sigma = 10.0 if use_mask else 25.0
        res = {"config": name, "workload": f"{type(net).__name__}(K={K},M={M},P={P},s={s},C={C}) on {tuple(shape)}"}
        outs = {}
        # arms: fp32 = exact CUDA-core kernels (CDL_TC2D=0), tc2 = tensor-core analysis + residual synthesis (the default),
        # tc2fm = the same with the JDD mask applied inside the footprint flush (CDL_TC2D_MASKPASS=0),
        # tc2x3 = the 3-term (hi/lo) analysis of cdl_tc2_analysis_x3.cuh (precision "tf32x3")
        arms = os.environ.get("TC2_ARMS", "fp32,tc2").split(",")
        for tag in arms:
            net.__dict__.pop("_plans", None)
            os.environ["CDL_TC2D"] = "0" if tag == "fp32" else "2"
            os.environ.pop("CDL_TC2D_MASKPASS", None)
            if tag == "tc2fm":
                os.environ["CDL_TC2D_MASKPASS"] = "0"
            net.precision = "fp32" if tag == "fp32" else ("tf32x3" if tag == "tc2x3" else "tf32")

            def fwd():
                with torch.no_grad():
                    return net(y, sigma, mask=mask)
            ms, (xhat, z) = timed(fwd)
            plan = net._last_plan
            res[f"{tag}_precision"] = plan.precision
            res[f"{tag}_forward_ms"] = ms
            outs[tag] = (xhat, (z != 0).float().mean().item())
            # the analysis step alone, on the final code and a residual-sized random image
            r = torch.randn(plan.fine_shape, device=dev) * 0.05
            c = torch.full((shape[0],), sigma / 255.0, device=dev)
            zz = z.clone()
            ams, _ = timed(lambda: plan.analysis_step(1, r, zz, c=c, first=False), steps=5, warmup=2)
            res[f"{tag}_analysis_ms"] = ams
            res[f"{tag}_analysis_GBs"] = 2 * z.numel() * 4 / (ams * 1e-3) / 1e9
            yp = torch.randn(plan.fine_shape, device=dev) * 0.05
            mp = (torch.rand(plan.fine_shape, device=dev) < 0.5).float() if use_mask else None
            sms, _ = timed(lambda: plan.synthesis_step(1, zz, r, yp=yp, mask_p=mp, residual=True), steps=5, warmup=2)
            res[f"{tag}_synthesis_ms"] = sms
            res[f"{tag}_synthesis_GBs"] = z.numel() * 4 / (sms * 1e-3) / 1e9
            del zz, r, yp, mp
        os.environ.pop("CDL_TC2D", None)
        os.environ.pop("CDL_TC2D_MASKPASS", None)
        vox = shape[0] * shape[2] * shape[3]
        for tag in arms:
            res[f"{tag}_Mpix_s"] = vox / (res[f"{tag}_forward_ms"] * 1e-3) / 1e6
            if tag != arms[0]:
                res[f"max_abs_xhat_{tag}_vs_{arms[0]}"] = (outs[tag][0] - outs[arms[0]][0]).abs().max().item()
        res["z_nonzero_frac"] = outs[arms[0]][1]
        print(json.dumps(res), flush=True)
        log.write(json.dumps(res) + "\n"); log.flush()
        del net, y, mask, outs
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
