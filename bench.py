#!/usr/bin/env python
"""bench.py — denoised Mvoxels/s of the K-iteration CDLNet-3D forward pass (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N ...            # the reference's CPU path (oracle port) on host cores

Workload (config 2 of BASELINE.json): CDLNetVideo(args3d.json: K=30, M=169, P=7 -> 7x7x7, s=2, C=1,
adaptive) blind-denoising synthetic 16x256x256 grayscale clips at sigma=25.  One step = one forward
pass (mean/pad preprocess + K ISTA iterations + D z + crop) over a batch of `--clips` clips per GPU;
the batch makes the sparse code (88.6 MB per clip) larger than the 126 MB L2.  Multi-GPU: one process
per GPU, clips are independent units sharded across ranks with no data-path collective (weak scaling).

`value` is timed with CUDA events with the clips resident in HBM; `e2e` is the same pass through the
C ABI's host-buffer entry (cdl_denoise_host): pinned host clip -> H2D -> forward -> D2H of xhat.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(K=30, M=169, P=7, s=2, C=1)          # /root/reference args3d.json:3-12 (scalar P = cubic)
CLIP = (16, 256, 256)
SIGMA = 25.0
L_SPECTRAL = 1.375e4                            # reference test.ipynb:171, power-method constant of this filter family


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--clips", type=int, default=4, help="clips per GPU per step")
    ap.add_argument("--precision", default=os.environ.get("CDL_PRECISION", "auto"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-breakdown", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# synthetic inputs and weights (no datasets / checkpoints exist: SURVEY F6)
# ------------------------------------------------------------------------------------------------
def synthetic_clip(torch, n, seed, device):
    """Random signed sums of 2-10 sin/cos plane waves on [-pi,pi]^3, min-max normalised to [0,1]
    (the reference's own synthetic-clip definition, syn_data/gen.py:12-31) + AWGN sigma/255 (utils.py:44-55)."""
    import math
    import random
    D, H, W = CLIP
    rng = random.Random(seed)
    gz = torch.linspace(-math.pi, math.pi, D, device=device).view(D, 1, 1)
    gx = torch.linspace(-math.pi, math.pi, H, device=device).view(1, H, 1)
    gy = torch.linspace(-math.pi, math.pi, W, device=device).view(1, 1, W)
    clips = []
    for _ in range(n):
        acc = None
        for i in range(rng.randint(2, 10)):
            cx, cy, cz = rng.uniform(-5, 5), rng.uniform(-5, 5), rng.uniform(-5, 5)
            f = rng.choice([torch.sin, torch.cos])
            term = f(cx * gx + cy * gy + cz * gz)
            acc = term if acc is None else (acc + term if rng.choice([True, False]) else acc - term)
        acc = (acc - acc.min()) / (acc.max() - acc.min())
        clips.append(acc)
    clean = torch.stack(clips).unsqueeze(1).float()
    g = torch.Generator(device=device).manual_seed(seed)
    noisy = clean + torch.randn(clean.shape, generator=g, device=device) * (SIGMA / 255.0)
    return clean, noisy


def synthetic_weights(torch, device, seed=1):
    """SURVEY 8(d) protocol: one randn bank / sqrt(L), 3 % per-layer perturbation; thresholds are set
    afterwards from the 85th percentile of |A_0 yp| (set_thresholds)."""
    K, M, C, P = CFG["K"], CFG["M"], CFG["C"], CFG["P"]
    g = torch.Generator().manual_seed(seed)
    W = torch.randn(M, C, P, P, P, generator=g) / (L_SPECTRAL ** 0.5)
    A = [(W * (1 + 0.03 * torch.randn(W.shape, generator=g))).to(device) for _ in range(K)]
    B = [(W * (1 + 0.03 * torch.randn(W.shape, generator=g))).to(device) for _ in range(K)]
    u = 0.8 + 0.4 * torch.rand(K, M, generator=g)
    return A, B, u.to(device)


def thresholds_from_quantile(torch, q, u):
    K, M = u.shape
    t = torch.zeros(K, 2, M, device=u.device)
    t[:, 0] = 0.3 * q[None] * u
    t[:, 1] = 0.7 * q[None] * u / (SIGMA / 255.0)
    return t


# ------------------------------------------------------------------------------------------------
# algorithmic work
# ------------------------------------------------------------------------------------------------
def algorithmic(n_clips):
    K, M, C, P, s = CFG["K"], CFG["M"], CFG["C"], CFG["P"], CFG["s"]
    V = CLIP[0] * CLIP[1] * CLIP[2]
    Q = V // s ** 3
    T = P ** 3
    conv_flops = 2.0 * n_clips * Q * M * C * T                 # one analysis or one synthesis
    flops_fwd = 2 * K * conv_flops
    bytes_fwd = 8.0 * K * n_clips * Q * M + 4.0 * n_clips * V * C * (K + 2)
    z_pass = 4.0 * n_clips * Q * M
    return dict(V=V, Q=Q, conv_flops=conv_flops, flops_fwd=flops_fwd, bytes_fwd=bytes_fwd, z_pass=z_pass)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "nvidia-smi unavailable"}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        sm.sort()
        # "under load": the upper half of the samples (idle samples before/after the region drag the median down)
        load = sm[len(sm) // 2:] if sm else []
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU reference arm: the oracle port of the reference's torch-CPU path
# ------------------------------------------------------------------------------------------------
def cpu_reference(steps, warmup):
    """Times oracle.forward_t (the reference's arithmetic: torch CPU conv3d / conv_transpose3d, fp32) on the host
    cores.  Sample = one 16xHxW clip with full K=30; H=W=256 unless the projected run would exceed the budget."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cdl_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    K, M, C, P, s = CFG["K"], CFG["M"], CFG["C"], CFG["P"], CFG["s"]
    g = torch.Generator().manual_seed(1)
    W = torch.randn(M, C, P, P, P, generator=g) / (L_SPECTRAL ** 0.5)
    A = [W * (1 + 0.03 * torch.randn(W.shape, generator=g)) for _ in range(K)]
    B = [W * (1 + 0.03 * torch.randn(W.shape, generator=g)) for _ in range(K)]
    t = torch.zeros(K, 2, M, 1, 1, 1)
    t[:, 0] = 0.012
    t[:, 1] = 0.28

    def run(hw, k_iters):
        y = torch.rand(1, 1, CLIP[0], hw, hw, generator=g)
        t0 = time.perf_counter()
        with torch.no_grad():
            O.forward_t(y, A[:k_iters], B[:k_iters], t[:k_iters], s, SIGMA, True, 1)
        return time.perf_counter() - t0

    probe = run(64, 3)                                          # warms the thread pool; projects the full-size cost
    est_full = probe * (256 * 256) / (64 * 64) * (K / 3.0)
    n = max(1, steps + warmup)
    hw = 256
    while hw > 64 and est_full * (hw * hw) / (256 * 256) > 120.0 / n:   # whole run within ~2 minutes
        hw //= 2
    times = []
    for i in range(n):
        dt = run(hw, K)
        if i >= warmup or n == 1:
            times.append(dt)
        if sum(times) > 150.0 and len(times) >= 1:            # hard wall: keep the whole run within minutes
            break
    vox = CLIP[0] * hw * hw
    dt = sum(times) / len(times)
    return dict(value=vox / dt / 1e6, ms=dt * 1e3, cores=cores, kind="port",
                sample=f"1 clip {CLIP[0]}x{hw}x{hw}, full K={K}, fp32 torch-CPU oracle (oracle/cdl_oracle.py forward_t), "
                       f"{len(times)} timed pass(es)")


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    workload = (f"cfg2: CDLNetVideo(args3d.json) K={CFG['K']} M={CFG['M']} P=7x7x7 s={CFG['s']} C=1 adaptive, "
                f"{args.clips} clip(s) x {CLIP[0]}x{CLIP[1]}x{CLIP[2]} per GPU per step, sigma={SIGMA:g}")

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference(args.steps, args.warmup)
        line = {"impl": "reference", "metric": "denoised Mvoxels/s (K-iter CDLNet-3D fwd)", "value": r["value"], "unit": "Mvoxels/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload, "reference_sample": r["sample"]},
                "cpu_baseline": {"value": r["value"], "unit": "Mvoxels/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": "Mvoxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return

    import torch
    import torch.distributed as dist
    import cdlnet_video_b200 as cb
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n_clips = args.clips
    alg = algorithmic(n_clips)
    prec = "fp32" if args.precision == "fp32" else "tf32"
    plan = cb.Plan(3, n_clips, CFG["C"], CFG["M"], CFG["K"], CLIP, (CFG["P"],) * 3, CFG["s"], has_mask=False,
                   precision=prec, device=local)
    A, B, u = synthetic_weights(torch, dev)
    clean, y = synthetic_clip(torch, n_clips, seed=rank, device=dev)
    c = torch.full((n_clips,), SIGMA / 255.0, dtype=torch.float32, device=dev)
    # thresholds from the 85th percentile of |A_0 yp|, using the library's own kernels
    plan.set_weights(A, B, torch.zeros(CFG["K"], 2, CFG["M"], device=dev))
    yp, _, mean = plan.preprocess(y)
    z0 = plan.new_code()
    plan.analysis_step(0, yp, z0, None, first=True)
    z0 = plan.export_code(z0)
    q = torch.quantile(z0[0].abs().reshape(CFG["M"], -1)[:, ::8].float(), 0.85, dim=1)
    t = thresholds_from_quantile(torch, q, u)
    plan.set_weights(A, B, t)
    del z0, yp
    torch.cuda.synchronize()

    z = torch.empty(plan.z_shape, device=dev)
    stream = torch.cuda.current_stream()

    def step():
        return plan.denoise(y, None, c, z_out=z)

    for _ in range(max(args.warmup, 3) if args.warmup else 0):
        xhat, _ = step()
    torch.cuda.synchronize()
    nnz = float((z != 0).float().mean())
    psnr_in = float(-10 * torch.log10(((y - clean) ** 2).mean()))
    psnr_out = float(-10 * torch.log10(((xhat - clean) ** 2).mean()))

    # ---- timed region: device-resident -----------------------------------------------------------
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = plan.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = plan.launch_count() - l0
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    tms = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_step = tms.item() / args.steps
    value = world * n_clips * alg["V"] / (ms_step * 1e-3) / 1e6

    # ---- e2e: pinned host buffers through cdl_denoise_host ------------------------------------------
    y_host = y.cpu().pin_memory()
    x_host = torch.empty_like(y_host).pin_memory()
    c_host = c.cpu().pin_memory()
    for _ in range(2):
        plan.denoise_host(y_host, x_host, None, c_host, None)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0.record(stream)
    for _ in range(args.steps):
        plan.denoise_host(y_host, x_host, None, c_host, None)
    e1.record(stream)
    torch.cuda.synchronize()
    te = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = te.item() / args.steps
    e2e_value = world * n_clips * alg["V"] / (e2e_ms * 1e-3) / 1e6
    e2e_diff = float((x_host.to(dev) - xhat).abs().max())     # scatter-add order: runs differ at the 1e-5 level (both within 1e-4 of the oracle)

    # ---- per-kernel breakdown (CUDA events on the launch stream) -> roofline of the dominant kernel ----
    roof = None
    if rank == 0 and not args.no_breakdown:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            peak_src = "measured (MEASURED_PEAKS.json)"
        except Exception:
            peaks = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
            peak_src = "fallback (B200_PROFILING.md)"
        ypb, _, meanb = plan.preprocess(y)
        r = torch.empty_like(ypb)
        code = plan.new_code()
        acc = {"analysis": [], "synthesis": []}

        def timed(kind, fn):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); fn(); b.record(stream)
            acc[kind].append((a, b))

        timed("analysis", lambda: plan.analysis_step(0, ypb, code, c, first=True))
        for k in range(1, CFG["K"]):
            timed("synthesis", lambda: plan.synthesis_step(k, code, r, ypb, None, residual=True))
            timed("analysis", lambda: plan.analysis_step(k, r, code, c))
        timed("synthesis", lambda: plan.synthesis_step(0, code, r, residual=False))
        torch.cuda.synchronize()
        tk = {k: [a.elapsed_time(b) for a, b in v] for k, v in acc.items()}
        tot = {k: sum(v) for k, v in tk.items()}
        dom = max(tot, key=tot.get)
        avg_ms = tot[dom] / len(tk[dom])
        eff = plan.precision
        tflops = alg["conv_flops"] / (avg_ms * 1e-3) / 1e12
        gbs = 2 * alg["z_pass"] / (avg_ms * 1e-3) / 1e9 if dom == "analysis" else alg["z_pass"] / (avg_ms * 1e-3) / 1e9
        tf32_peak = peaks.get("bf16_tflops_sustained", 1400.0) / 2.0
        traffic = None
        try:                                     # DRAM bytes per launch of this kernel from the committed ncu --set full capture
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            if eff == "tf32" and tj.get("clips") == n_clips:
                traffic = tj[dom]["dram_bytes_per_launch"]
        except Exception:
            pass
        roof = {"kernel": f"{dom} ({'tcgen05 tf32' if eff == 'tf32' else 'CUDA-core fp32'})", "bound": "tensor",
                "achieved": tflops, "peak": tf32_peak, "unit": "TFLOP/s", "frac": tflops / tf32_peak,
                "traffic": traffic,
                "peak_source": f"{peak_src}: bf16 sustained / 2 (tf32 is not measured separately; cfg-2 AI 167 FLOP/B is tensor-bound for tf32)",
                "hbm": {"achieved": gbs, "peak": peaks.get("hbm_gbs"), "unit": "GB/s", "frac": gbs / peaks.get("hbm_gbs", 6650.0)},
                "avg_launch_ms": avg_ms, "launches_per_step": len(tk[dom]),
                "share_of_step": tot[dom] / sum(tot.values()),
                "per_kernel_ms_per_step": tot, "algorithmic_flops_per_launch": alg["conv_flops"]}
        # both kernels against both ceilings: the analysis step (reads AND rewrites the code: AI 82 FLOP/B) is HBM-bound,
        # the synthesis step (reads it once: AI 165 FLOP/B) is tensor-bound
        roof["kernels"] = {}
        for kname in tot:
            ms = tot[kname] / len(tk[kname])
            zb = (2 if kname == "analysis" else 1) * alg["z_pass"]
            roof["kernels"][kname] = {"bound": "hbm" if kname == "analysis" else "tensor", "avg_launch_ms": ms,
                                      "tflops": alg["conv_flops"] / (ms * 1e-3) / 1e12, "frac_tensor": alg["conv_flops"] / (ms * 1e-3) / 1e12 / tf32_peak,
                                      "gbs": zb / (ms * 1e-3) / 1e9, "frac_hbm": zb / (ms * 1e-3) / 1e9 / peaks.get("hbm_gbs", 6650.0),
                                      "algorithmic_bytes_per_launch": zb,
                                      "floor_ms": max(zb / (peaks.get("hbm_gbs", 6650.0) * 1e9), alg["conv_flops"] / (tf32_peak * 1e12)) * 1e3}
        # north_star's "fraction of the per-iteration roofline": each kernel at the higher of its HBM and tensor floors
        it_floor = sum(v["floor_ms"] for v in roof["kernels"].values())
        it_ms = sum(v["avg_launch_ms"] for v in roof["kernels"].values())
        roof["per_iteration"] = {"floor_ms": it_floor, "achieved_ms": it_ms, "frac": it_floor / it_ms}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rcpu = cpu_reference(1, 0)
        cpu = {"value": rcpu["value"], "unit": "Mvoxels/s", "cores": rcpu["cores"], "kind": rcpu["kind"], "sample": rcpu["sample"]}

    if rank == 0:
        line = {"metric": "denoised Mvoxels/s (K-iter CDLNet-3D fwd)", "value": value, "unit": "Mvoxels/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "tf32" if plan.precision == "tf32" else "f32", "data": "synthetic",
                "config": {"workload": workload, "precision": plan.precision, "parallelism": f"clips sharded over {world} GPU(s), no collective",
                           "l2_policy": f"inputs larger than L2: z is {alg['z_pass'] / 1e6:.0f} MB per pass, updated in place",
                           "z_nonzero_frac": nnz, "psnr_in_db": psnr_in, "psnr_out_db": psnr_out},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "Mvoxels/s", "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": int(y_host.numel() * 4 + c_host.numel() * 4), "d2h_bytes_per_step": int(x_host.numel() * 4),
                        "max_abs_diff_vs_device_path": e2e_diff},
                "gpu_launches": int(launches),
                "algorithmic": {"flops_per_step": alg["flops_fwd"], "bytes_per_step": alg["bytes_fwd"],
                                "tflops": alg["flops_fwd"] / (ms_step * 1e-3) / 1e12, "gbs": alg["bytes_fwd"] / (ms_step * 1e-3) / 1e9}}
        if roof:
            line["roofline"] = roof
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
