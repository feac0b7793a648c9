#!/usr/bin/env python
"""bench.py — denoised Mvoxels/s of the K-iteration CDLNet-3D forward pass (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (N > 1: under torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N ...            # the reference's CPU path (oracle port) on the host cores
  python bench.py --workload cfg2 ...                      # secondary line: BASELINE config 2 (clips sharded, weak scaling)

Default workload = BASELINE config 5, the configuration north_star's multi-GPU design is about: CDLNetVideo(args3d.json:
K=30, M=169, P=7 -> 7x7x7, s=2, C=1, adaptive) denoising ONE synthetic 240-frame 1080p clip at sigma=25.  One step =
one forward pass (mean/pad preprocess + K ISTA iterations + D z + crop) over the whole clip.  N GPUs split the clip
TEMPORALLY (cdlnet_video_b200.sharded.ShardedVideoDenoiser -> cdl_forward_sharded): every rank keeps its slab of the
42 GB sparse code resident and exchanges the Pt - s = 5 seam frames of its partial synthesis with its ring neighbours
once per iteration over NCCL P2P.  STRONG scaling: the clip is fixed, N = 1 runs the same clip, same entry point, on
one GPU.  `value` = D*H*W / step time (max over ranks, CUDA events, clip resident in HBM); `e2e` = the same through
ShardedVideoDenoiser.denoise_host: pinned host slab -> H2D -> forward -> D2H of the owned frames of xhat.

The reference arm (`--impl reference`, and the in-line `cpu_baseline`) times the oracle port of the reference's
torch-CPU arithmetic on a bounded sample OF THE SAME clip with the SAME weights and thresholds: frames 0..15, rows and
columns 0..255 (the 16-frame window the reference's own driver would cut, analyze3d.py:62,105-106, cropped to 256^2).
"""
import argparse
import json
import math
import os
import random
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(K=30, M=169, P=7, s=2, C=1)          # /root/reference args3d.json:3-12 (scalar P = cubic)
CLIP = (16, 256, 256)                           # BASELINE config 2 clip (also the CPU arm's sample window)
CLIP5 = (240, 1080, 1920)                       # BASELINE config 5 clip
SIGMA = 25.0
L_SPECTRAL = 1.375e4                            # reference test.ipynb:171, power-method constant of this filter family
METRIC = "denoised Mvoxels/s (K-iter CDLNet-3D fwd)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg5", choices=["cfg5", "cfg2"])
    ap.add_argument("--frames", type=int, default=CLIP5[0])
    ap.add_argument("--height", type=int, default=CLIP5[1])
    ap.add_argument("--width", type=int, default=CLIP5[2])
    ap.add_argument("--clips", type=int, default=4, help="cfg2: clips per GPU per step")
    ap.add_argument("--precision", default=os.environ.get("CDL_PRECISION", "auto"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-breakdown", action="store_true")
    ap.add_argument("--no-check", action="store_true", help="cfg5, N > 1: skip the sharded-vs-unsharded parity check")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# synthetic inputs and weights (no datasets / checkpoints exist: SURVEY F6)
# ------------------------------------------------------------------------------------------------
def _waves(seed):
    """The random plane-wave terms of one synthetic clip (syn_data/gen.py:12-24): [(sign, fn name, cx, cy, cz)]."""
    rng = random.Random(seed)
    terms = []
    for i in range(rng.randint(2, 10)):
        cx, cy, cz = rng.uniform(-5, 5), rng.uniform(-5, 5), rng.uniform(-5, 5)
        fn = rng.choice(["sin", "cos"])
        sign = 1.0 if (i == 0 or rng.choice([True, False])) else -1.0
        terms.append((sign, fn, cx, cy, cz))
    return terms


def clip_window(torch, shape, seed, device, frames=None, rows=None, cols=None):
    """Frames [f0,f1) x rows [r0,r1) x cols [c0,c1) of the synthetic clip `seed` of extents `shape`: a random signed sum of
    2-10 sin/cos plane waves on [-pi,pi]^3 (syn_data/gen.py:12-31) normalised to [0,1] with the min/max over the clip's
    every-4th-voxel lattice (the reference min-max normalises per clip, syn_data/gen_data_draft.py:33-38; the lattice makes
    the same clip computable window by window on any rank and on the CPU arm), plus AWGN sigma/255 seeded per frame
    (utils.py:44-55).  Returns (clean, noisy), each (1,1,f,r,c) fp32."""
    D, H, W = shape
    f0, f1 = frames or (0, D)
    r0, r1 = rows or (0, H)
    c0, c1 = cols or (0, W)
    terms = _waves(seed)

    def field(gz, gx, gy):
        acc = None
        for sign, fn, cx, cy, cz in terms:
            t = getattr(torch, fn)(cx * gx + cy * gy + cz * gz)
            acc = sign * t if acc is None else acc + sign * t
        return acc
    lin = lambda n: torch.linspace(-math.pi, math.pi, n, device=device, dtype=torch.float32)
    gz, gx, gy = lin(D), lin(H), lin(W)
    lat = field(gz[::4].view(-1, 1, 1), gx[::4].view(1, -1, 1), gy[::4].view(1, 1, -1))
    lo, hi = float(lat.min()), float(lat.max())
    del lat
    clean = field(gz[f0:f1].view(-1, 1, 1), gx[r0:r1].view(1, -1, 1), gy[c0:c1].view(1, 1, -1))
    clean = ((clean - lo) / (hi - lo)).clamp_(0.0, 1.0)
    noisy = torch.empty_like(clean)
    for i, f in enumerate(range(f0, f1)):                     # per-frame generator: every rank / arm draws the same noise
        g = torch.Generator(device="cpu").manual_seed(1000003 * (seed + 1) + f)
        n = torch.randn(H, W, generator=g)[r0:r1, c0:c1]
        noisy[i] = clean[i] + n.to(device) * (SIGMA / 255.0)
    return clean[None, None], noisy[None, None]


def synthetic_clip(torch, n, seed, device, shape=CLIP):
    """n independent clips of extents `shape` (config 2 and the tests): (clean, noisy), each (n,1,D,H,W)."""
    outs = [clip_window(torch, shape, seed * 131 + i, device) for i in range(n)]
    return torch.cat([o[0] for o in outs]), torch.cat([o[1] for o in outs])


def synthetic_weights(torch, device, seed=1):
    """SURVEY 8(d) protocol: one randn bank / sqrt(L), 3 % per-layer perturbation; thresholds are set afterwards from the
    85th percentile of |A_0 yp| (calibrate_thresholds)."""
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


def calibrate_thresholds(torch, A, B, u, y_window, device):
    """t (K,2,M) from the 85th percentile q_m of |A_0 (w - mean(w))| over the window w.  Plain torch conv3d in
    fp32 (TF32 off) on whatever device the window lives on: the GPU arm and the CPU arm get the same thresholds to ~1e-7.
    Set-up code, not the product path."""
    import torch.nn.functional as F
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        w = y_window.to(device)
        a0 = F.conv3d(w - w.mean(), A[0].to(device), stride=CFG["s"], padding=CFG["P"] // 2).abs()
    finally:
        torch.backends.cudnn.allow_tf32 = old
    q = torch.quantile(a0[0].reshape(CFG["M"], -1)[:, ::8].float(), 0.85, dim=1)
    return thresholds_from_quantile(torch, q, u.to(device))


def sample_window(torch, shape, device):
    """The CPU arm's bounded sample of the cfg-5 clip: frames 0..15, rows / cols 0..255."""
    D, H, W = shape
    return clip_window(torch, shape, 0, device, frames=(0, min(16, D)), rows=(0, min(256, H)), cols=(0, min(256, W)))


# ------------------------------------------------------------------------------------------------
# algorithmic work
# ------------------------------------------------------------------------------------------------
def algorithmic(n_clips, clip=CLIP):
    K, M, C, P, s = CFG["K"], CFG["M"], CFG["C"], CFG["P"], CFG["s"]
    V = clip[0] * clip[1] * clip[2]
    Q = V // s ** 3
    T = P ** 3
    conv_flops = 2.0 * n_clips * Q * M * C * T                 # one analysis or one synthesis
    flops_fwd = 2 * K * conv_flops
    bytes_fwd = 8.0 * K * n_clips * Q * M + 4.0 * n_clips * V * C * (K + 2)
    z_pass = 4.0 * n_clips * Q * M
    return dict(V=V, Q=Q, conv_flops=conv_flops, flops_fwd=flops_fwd, bytes_fwd=bytes_fwd, z_pass=z_pass)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms during the timed region."""
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


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


def measure_tf32_peak(torch, dev, seconds=1.5):
    """cuBLAS TF32 GEMM 8192^3 measured the way MEASURED_PEAKS.json measures bf16: best of 10 (burst) and a back-to-back
    loop under the power cap (sustained).  A library GEMM as a yardstick, not part of the product path."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn(n, n, device=dev)
        b = torch.randn(n, n, device=dev)
        for _ in range(3):
            a @ b
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        reps = max(10, int(seconds * 1e3 / best))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            a @ b
        e1.record(); torch.cuda.synchronize()
        fl = 2.0 * n ** 3
        return {"tf32_tflops_burst": fl / (best * 1e-3) / 1e12, "tf32_tflops_sustained": fl * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12,
                "how": f"torch.matmul fp32 with allow_tf32 (cuBLAS) {n}^3: best of 10 and {reps} back to back"}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


# ------------------------------------------------------------------------------------------------
# CPU reference arm: the oracle port of the reference's torch-CPU path, on the bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_reference(steps, warmup, shape, budget_s=150.0):
    """oracle.forward_t (the reference's arithmetic: torch CPU conv3d / conv_transpose3d, fp32, all host cores) on the
    sample window of the cfg-5 clip, same weights / thresholds / noise as the GPU arm."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cdl_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    K, M, s = CFG["K"], CFG["M"], CFG["s"]
    cpu = torch.device("cpu")
    A, B, u = synthetic_weights(torch, cpu)
    clean, y = sample_window(torch, shape, cpu)
    t = calibrate_thresholds(torch, A, B, u, y, cpu).reshape(K, 2, M, 1, 1, 1)

    def run():
        t0 = time.perf_counter()
        with torch.no_grad():
            O.forward_t(y, A, B, t, s, SIGMA, True, 1)
        return time.perf_counter() - t0
    times, spent = [], 0.0
    for i in range(max(1, warmup) + max(1, steps)):
        dt = run()
        spent += dt
        if i >= max(1, warmup):
            times.append(dt)
        if spent + dt > budget_s and times:                   # keep the whole run within a few minutes: fewer timed passes,
            break                                             # never a smaller sample
    vox = y.numel()
    dt = sum(times) / len(times)
    dims = "x".join(str(v) for v in y.shape[2:])
    return dict(value=vox / dt / 1e6, ms=dt * 1e3, cores=cores, kind="port", passes=len(times),
                sample=f"window {dims} (frames 0-15, rows/cols 0-255) of the same clip, same weights and thresholds, full K={K}; "
                       f"fp32 torch-CPU oracle (oracle/cdl_oracle.py forward_t), {max(1, warmup)} warm-up + {len(times)} timed pass(es)")


def workload_name(args, world):
    if args.workload == "cfg2":
        return (f"cfg2: CDLNetVideo(args3d.json) K={CFG['K']} M={CFG['M']} P=7x7x7 s={CFG['s']} C=1 adaptive, "
                f"{args.clips} clip(s) x {CLIP[0]}x{CLIP[1]}x{CLIP[2]} per GPU per step, sigma={SIGMA:g}")
    return (f"cfg5: CDLNetVideo(args3d.json) K={CFG['K']} M={CFG['M']} P=7x7x7 s={CFG['s']} C=1 adaptive, ONE "
            f"{args.frames}x{args.height}x{args.width} clip per step, sigma={SIGMA:g}, split temporally over the GPUs")


def reference_arm(args):
    shape = (args.frames, args.height, args.width) if args.workload == "cfg5" else CLIP
    r = cpu_reference(args.steps, args.warmup, shape)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "Mvoxels/s",
            "n_gpus": args.gpus, "steps": r["passes"], "warmup": max(1, args.warmup), "ms_per_step": r["ms"],
            "higher_is_better": True, "scaling": "strong" if args.workload == "cfg5" else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args, args.gpus), "reference_sample": r["sample"]},
            "cpu_baseline": {"value": r["value"], "unit": "Mvoxels/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": "Mvoxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# per-kernel breakdown -> roofline
# ------------------------------------------------------------------------------------------------
def roofline_from_times(tk, alg, peaks, peak_src, eff, n_units, tf32_meas):
    """tk: {"analysis": [ms...], "synthesis": [ms...]} CUDA-event times of every launch of one forward on this rank;
    alg: algorithmic() of THIS RANK's share."""
    tot = {k: sum(v) for k, v in tk.items()}
    dom = max(tot, key=tot.get)
    avg_ms = tot[dom] / len(tk[dom])
    tf32_peak = peaks.get("bf16_tflops_sustained", 1400.0) / 2.0
    hbm = peaks.get("hbm_gbs", 6650.0)
    tflops = alg["conv_flops"] / (avg_ms * 1e-3) / 1e12
    zb_dom = (2 if dom == "analysis" else 1) * alg["z_pass"]
    gbs = zb_dom / (avg_ms * 1e-3) / 1e9
    traffic = None
    try:                                     # DRAM bytes per launch of this kernel from the committed ncu --set full capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        for cap in tj.get("captures", []):     # the capture of THIS workload size (per-launch bytes do not scale otherwise)
            if eff == "tf32" and cap.get("units") == n_units and dom in cap:
                traffic = cap[dom]["dram_bytes_per_launch"]
                break
    except Exception:
        pass
    bound = "tensor" if dom == "synthesis" else "hbm"
    roof = {"kernel": f"{dom} ({'tcgen05 tf32' if eff == 'tf32' else 'CUDA-core fp32'})", "bound": bound,
            "achieved": tflops if bound == "tensor" else gbs, "peak": tf32_peak if bound == "tensor" else hbm,
            "unit": "TFLOP/s" if bound == "tensor" else "GB/s",
            "frac": (tflops / tf32_peak) if bound == "tensor" else gbs / hbm, "traffic": traffic,
            "peak_source": f"{peak_src}: tensor = bf16 sustained / 2 (the driver file has no tf32 entry), hbm = copy bandwidth",
            "avg_launch_ms": avg_ms, "launches_per_step": len(tk[dom]), "share_of_step": tot[dom] / sum(tot.values()),
            "per_kernel_ms_per_step": tot, "algorithmic_flops_per_launch": alg["conv_flops"],
            "algorithmic_bytes_per_launch": zb_dom}
    if tf32_meas:
        roof["tf32_peak_measured"] = tf32_meas
        roof["frac_of_measured_tf32_sustained"] = tflops / tf32_meas["tf32_tflops_sustained"]
        roof["frac_of_measured_tf32_burst"] = tflops / tf32_meas["tf32_tflops_burst"]
    roof["kernels"] = {}
    for kname in tot:
        ms = tot[kname] / len(tk[kname])
        zb = (2 if kname == "analysis" else 1) * alg["z_pass"]
        roof["kernels"][kname] = {"bound": "hbm" if kname == "analysis" else "tensor", "avg_launch_ms": ms,
                                  "tflops": alg["conv_flops"] / (ms * 1e-3) / 1e12, "frac_tensor": alg["conv_flops"] / (ms * 1e-3) / 1e12 / tf32_peak,
                                  "gbs": zb / (ms * 1e-3) / 1e9, "frac_hbm": zb / (ms * 1e-3) / 1e9 / hbm,
                                  "algorithmic_bytes_per_launch": zb,
                                  "floor_ms": max(zb / (hbm * 1e9), alg["conv_flops"] / (tf32_peak * 1e12)) * 1e3}
    # north_star's per-iteration roofline, SURVEY 8(d): max(bytes / BW_HBM, flops / peak) of the WHOLE iteration
    # (two z passes + the image passes; two convolutions)
    it_bytes = 2 * alg["z_pass"] + 4.0 * alg["V_rank"] * 1
    it_flops = 2 * alg["conv_flops"]
    it_floor = max(it_bytes / (hbm * 1e9), it_flops / (tf32_peak * 1e12)) * 1e3
    it_ms = sum(tot.values()) / max(len(tk["analysis"]), 1)
    roof["per_iteration"] = {"floor_ms": it_floor, "achieved_ms": it_ms, "frac": it_floor / it_ms,
                             "definition": "max(bytes/BW_HBM, flops/tf32 peak) of one iteration (SURVEY 8d)"}
    return roof


# ------------------------------------------------------------------------------------------------
# config 5: one long clip, temporal slabs (strong scaling)
# ------------------------------------------------------------------------------------------------
def run_cfg5(args, rank, world, local):
    import torch
    import torch.distributed as dist
    import cdlnet_video_b200 as cb
    from cdlnet_video_b200 import sharded
    dev = torch.device("cuda", local)
    shape = (args.frames, args.height, args.width)
    D, H, W = shape
    K, M, s = CFG["K"], CFG["M"], CFG["s"]
    prec = "fp32" if args.precision == "fp32" else "tf32"

    A, B, u = synthetic_weights(torch, dev)
    _, ywin = sample_window(torch, shape, dev)
    t = calibrate_thresholds(torch, A, B, u, ywin, dev)
    net = cb.CDLNetVideo(K=K, M=M, P=CFG["P"], s=s, C=1, adaptive=True, init=False)
    with torch.no_grad():
        for k in range(K):
            net.A[k].weight.copy_(A[k]); net.B[k].weight.copy_(B[k])
        net.t.copy_(t.reshape(net.t.shape))
    net = net.to(dev).eval()
    den = sharded.ShardedVideoDenoiser(net, (1, 1, D, H, W), rank, world, dev, precision=prec)
    g = den.geo
    clean, y = clip_window(torch, shape, 0, dev, frames=(g["f0"], g["f1"]))
    own = slice(g["hf"], y.shape[2] - g["hb"])
    clean_owned = clean[:, :, own].clone()
    del clean
    stream = torch.cuda.current_stream()

    def step():
        return den.forward_resident(y, SIGMA)[0]

    for _ in range(max(args.warmup, 3)):
        xhat = step()
    torch.cuda.synchronize()
    mse = torch.stack([((xhat - clean_owned) ** 2).sum(), ((y[:, :, own] - clean_owned) ** 2).sum()]).double()
    if world > 1:
        dist.all_reduce(mse)
    psnr_out, psnr_in = [float(-10 * torch.log10(v / (D * H * W))) for v in mse]

    # ---- sharded == unsharded?  rank 0 runs the whole clip on its own GPU once, every rank compares its owned frames
    check = None
    if world > 1 and not args.no_check:
        full = torch.empty(1, 1, D, H, W, device=dev)
        if rank == 0:
            _, yfull = clip_window(torch, shape, 0, dev)
            one = sharded.ShardedVideoDenoiser(net, (1, 1, D, H, W), 0, 1, dev, precision=prec)
            full.copy_(one.forward_resident(yfull, SIGMA)[0])
            del one, yfull
            torch.cuda.empty_cache()
        dist.broadcast(full, src=0)
        diff = (full[:, :, s * g["q0"]:s * g["q1"]] - xhat).abs().max().reshape(1)
        dist.all_reduce(diff, op=dist.ReduceOp.MAX)
        check = float(diff)
        del full
        torch.cuda.empty_cache()

    # ---- timed region: device-resident -----------------------------------------------------------
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = den.plan.launch_count() + den.sum_plan.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = den.plan.launch_count() + den.sum_plan.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    tms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_step = tms.item() / args.steps
    V = D * H * W
    value = V / (ms_step * 1e-3) / 1e6

    # ---- e2e: pinned host slab -> H2D -> forward -> D2H of the owned frames ----------------------------
    e2e = None
    if not args.no_e2e:
        y_host = y.cpu().pin_memory()
        x_host = torch.empty_like(xhat, device="cpu").pin_memory()
        n_e2e = max(2, min(args.steps, 5))
        den.denoise_host(y_host, x_host, SIGMA)
        den.wait()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0.record(stream)
        for _ in range(n_e2e):
            den.denoise_host(y_host, x_host, SIGMA)      # every step: its own H2D, forward, D2H; the copies of neighbouring steps overlap
        den.join()                                        # the timed region ends when the LAST download has landed
        e1.record(stream)
        torch.cuda.synchronize()
        te = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        nb = torch.tensor([y_host.numel() * 4.0, x_host.numel() * 4.0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
            dist.all_reduce(nb)
        e2e_ms = te.item() / n_e2e
        e2e = {"value": V / (e2e_ms * 1e-3) / 1e6, "unit": "Mvoxels/s", "ms_per_step": e2e_ms, "steps": n_e2e,
               "h2d_bytes_per_step": int(nb[0].item()), "d2h_bytes_per_step": int(nb[1].item()),
               "max_abs_diff_vs_device_path": float((x_host.to(dev) - xhat).abs().max()),
               "api": "ShardedVideoDenoiser.denoise_host (pinned host slab in, owned xhat frames out; uploads / downloads on copy streams, double-buffered)"}
        del y_host, x_host

    # ---- per-kernel breakdown on this rank's slab (all ranks run it: the exchange is collective) ------
    roof = None
    if not args.no_breakdown:
        plan = den.plan
        den._buffers()
        code, r, halo = den.code, den.r, den.halo
        hb = plan.halo_bytes // 4
        rp, rn = halo[:hb], halo[hb:]
        c = net._c_vector(SIGMA, 1, dev)
        mean = plan.mean_from_sums(den.sum_plan.reduce_sums(y[:, :, own].contiguous()))     # (local mean: timing only)
        yp = plan.center_pad(y, mean)[0]
        acc = {"analysis": [], "synthesis": [], "exchange": []}

        def timed(kind, fn):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); fn(); b.record(stream)
            acc[kind].append((a, b))
        lib, comm = plan.lib, den.comm
        from cdlnet_video_b200.plan import _ptr, _stream

        def exchange():
            if world > 1:
                cb._lib.check(lib.cdl_halo_exchange(plan.handle, comm.handle, _ptr(r), _ptr(rp), _ptr(rn), _stream()), "cdl_halo_exchange")
        plan.set_rearm(True)
        timed("analysis", lambda: plan.analysis_step(0, yp, code, c, first=True))
        for k in range(1, K):
            timed("synthesis", lambda: plan.synthesis_step(k, code, r, yp, None, residual=True))
            timed("exchange", exchange)
            timed("analysis", lambda: plan.analysis_step_halo(k, r, code, c, rp, rn, yp))
        plan.set_rearm(False)
        timed("synthesis", lambda: plan.synthesis_step(0, code, r, residual=False))
        torch.cuda.synchronize()
        if rank == 0:
            peaks, peak_src = load_peaks()
            tf32_meas = measure_tf32_peak(torch, dev) if world == 1 else None
            tk = {k: [a.elapsed_time(b) for a, b in v] for k, v in acc.items()}
            xl = sorted(tk.pop("exchange"))
            xch_ms = sum(xl)
            Vr = (g["q1"] - g["q0"]) * s * H * W                 # this rank's owned voxels
            alg = algorithmic(1, ((g["q1"] - g["q0"]) * s, H, W))
            alg["V_rank"] = Vr
            roof = roofline_from_times(tk, alg, peaks, peak_src, plan.precision, f"{(g['q1'] - g['q0']) * s}x{H}x{W}", tf32_meas)
            roof["exchange_ms_per_step"] = xch_ms
            roof["exchange_ms"] = {"min": xl[0], "median": xl[len(xl) // 2], "max": xl[-1], "n": len(xl),
                                   "bytes_per_direction_per_seam": den.plan.halo_bytes}
            roof["rank"] = 0
        del yp

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rcpu = cpu_reference(2, 1, shape, budget_s=40.0)
        cpu = {"value": rcpu["value"], "unit": "Mvoxels/s", "cores": rcpu["cores"], "kind": rcpu["kind"], "sample": rcpu["sample"]}

    if rank == 0:
        alg = algorithmic(1, shape)
        line = {"metric": METRIC, "value": value, "unit": "Mvoxels/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "tf32" if den.plan.precision == "tf32" else "f32", "data": "synthetic",
                "config": {"workload": workload_name(args, world), "precision": den.plan.precision,
                           "parallelism": (f"temporal slabs over {world} GPU(s): {g['q1'] - g['q0']} coarse frames per rank, "
                                           f"{g['overlap']}-frame image-domain halo exchange per seam per iteration (NCCL P2P, cdl_forward_sharded)"
                                           if world > 1 else "1 GPU: whole clip through cdl_forward_sharded without a communicator"),
                           "l2_policy": f"inputs larger than L2: the code is {alg['z_pass'] / 1e9 / world:.1f} GB per rank, updated in place",
                           "psnr_in_db": psnr_in, "psnr_out_db": psnr_out,
                           "sharded_vs_unsharded_max_abs": check},
                "clocks": clocks, "gpu_launches": int(launches),
                "algorithmic": {"flops_per_step": alg["flops_fwd"], "bytes_per_step": alg["bytes_fwd"],
                                "tflops": alg["flops_fwd"] / (ms_step * 1e-3) / 1e12, "gbs": alg["bytes_fwd"] / (ms_step * 1e-3) / 1e9}}
        if e2e:
            line["e2e"] = e2e
        if roof:
            line["roofline"] = roof
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# config 2: independent 16x256x256 clips sharded over the GPUs (weak scaling; secondary line)
# ------------------------------------------------------------------------------------------------
def run_cfg2(args, rank, world, local):
    import torch
    import torch.distributed as dist
    import cdlnet_video_b200 as cb
    dev = torch.device("cuda", local)
    n_clips = args.clips
    alg = algorithmic(n_clips)
    alg["V_rank"] = n_clips * alg["V"]
    prec = "fp32" if args.precision == "fp32" else "tf32"
    plan = cb.Plan(3, n_clips, CFG["C"], CFG["M"], CFG["K"], CLIP, (CFG["P"],) * 3, CFG["s"], has_mask=False,
                   precision=prec, device=local)
    A, B, u = synthetic_weights(torch, dev)
    clean, y = synthetic_clip(torch, n_clips, seed=rank, device=dev)
    c = torch.full((n_clips,), SIGMA / 255.0, dtype=torch.float32, device=dev)
    t = calibrate_thresholds(torch, A, B, u, y[:1], dev)
    plan.set_weights(A, B, t)
    torch.cuda.synchronize()
    z = torch.empty(plan.z_shape, device=dev)
    stream = torch.cuda.current_stream()

    def step():
        return plan.denoise(y, None, c, z_out=z)

    for _ in range(max(args.warmup, 3)):
        xhat, _ = step()
    torch.cuda.synchronize()
    nnz = float((z != 0).float().mean())
    psnr_in = float(-10 * torch.log10(((y - clean) ** 2).mean()))
    psnr_out = float(-10 * torch.log10(((xhat - clean) ** 2).mean()))
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
    clocks = sampler.stop() if sampler else None
    tms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_step = tms.item() / args.steps
    value = world * n_clips * alg["V"] / (ms_step * 1e-3) / 1e6

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

    roof = None
    if rank == 0 and not args.no_breakdown:
        peaks, peak_src = load_peaks()
        ypb, _, meanb = plan.preprocess(y)
        r = torch.empty_like(ypb)
        code = plan.new_code()
        acc = {"analysis": [], "synthesis": []}

        def timed(kind, fn):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); fn(); b.record(stream)
            acc[kind].append((a, b))
        plan.set_rearm(True)
        timed("analysis", lambda: plan.analysis_step(0, ypb, code, c, first=True))
        for k in range(1, CFG["K"]):
            timed("synthesis", lambda: plan.synthesis_step(k, code, r, ypb, None, residual=True))
            timed("analysis", lambda: plan.analysis_step(k, r, code, c))
        plan.set_rearm(False)
        timed("synthesis", lambda: plan.synthesis_step(0, code, r, residual=False))
        torch.cuda.synchronize()
        tk = {k: [a.elapsed_time(b) for a, b in v] for k, v in acc.items()}
        tf32_meas = measure_tf32_peak(torch, dev) if world == 1 else None
        roof = roofline_from_times(tk, alg, peaks, peak_src, plan.precision, f"{n_clips} clips", tf32_meas)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rcpu = cpu_reference(2, 1, CLIP, budget_s=40.0)
        cpu = {"value": rcpu["value"], "unit": "Mvoxels/s", "cores": rcpu["cores"], "kind": rcpu["kind"], "sample": rcpu["sample"]}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "Mvoxels/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "tf32" if plan.precision == "tf32" else "f32", "data": "synthetic",
                "config": {"workload": workload_name(args, world), "precision": plan.precision,
                           "parallelism": f"clips sharded over {world} GPU(s), no collective",
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


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        if rank == 0:
            reference_arm(args)
        return
    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    try:
        (run_cfg5 if args.workload == "cfg5" else run_cfg2)(args, rank, world, local)
    finally:
        if world > 1:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
