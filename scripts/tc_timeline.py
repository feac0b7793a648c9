#!/usr/bin/env python
"""Development aid: per-warp-role cycle accounting of the tcgen05 video kernels (uses the undocumented
cdl__debug_set_buffer hook; needs a library built with -DCDL_TC_PROFILE, e.g.
  CDL_TC_PROFILE=1 CDL_LIB_PATH=$PWD/cdlnet-video_b200/libcdl_b200_prof.so python -c "import cdlnet_video_b200 as c; c.build(force=True)"
  CDL_LIB_PATH=$PWD/cdlnet-video_b200/libcdl_b200_prof.so python scripts/tc_timeline.py [clips])"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import cdlnet_video_b200 as cb

clips = int(sys.argv[1]) if len(sys.argv) > 1 else 4
d = torch.device("cuda", 0)
K, M = bench.CFG["K"], bench.CFG["M"]
plan = cb.Plan(3, clips, 1, M, K, bench.CLIP, (7, 7, 7), 2, precision="tf32")
A, B, u = bench.synthetic_weights(torch, d)
clean, y = bench.synthetic_clip(torch, clips, 0, d)
plan.set_weights(A, B, bench.calibrate_thresholds(torch, A, B, u, y[:1], d))
c = torch.full((clips,), 25 / 255.0, device=d)
yp, _, mean = plan.preprocess(y)
z = plan.new_code()
r = torch.empty_like(yp)
plan.set_rearm(True)
plan.analysis_step(0, yp, z, c, first=True)
for k in range(1, 4):
    plan.synthesis_step(k, z, r, yp, None, residual=True)
    plan.analysis_step(k, r, z, c)
torch.cuda.synchronize()
lib = cb.load_library()
dbg = torch.zeros(148 * 24 * 8, dtype=torch.int64, device=d)
lib.cdl__debug_set_buffer.argtypes = [ctypes.c_void_p]
lib.cdl__debug_set_buffer(ctypes.c_void_p(dbg.data_ptr()))


def show(name, roles):
    torch.cuda.synchronize()
    t = dbg.view(148, 24, 8).cpu().double()
    print(f"== {name}: cycles (mean over CTAs; rank0 = even blocks)")
    for rname, warps, labels in roles:
        for rk in (0, 1):
            sel = t[rk::2][:, warps, :].mean(dim=(0, 1))
            tot = sel[0].item()
            print(f"  {rname:9s} rank{rk}: total {tot:9.0f} | " + " | ".join(f"{lab} {sel[i + 1].item():9.0f} ({100 * sel[i + 1].item() / max(tot, 1):4.1f}%)" for i, lab in enumerate(labels)))
    dbg.zero_()


e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); plan.synthesis_step(4, z, r, yp, None, residual=True); e1.record()
show("synthesis", [("col2im", list(range(0, 16)), ["wait dfull", "tmem->ring", "barriers+flush"]),
                   ("mma", [16], ["wait dempty", "wait A chunk", "wait weights"]),
                   ("tma", [17], ["wait aempty", "-", "-"])])
print("   launch ms", e0.elapsed_time(e1), " tiles per CTA", clips * 8 * 128 / 148.0)
e0.record(); plan.analysis_step(4, r, z, c); e1.record()
show("analysis", [("epilogue", list(range(0, 16)), ["wait dfull", "-", "-"]),
                  ("mma", [16], ["wait dempty", "wait r tiles", "wait weights"]),
                  ("loader", [17], ["wait rempty", "wait rfull", "-"])])
print("   launch ms", e0.elapsed_time(e1))
