#!/usr/bin/env python
"""Phase ceilings of the video synthesis kernel (results invalid in the debug modes): time one residual synthesis launch
on the bench's config-2 workload with parts of the kernel compiled out by CDL_TC_DBG_MODE (read at plan creation):
  0 = full kernel, 64 = no col2im, 128 = no flush, 192 = neither, 256 = TMA signals without moving data,
  448 = the MMA / barrier skeleton alone.   python scripts/syn_phase.py [clips] [modes...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
import cdlnet_video_b200 as cb


def main():
    clips = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    modes = [int(m) for m in sys.argv[2:]] or [0, 64, 128, 192, 256, 448]
    dev = torch.device("cuda", 0)
    A, B, u = bench.synthetic_weights(torch, dev)
    clean, y = bench.synthetic_clip(torch, clips, seed=0, device=dev)
    t = bench.calibrate_thresholds(torch, A, B, u, y[:1], dev)
    c = torch.full((clips,), bench.SIGMA / 255.0, device=dev)
    out = {}
    for mode in modes:
        os.environ["CDL_TC_DBG_MODE"] = str(mode)
        plan = cb.Plan(3, clips, 1, bench.CFG["M"], bench.CFG["K"], bench.CLIP, (7, 7, 7), 2, precision="tf32")
        os.environ.pop("CDL_TC_DBG_MODE")
        plan.set_weights(A, B, t)
        yp, _, mean = plan.preprocess(y)
        code, r = plan.new_code(), torch.empty_like(yp)
        plan.analysis_step(0, yp, code, c, first=True)
        plan.set_rearm(True)
        for k in range(1, 4):                      # a realistic code (a few iterations), the buffer armed with -yp
            plan.synthesis_step(k, code, r, yp, None, residual=True)
            plan.analysis_step(k, r, code, c)
        torch.cuda.synchronize()
        ts = {"synthesis": [], "analysis": []}
        for k in range(4, 14):
            for kind, fn in (("synthesis", lambda: plan.synthesis_step(k, code, r, yp, None, residual=True)),
                             ("analysis", lambda: plan.analysis_step(k, r, code, c))):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record()
                ts[kind].append((e0, e1))
        torch.cuda.synchronize()
        out[mode] = {k: round(sum(a.elapsed_time(b) for a, b in v) / len(v) * 1e3, 1) for k, v in ts.items()}
        plan.close()
    tiles = clips * 8 * 128 / 148.0
    for mode, v in out.items():
        v["cycles_per_tile_at_1.9GHz"] = round(v["synthesis"] * 1e-6 * 1.9e9 / tiles)
    print(json.dumps({"clips": clips, "us_per_launch": out}))


if __name__ == "__main__":
    main()
