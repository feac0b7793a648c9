#!/usr/bin/env python
"""Development aid: raw event trace of one analysis tile (block 0, 4th tile) - needs a CDL_TC_PROFILE build and
CDL_TC_DBG_MODE=8 (or 9/10/11 for the skeleton variants).  python scripts/tc_trace.py [clips]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CDL_TC_DBG_MODE", "8")
import torch
import bench
import cdlnet_video_b200 as cb

clips = int(sys.argv[1]) if len(sys.argv) > 1 else 4
d = torch.device("cuda", 0)
K, M = bench.CFG["K"], bench.CFG["M"]
plan = cb.Plan(3, clips, 1, M, K, bench.CLIP, (7, 7, 7), 2, precision="tf32")
A, B, u = bench.synthetic_weights(torch, d)
plan.set_weights(A, B, torch.rand(K, 2, M, device=d) * 0.01)
clean, y = bench.synthetic_clip(torch, clips, 0, d)
c = torch.full((clips,), 25 / 255.0, device=d)
yp, _, mean = plan.preprocess(y)
z = plan.new_code()
r = torch.empty_like(yp)
plan.analysis_step(0, yp, z, c, first=True)
plan.synthesis_step(1, z, r, yp, None, residual=True)
lib = cb.load_library()
dbg = torch.zeros(148 * 24 * 8 + 512, dtype=torch.int64, device=d)
lib.cdl__debug_set_buffer.argtypes = [ctypes.c_void_p]
lib.cdl__debug_set_buffer(ctypes.c_void_p(dbg.data_ptr()))
plan.analysis_step(1, r, z, c)
torch.cuda.synchronize()
ev = dbg[148 * 24 * 8:].cpu().tolist()
t0 = min(v for v in ev if v > 0)
rows = []
names = {30: "mma tile start (dempty ok)", 31: "mma dfull commit issued", 63: "prod tile start", 120: "epi z loads issued", 121: "epi dfull ok", 122: "epi done"}
for i, v in enumerate(ev):
    if v <= 0:
        continue
    if i < 30:
        nm = f"mma ch{i // 3} " + ["afull ok", "mmas issued", "commit issued"][i % 3]
    elif 64 <= i < 120:
        j = i - 64
        nm = f"prod ch{j // 4} " + ["lds issued", "aempty ok", "st done", "arrived"][j % 4]
    else:
        nm = names.get(i, str(i))
    rows.append((v - t0, nm))
for t, nm in sorted(rows):
    print(f"{t:8d}  {nm}")
