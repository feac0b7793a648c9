#!/usr/bin/env python
"""Top stall locations of one kernel in an ncu report: python scripts/ncu_hot.py rep.ncu-rep kernel_regex [n]"""
import csv, subprocess, sys, collections
rep, kre = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"], capture_output=True, text=True).stdout
lines = out.splitlines()
# several launches are concatenated; take the first block
blocks, cur = [], []
for l in lines:
    if l.startswith('"Kernel Name"'):
        if cur: blocks.append(cur)
        cur = []
    else:
        cur.append(l)
if cur: blocks.append(cur)
r = list(csv.reader(blocks[0])); h = r[0]
si = h.index("# Samples"); src = h.index("Source"); ie = h.index("Instructions Executed")
stalls = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
rows = []
tot = 0
for row in r[1:]:
    try: s = int(row[si])
    except: continue
    tot += s
    rows.append((s, row))
rows.sort(key=lambda x: -x[0])
print("total samples", tot)
for s, row in rows[:n]:
    top = sorted(((int(row[i] or 0), h[i]) for i in stalls), reverse=True)[:3]
    print(f"{s:7d} {100*s/tot:5.1f}%  {row[src][:70]:70s} exec={row[ie]:>9s} " + " ".join(f"{nm[6:]}={v}" for v, nm in top if v))
