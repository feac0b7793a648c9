#!/usr/bin/env python
"""Aggregate warp-stall samples of one kernel by reason, and by source line ranges: python scripts/ncu_stalls.py rep kernel_regex"""
import csv, subprocess, sys, collections
rep, kre = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"], capture_output=True, text=True).stdout
lines = out.splitlines()
blocks, cur = [], []
for l in lines:
    if l.startswith('"Kernel Name"'):
        if cur: blocks.append(cur)
        cur = []
    else:
        cur.append(l)
if cur: blocks.append(cur)
r = list(csv.reader(blocks[0])); h = r[0]
si = h.index("# Samples")
stalls = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
tot = collections.Counter(); n = 0
for row in r[1:]:
    try: s = int(row[si])
    except: continue
    n += s
    for i in stalls:
        try: tot[h[i]] += int(row[i] or 0)
        except: pass
print("samples", n)
for k, v in tot.most_common(14):
    print(f"  {k:28s} {v:7d} {100*v/max(n,1):5.1f}%")
