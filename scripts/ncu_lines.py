#!/usr/bin/env python
"""Stall samples per CUDA source line: python scripts/ncu_lines.py rep.ncu-rep kernel_regex [n]"""
import csv, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = []
fname = None; hdr = None; seen_fn = 0
for row in csv.reader(out.splitlines()):
    if not row: continue
    if row[0] == "File Path": fname = row[1].split("/")[-1]; continue
    if row[0] == "Function Name":
        continue
    if row[0] == "Line No": hdr = row; continue
    if hdr is None: continue
    if row[2] != "-":      # sass rows carry an address; keep only the per-line aggregate rows
        continue
    try:
        s = int(row[hdr.index("# Samples")])
    except Exception:
        continue
    stalls = {hdr[i][6:]: int(row[i] or 0) for i in range(len(hdr)) if hdr[i].startswith("stall_") and "Not Issued" not in hdr[i]}
    rows.append((s, fname, row[0], row[1].strip()[:90], stalls, row[hdr.index("Instructions Executed")]))
# the report holds several launches of the kernel: lines repeat; merge
agg = {}
for s, f, ln, src, st, ie in rows:
    k = (f, ln, src)
    if k not in agg: agg[k] = [0, {}, 0]
    agg[k][0] += s
    agg[k][2] += int(ie or 0)
    for a, b in st.items(): agg[k][1][a] = agg[k][1].get(a, 0) + b
tot = sum(v[0] for v in agg.values())
print("total samples", tot)
for (f, ln, src), (s, st, ie) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n]:
    top = sorted(st.items(), key=lambda x: -x[1])[:3]
    print(f"{100*s/tot:5.1f}% {f}:{ln:>4s} {src:90s} inst={ie} " + " ".join(f"{a}={b}" for a, b in top if b))
