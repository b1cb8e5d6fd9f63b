#!/usr/bin/env python
"""Per-source-line summary of an `ncu --page source --csv --print-source cuda,sass` export:
samples, executed warp instructions and the dominant stall reasons."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10**9
hdr = next(r for r in rows if r and r[0] == "Line No")
col = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
tot = 0
out = []
for r in rows:
    if len(r) < len(hdr) or not r[0].isdigit():
        continue
    ln = int(r[0])
    s = int(r[col["# Samples"]] or 0)
    tot += s
    if lo <= ln <= hi and s:
        inst = int(r[col["Instructions Executed"]] or 0)
        top = sorted(((int(r[col[n]] or 0), n[6:]) for n in stalls), reverse=True)[:3]
        out.append((ln, s, inst, ", ".join(f"{n}:{v}" for v, n in top if v), r[1].strip()[:70]))
for ln, s, inst, top, src in out:
    print(f"{ln:5d} {s:7d} {100.0 * s / tot:5.1f}% inst={inst:9d}  {top:40s} | {src}")
print("total samples", tot)
