#!/usr/bin/env python
"""Per-source-line summary of an `ncu --page source --csv --print-source cuda,sass` export:
samples, executed warp instructions and the dominant stall reasons.

    python tools/ncu_lines.py export.csv [file-substring] [top-N]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else ""
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40


def num(s):
    try:
        return int(float(s))
    except ValueError:
        return 0


cur_file, hdr, col, stalls = "", None, None, None
agg = {}
tot = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        col = {n: i for i, n in enumerate(hdr)}
        stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
        continue
    s = num(r[col["# Samples"]])
    tot += s
    if want and want not in cur_file:
        continue
    key = (cur_file.rsplit("/", 1)[-1], int(r[0]))
    a = agg.setdefault(key, {"s": 0, "i": 0, "st": {}, "src": r[1].strip()[:80]})
    a["s"] += s
    a["i"] += num(r[col["Instructions Executed"]])
    for n in stalls:
        a["st"][n[6:]] = a["st"].get(n[6:], 0) + num(r[col[n]])
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1]["s"])[:top_n]:
    top = sorted(((v, n) for n, v in a["st"].items()), reverse=True)[:3]
    print(f"{f}:{ln:<5d} {a['s']:7d} {100.0 * a['s'] / max(tot, 1):5.1f}% inst={a['i']:9d}  "
          f"{', '.join(f'{n}:{v}' for v, n in top if v):42s} | {a['src']}")
print("total samples", tot)
