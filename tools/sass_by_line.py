#!/usr/bin/env python
"""Static SASS instruction count per source line of one kernel (needs -lineinfo).

    python tools/sass_by_line.py <cubin> <kernel-substring> [source-file-substring]
"""
import collections
import re
import subprocess
import sys

cubin, kern = sys.argv[1], sys.argv[2]
srcsub = sys.argv[3] if len(sys.argv) > 3 else ""
out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
counts = collections.Counter()
ops = collections.defaultdict(collections.Counter)
active, cur = False, None
for ln in out.splitlines():
    if ln.startswith(".text.") or ln.startswith("\t.section\t.text."):
        active = kern in ln
        continue
    if not active:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and cur:
        if srcsub and srcsub not in cur[0]:
            continue
        counts[cur] += 1
        ops[cur][m.group(1).split(".")[0]] += 1
for (f, l), c in sorted(counts.items()):
    top = ", ".join(f"{k}:{v}" for k, v in ops[(f, l)].most_common(4))
    print(f"{f}:{l:5d} {c:5d}  {top}")
print("total", sum(counts.values()))
