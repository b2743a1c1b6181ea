#!/usr/bin/env python
"""Per-source-line warp-stall samples of one ncu report (needs --import-source on and -lineinfo).
  python tools/ncu_lines.py <report.ncu-rep> [top_n]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, H = None, None
agg = {}
tot = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        H = r
        si = H.index("# Samples")
        ie = H.index("Instructions Executed")
        stall = [(i, h) for i, h in enumerate(H) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if H is None or r[0] in ("Function Name",) or not r[0].isdigit():
        continue
    try:
        n = int(r[si])
    except ValueError:
        continue
    key = (cur_file, int(r[0]))
    a = agg.setdefault(key, {"n": 0, "inst": 0, "src": r[1].strip()[:110], "st": collections.Counter()})
    a["n"] += n
    a["inst"] += int(r[ie] or 0)
    for i, h in stall:
        try:
            a["st"][h[6:]] += int(r[i])
        except ValueError:
            pass
    tot += n
print(f"total samples {tot}")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["n"])[:top]:
    st = ", ".join(f"{k}={v}" for k, v in a["st"].most_common(4) if v)
    print(f"{a['n']:6d} {100 * a['n'] / max(tot, 1):5.1f}%  inst={a['inst']:9d}  {key[0]}:{key[1]:<5d} {a['src']}\n        {st}")
