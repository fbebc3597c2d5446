#!/usr/bin/env python
"""Per source line: stall samples by reason, from `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`.
Usage: python tools/ncu_stall_lines.py prof_cs.csv [top_n] [reason]"""
import csv
import os
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
reason = sys.argv[3] if len(sys.argv) > 3 else None
cur, hdr = None, None
agg = defaultdict(lambda: defaultdict(int))
src = {}
for r in csv.reader(open(path)):
    if len(r) >= 2 and r[0] == "File Path":
        cur = os.path.basename(r[1]); continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-":
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    src[(cur, ln)] = r[1][:110]
    for k, name in enumerate(hdr):
        if name.startswith("stall_") and "Not Issued" not in name:
            agg[(cur, ln)][name] += int(r[k] or 0)
    agg[(cur, ln)]["inst"] += int(r[7] or 0)
tot = defaultdict(int)
for v in agg.values():
    for k, x in v.items():
        tot[k] += x
allsamp = sum(x for k, x in tot.items() if k.startswith("stall_"))
print("totals:", {k: x for k, x in sorted(tot.items(), key=lambda t: -t[1])})
key = (lambda kv: -kv[1][reason]) if reason else (lambda kv: -sum(x for k, x in kv[1].items() if k.startswith("stall_")))
for (f, ln), v in sorted(agg.items(), key=key)[:top]:
    s = sum(x for k, x in v.items() if k.startswith("stall_"))
    topr = sorted(((x, k[6:]) for k, x in v.items() if k.startswith("stall_") and x), reverse=True)[:3]
    print(f"{f}:{ln:<5d} {100.0 * s / allsamp:5.1f}% inst {v['inst']:>8d}  {topr}  | {src[(f, ln)]}")
