#!/usr/bin/env python
"""Where a kernel's warps spend their time, by source phase: aggregates `ncu --page source --print-source cuda,sass --csv`
(stall samples and executed instructions per CUDA line) over the `// ----------------` phase markers of a .cuh file.

  ncu -i prof.ncu-rep --page source --print-source cuda,sass --csv > prof_cs.csv
  python tools/ncu_phase_report.py prof_cs.csv hcr_genesis_lr_cl_b200/csrc/dynamics_kernel.cuh [top_lines]
"""
import csv
import os
import re
import sys


def main():
    path, src_path = sys.argv[1], sys.argv[2]
    top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    rows = list(csv.reader(open(path)))
    cur, agg = None, {}
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur = os.path.basename(r[1])
            continue
        if len(r) < 8 or r[2] != "-":
            continue
        try:
            ln = int(r[0])
        except ValueError:
            continue
        agg[(cur, ln)] = (int(r[6] or 0), int(r[7] or 0), r[1][:100])
    tot_s = sum(v[0] for v in agg.values()) or 1
    tot_i = sum(v[1] for v in agg.values()) or 1
    src = open(src_path).read().split("\n")
    base = os.path.basename(src_path)
    marks = [(i + 1, re.sub(r"[-=/ ]+$", "", ln.strip().lstrip("/ -=")).strip()[:48]) for i, ln in enumerate(src) if re.match(r"\s*// [-=]{8,} ", ln)]
    marks = [(1, "(file head / helpers)")] + marks + [(len(src) + 1, "end")]
    ph = {}
    other = [0, 0]
    for (f, ln), (s, i, _) in agg.items():
        if f != base:
            other[0] += s
            other[1] += i
            continue
        name = next(m[1] for k, m in enumerate(marks[:-1]) if m[0] <= ln < marks[k + 1][0])
        a = ph.setdefault(name, [0, 0])
        a[0] += s
        a[1] += i
    print(f"total: {tot_s} samples, {tot_i} warp-instructions")
    for _, name in marks[:-1]:
        if name in ph:
            s, i = ph[name]
            print(f"  {name:50s} samples {100 * s / tot_s:5.1f}%   instructions {100 * i / tot_i:5.1f}%")
    print(f"  {'(other files: inlined helpers, intrinsics)':50s} samples {100 * other[0] / tot_s:5.1f}%   instructions {100 * other[1] / tot_i:5.1f}%")
    print("hottest lines:")
    for (f, ln), (s, i, t) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top_n]:
        print(f"  {f}:{ln:<5d} {100 * s / tot_s:5.1f}% samples  {100 * i / tot_i:5.1f}% instr | {t}")


if __name__ == "__main__":
    main()
