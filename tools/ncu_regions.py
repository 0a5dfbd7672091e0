#!/usr/bin/env python
"""Executed warp instructions of an .ncu-rep summed over source-line ranges: tools/ncu_regions.py rep lo-hi:name ..."""
import csv, subprocess, sys
rep = sys.argv[1]
ranges = []
for a in sys.argv[2:]:
    r, name = a.split(":")
    lo, hi = r.split("-")
    ranges.append((int(lo), int(hi), name))
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hi_ = [i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r][0]
h = rows[hi_]
ie, isamp, it = h.index("Instructions Executed"), h.index("# Samples"), h.index("Thread Instructions Executed")
body = [r for r in rows[hi_ + 1:] if len(r) > it and r[ie].isdigit() and r[0].strip().isdigit()]
tot = sum(int(r[ie]) for r in body); ts = sum(int(r[isamp]) for r in body if r[isamp].isdigit())
print(f"total warp-instr {tot:,} samples {ts:,}")
for lo, hi, name in ranges:
    sel = [r for r in body if lo <= int(r[0]) <= hi]
    e = sum(int(r[ie]) for r in sel); t = sum(int(r[it]) for r in sel); s = sum(int(r[isamp]) for r in sel if r[isamp].isdigit())
    print(f"{name:28s} lines {lo}-{hi}: {e / tot * 100:5.1f}% instr  thr={t / max(e, 1):4.1f}  samples {s / max(ts, 1) * 100:5.1f}%")
