#!/usr/bin/env python
"""Top SASS lines of an .ncu-rep by stall samples and by executed instructions, with source line mapping."""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hi = [i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r]
if not hi:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines())); hi = [i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r]
h = rows[hi[0]]
ie, isamp, ia, isrc = h.index("Instructions Executed"), h.index("# Samples"), h.index("Avg. Threads Executed"), h.index("Source")
body = [r for r in rows[hi[0] + 1:] if len(r) > ia and r[ie].isdigit()]
tot_i = sum(int(r[ie]) for r in body); tot_s = sum(int(r[isamp]) for r in body)
print(f"total warp-instr {tot_i:,}  samples {tot_s:,}")
stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_")]
agg = collections.Counter()
for r in body:
    for i in stall_cols:
        if r[i].isdigit(): agg[h[i]] += int(r[i])
print("stall totals:", ", ".join(f"{k}={v}" for k, v in agg.most_common(8)))
print("-- by samples")
for r in sorted(body, key=lambda r: -int(r[isamp]))[:top]:
    print(f"{r[0][-5:]} {r[isrc][:78]:78s} inst={int(r[ie]):>10,} samp={int(r[isamp]):>8,} thr={r[ia]}")
print("-- by instructions")
for r in sorted(body, key=lambda r: -int(r[ie]))[:top]:
    print(f"{r[0][-5:]} {r[isrc][:78]:78s} inst={int(r[ie]):>10,} samp={int(r[isamp]):>8,} thr={r[ia]}")
