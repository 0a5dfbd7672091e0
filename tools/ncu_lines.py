#!/usr/bin/env python
"""Per-CUDA-source-line shares of executed warp instructions + key launch metrics for an .ncu-rep."""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hi = [i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r][0]
h = rows[hi]
ie, isamp, isrc, it = h.index("Instructions Executed"), h.index("# Samples"), h.index("Source"), h.index("Thread Instructions Executed")
body = [r for r in rows[hi + 1:] if len(r) > it and r[ie].isdigit() and r[isrc].strip() and r[0].strip().isdigit()]
tot = sum(int(r[ie]) for r in body); tt = sum(int(r[it]) for r in body); ts = sum(int(r[isamp]) for r in body if r[isamp].isdigit())
print(f"warp-instr {tot:,}  thread-instr {tt:,}  avg active threads {tt / tot:.1f}  samples {ts:,}")
for r in sorted(body, key=lambda r: -int(r[ie]))[:top]:
    print(f"{r[0]:>5} {r[isrc].strip()[:92]:92s} {int(r[ie]) / tot * 100:5.1f}% thr={int(r[it]) / max(1, int(r[ie])):4.1f} samp={(int(r[isamp]) if r[isamp].isdigit() else 0) / max(1, ts) * 100:4.1f}%")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines())); hh = rr[0]
for n in ("gpu__time_duration.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__warps_active.avg.per_cycle_active",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__average_warp_latency_per_inst_issued.ratio",
          "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__registers_per_thread",
          "dram__bytes_read.sum", "dram__bytes_write.sum"):
    if n in hh:
        print(f"  {n}: {[r[hh.index(n)] for r in rr[2:]]} {rr[1][hh.index(n)]}")
for i, n in enumerate(hh):
    if "issue_stalled" in n and "per_issue_active" in n and "not_issued" not in n:
        v = float(rr[2][i])
        if v > 0.3: print(f"  stall {n.split('issue_stalled_')[1].split('_per_issue')[0]}: {v:.2f}")
