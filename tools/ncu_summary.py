#!/usr/bin/env python
"""Summarise gpurun_out/launches.csv and prof_*.ncu-rep into text (needs ncu on PATH, no GPU)."""
import collections
import csv
import subprocess
import sys
from pathlib import Path

import os

OUT = Path(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out")
TAG = os.environ.get("TAG", "")  # file prefix of tools/profile_r2.sh, e.g. TAG=r2p -> r2p_launches.csv, r2p_prof_*.ncu-rep
PRE = TAG + "_" if TAG else ""
WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_read.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
    "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
    "smsp__warp_issue_stalled_membar_per_warp_active.pct", "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct",
    "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
]  # fmt: skip


def launches():
    f = OUT / f"{PRE}launches.csv"
    if not f.exists():
        return
    rows = list(csv.reader(open(f)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1 :]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(r[ui], v)
        name = r[ki].split("(")[0]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"== launch list ({f}): cold-cache, serialised; compare shares")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:70]:70s} n={v[0]:4d} total={v[1]:10.1f} us avg={v[1] / v[0]:9.1f} us share={v[1] / tot * 100:5.1f}%")


TRAFFIC = {}


def reports():
    for rep in sorted(OUT.glob(f"{PRE}prof_*.ncu-rep")):
        txt = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        if len(rows) < 3:
            continue
        h, units = rows[0], rows[1]
        print(f"== {rep.name}: {rows[2][h.index('Kernel Name')][:90]}")
        try:
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd, wr = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
            total = float(rows[2][rd].replace(",", "")) * scale[units[rd]] + float(rows[2][wr].replace(",", "")) * scale[units[wr]]
            name = rep.name[len(PRE + "prof_"):-len(".ncu-rep")]
            TRAFFIC[name] = {
                "dram_bytes_per_launch": total,
                "duration_us": rows[2][h.index("gpu__time_duration.sum")] + " " + units[h.index("gpu__time_duration.sum")],
                "issue_active_pct": round(float(rows[2][h.index("smsp__issue_active.avg.pct_of_peak_sustained_active")]), 1),
                "grid": rows[2][h.index("launch__grid_size")],
                "warp_inst_per_launch": float(rows[2][h.index("smsp__inst_executed.sum")].replace(",", "")),
            }
        except (ValueError, KeyError) as exc:
            print("  (no traffic record:", exc, ")")
        for w in WANT:
            if w in h:
                i = h.index(w)
                print(f"  {w:75s} {units[i]:14s} {[r[i] for r in rows[2:]]}")


launches()
reports()

if len(sys.argv) > 2:  # python tools/ncu_summary.py gpurun_out profiles/traffic.json [workload events_per_launch source]
    import json

    workload = sys.argv[3] if len(sys.argv) > 3 else "c16dd"
    events = int(sys.argv[4]) if len(sys.argv) > 4 else 32768
    source = sys.argv[5] if len(sys.argv) > 5 else "ncu --set full"
    for rec in TRAFFIC.values():
        rec.update(workload=workload, events_per_launch=events, source=source)
    Path(sys.argv[2]).write_text(json.dumps(TRAFFIC, indent=1) + "\n")
