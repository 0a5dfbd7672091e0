"""Summarise bench JSON lines of gpurun_out/<tag>_*.log: step time, value, stage split, e2e."""
import glob
import json
import sys

for f in sorted(glob.glob(f"gpurun_out/{sys.argv[1]}_*.log")):
    for line in open(f):
        if line.startswith("{"):
            d = json.loads(line)
            r = d.get("roofline") or {}
            st = r.get("stage_ms_per_step") or {}
            print(f.split("/")[-1][len(sys.argv[1]) + 1:-4].ljust(30), d.get("ms_per_step"), round(d.get("value", 0) / 1e6, 3),
                  [round(v, 2) for v in st.values()], "e2e", (d.get("e2e") or {}).get("value"))
