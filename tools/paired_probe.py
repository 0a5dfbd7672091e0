#!/usr/bin/env python
"""Per-event observables of the CUDA path on the events tools/tolerance_study.py ran through the oracle (GPU box)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from attpc_engine_b200.detector import simulate_batch  # noqa: E402

name, n = sys.argv[1], int(sys.argv[2])
cfg, m, v, zs, as_, idx = bench.build_workload(name, 4096)
rows = []
for seed in (1, 2, 3):
    b = simulate_batch(m[:n], v[:n], zs, as_, cfg, seed, idx, columns=True)
    out = np.zeros((n, 4))
    for e in range(n):
        c, _ = b.event(e)
        q = c[:, 2] if len(c) else np.zeros(1)
        out[e] = (len(c), q.sum(), np.median(q), q.max())
    rows.append(out)
np.save(ROOT / "gpurun_out" / f"gpu_obs_{name}.npy", np.stack(rows))
print("saved", np.stack(rows).shape)
