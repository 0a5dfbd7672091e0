#!/usr/bin/env python
"""KS p-values of every observable of tests/test_gpu_statistics.py for several independent CUDA samples (GPU box)."""
import sys
from pathlib import Path

import numpy as np
from scipy.stats import ks_2samp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from attpc_engine_b200.detector import simulate_batch  # noqa: E402
from tests.common import WORKLOAD_NAMES, load_golden  # noqa: E402
from tests.test_gpu_statistics import N_GPU, _event_observables  # noqa: E402

dist = load_golden("distributions.npz")
names = sys.argv[1].split(",") if len(sys.argv) > 1 else WORKLOAD_NAMES
for name in names:
    for trial in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3):
        n = N_GPU[name]
        cfg, m, v, zs, as_, idx = bench.build_workload(name, n, seed_offset=10 + trial)
        batch = simulate_batch(m, v, zs, as_, cfg, 777 + trial, idx, columns=True)
        obs, per_label = _event_observables(batch, idx, 50 + trial)
        line = [f"{k}={ks_2samp(o, dist[f'{name}/{k}']).pvalue:.3f}" for k, o in obs.items()]
        line += [f"ppt{idx[k]}={ks_2samp(per_label[:, k], dist[f'{name}/points_per_track'][:, k]).pvalue:.3f}"
                 for k in range(len(idx)) if zs[idx[k]] != 0]
        print(name, trial, " ".join(line), flush=True)
