#!/usr/bin/env python
"""Does the reference's loose ODE tolerance (scipy Radau, rtol 1e-3) move the charge observables?

Runs the CPU oracle (oracle/attpc_oracle.py, pinned to the reference) on the same events and the same random
numbers twice: with the reference's solver settings and with a converged trajectory (DOP853, rtol 1e-10), and
prints the paired ratio of the per-event observables plus the two-sample KS p-value between the two runs.
    python tools/tolerance_study.py c16dd 1600
"""
import multiprocessing as mp
import os
import sys
from pathlib import Path

import numpy as np
from scipy.stats import ks_2samp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
_S = {}


def _run(args):
    name, a, b = args
    import bench
    from attpc_engine_b200 import nuclear_map
    from oracle import attpc_oracle as oracle

    if name not in _S:
        _S[name] = bench.build_workload(name, 4096)
    cfg, m, v, zs, as_, idx = _S[name]
    out = []
    for i in range(a, b):
        row = []
        for kw in (None, dict(method="DOP853", rtol=1e-10, atol=1e-13)):
            cloud, _ = oracle.simulate_event(m[i], v[i], zs, as_, cfg, np.random.default_rng(1000 + i), idx, nuclear_map,
                                             solver_kwargs=kw)
            q = cloud[:, 2] if len(cloud) else np.zeros(1)
            row += [len(cloud), q.sum(), np.median(q), q.max()]
        out.append(row)
    return out


if __name__ == "__main__":
    name, n = sys.argv[1], int(sys.argv[2])
    jobs = [(name, a, min(a + 20, n)) for a in range(0, n, 20)]
    with mp.get_context("spawn").Pool(len(os.sched_getaffinity(0))) as pool:
        rows = np.array([r for part in pool.map(_run, jobs) for r in part], dtype=np.float64)
    np.save(ROOT / "gpurun_out" / f"oracle_obs_{name}.npy", rows)
    for k, label in enumerate(("n_points", "sum_charge", "median_charge", "max_charge")):
        a, b = rows[:, k], rows[:, 4 + k]
        ok = a > 0
        print(f"{label:14s} reference-settings mean {a.mean():.6g}  converged mean {b.mean():.6g}  "
              f"paired ratio converged/reference: mean {np.mean(b[ok] / a[ok]):.4f} median {np.median(b[ok] / a[ok]):.4f}  "
              f"KS p = {ks_2samp(a, b).pvalue:.3g}")
