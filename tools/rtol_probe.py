"""Integrator tolerance study (run under gpurun): error against the converged scipy solution and step counts."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

import bench
from attpc_engine_b200 import nuclear_map
from attpc_engine_b200.detector.engine import engine_for
from attpc_engine_b200.detector.simulator import _nuclei_for
from tests.common import make_config

g = np.load(Path(__file__).resolve().parent.parent / "tests" / "golden" / "trajectories.npz")
CASES = ["d_exit", "d_stop", "d_loop", "c16_fwd", "p_back", "alpha"]


def ke_of(rows, mass):
    g2 = np.sum(rows[:, 3:] ** 2, axis=1)
    return mass * g2 / (np.sqrt(1.0 + g2) + 1.0)


for rtol in (1e-8, 1e-7, 1e-6, 1e-5):
    worst_p = worst_k = 0.0
    line = []
    for name in CASES:
        z, a = (int(v) for v in g[f"traj/{name}/za"])
        nucleus = nuclear_map.get_data(z, a)
        eng = engine_for(make_config(), [nucleus], freeze_ke_mev=0.0, ode_rtol=rtol, ode_atol=rtol * 1e-4)
        want = g[f"traj/{name}/tight_every8"]
        pts, counts = eng.trajectories(g[f"traj/{name}/momentum"], g[f"traj/{name}/vertex"], [nucleus], stride=8, max_points=len(want))
        n_ref = int(g[f"traj/{name}/npoints"][1])
        n = min(len(want), (int(counts[0]) - 1) // 8 + 1)
        got, w = pts[0, :n], want[:n]
        path = np.concatenate([[0.0], np.cumsum(np.linalg.norm(np.diff(w[:, :3], axis=0), axis=1))])
        pe = (np.linalg.norm(got[:, :3] - w[:, :3], axis=1) / np.maximum(path, 0.01)).max()
        kw = ke_of(w, nucleus.mass)
        kerr = (np.abs(ke_of(got, nucleus.mass) - kw) / kw[0]).max()
        line.append(f"{name}: dn={int(counts[0]) - n_ref} pos={pe:.1e} ke={kerr:.1e}")
        worst_p, worst_k = max(worst_p, pe), max(worst_k, kerr)
    print(f"rtol={rtol:g} worst pos={worst_p:.2e} ke={worst_k:.2e} | " + " | ".join(line), flush=True)

for name in ("c16dd", "c14dp"):
    config, momenta, vertices, zs, as_, indices = bench.build_workload(name, 32768)
    K = momenta.shape[1]
    m = torch.from_numpy(momenta).cuda()
    v = torch.from_numpy(vertices).cuda()
    for rtol in (1e-8, 1e-7, 1e-6, 1e-5):
        eng = engine_for(config, _nuclei_for(zs, as_, indices, nuclear_map), max_events_per_launch=32768, ode_rtol=rtol, ode_atol=rtol * 1e-4)
        for n in (512, 32768):
            for rep in range(3):
                st = eng.simulate_device(m.data_ptr(), v.data_ptr(), n, K, zs, as_, indices, seed=1 + rep).stats
            print(name, "rtol", rtol, "events", n, "ms_tracks", round(st["ms_tracks"], 3), "steps/track", round(st["n_rk_steps"] / st["n_tracks"], 1),
                  "rejects/track", round(st["n_rk_rejects"] / st["n_tracks"], 2), "max passes", st["max_track_passes"],
                  "traj pts", st["n_trajectory_points"], "active", st["n_active_points"], "prim e", st["n_primary_electrons"], flush=True)
