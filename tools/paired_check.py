"""Paired diagnostic: CUDA path vs CPU oracle on the same events (different random numbers)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np

from attpc_engine_b200 import nuclear_map
from attpc_engine_b200.detector import simulate_batch
from oracle import attpc_oracle as oracle
from tests.common import load_golden, make_config

d = load_golden("distributions.npz")
name = "c16dd"
n = 120
cfg = make_config()
m, v = d[f"{name}/momenta"][:n], d[f"{name}/vertices"][:n]
zs, as_, idx = d[f"{name}/Z"], d[f"{name}/A"], list(d[f"{name}/indices"])
batch = simulate_batch(m, v, zs, as_, cfg, 1234, idx)
rng = np.random.default_rng(7)
rows = []
pooled_o, pooled_g = [], []
for i in range(n):
    rec = {}
    c, l = oracle.simulate_event(m[i], v[i], zs, as_, cfg, rng, idx, nuclear_map, record=rec,
                                 solver_kwargs=dict(rtol=1e-9, atol=1e-12, method="DOP853"))
    g, gl = batch.event(i)
    rows.append((len(c), len(g), c[:, 2].sum() if len(c) else 0, g[:, 2].sum() if len(g) else 0,
                 rec["stats"]["active_points"], rec["stats"]["primary_electrons"]))
    pooled_o.append(c[:, 2]); pooled_g.append(g[:, 2])
r = np.array(rows, dtype=float)
ok = r[:, 0] > 0
print("events", n, "oracle mean N", r[:, 0].mean(), "gpu mean N", r[:, 1].mean())
print("sum charge ratio gpu/oracle: mean", (r[ok, 3] / r[ok, 2]).mean(), "std", (r[ok, 3] / r[ok, 2]).std())
print("N ratio: mean", (r[ok, 1] / r[ok, 0]).mean(), "std", (r[ok, 1] / r[ok, 0]).std())
print("gpu stats", {k: batch.stats[k] for k in ("n_active_points", "n_primary_electrons", "n_trajectory_points")},
      "oracle active", r[:, 4].sum(), "oracle primary", r[:, 5].sum())
po, pg = np.concatenate(pooled_o), np.concatenate(pooled_g)
qs = [0.01, 0.05, 0.1, 0.25, 0.5, 0.75, 0.9, 0.99]
print("quantiles oracle", np.quantile(po, qs))
print("quantiles gpu   ", np.quantile(pg, qs))
print("frac zero-charge points oracle", (po == 0).mean(), "gpu", (pg == 0).mean())
