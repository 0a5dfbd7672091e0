"""Statistical half of parity part (b): the CUDA path against distribution samples of the UNMODIFIED reference
(tests/golden/make_distributions.py: 6000 / 4000 / 2400 / 2400 events of the four bench workloads).

The CUDA events come from kinematics drawn with a DIFFERENT seed than the reference sample (unpaired): a two-sample
Kolmogorov-Smirnov test at p > 0.01 per observable, north_star's bar.  Observables: cloud points per event, charge per
event, time-bucket extent, pads per event, median / maximum charge per point, charge of one random point, charge of
one random pad ("charge per pad"), cloud points per track (label), trajectory rows per track and path length per
track ("track-length and point-count distributions").

Also here: the production-only early stop of stalled ions (`inert_forever`) must not change a single row.
"""

import numpy as np
import pytest
from scipy.stats import ks_2samp

from attpc_engine_b200 import nuclear_map
from tests.common import WORKLOAD_NAMES, load_golden

pytestmark = pytest.mark.gpu

P_MIN = 0.01  # north_star: KS tests against the reference at p > 0.01
N_GPU = {"c16dd": 6000, "c14dp": 4000, "c12aa": 2400, "sn132dp": 2400}
STRIDE = 4  # tests/golden/make_distributions.py: STRIDE


@pytest.fixture(scope="module")
def dist():
    return load_golden("distributions.npz")


def _event_observables(batch, indices, seed):
    picker = np.random.default_rng(seed)
    n = len(batch)
    obs = {k: [] for k in ("n_points", "sum_charge", "tb_extent", "n_pads", "median_charge", "max_charge",
                           "point_charge_sample", "pad_charge_sample")}  # fmt: skip
    per_label = np.zeros((n, len(indices)), dtype=np.int64)
    for e in range(n):
        cloud, labels = batch.event(e)
        for k, idx in enumerate(indices):
            per_label[e, k] = int((labels == idx).sum())
        if len(cloud) == 0:
            for k, v in zip(("n_points", "sum_charge", "tb_extent", "n_pads", "median_charge", "max_charge"),
                            (0, 0.0, 0.0, 0, 0.0, 0.0)):  # fmt: skip
                obs[k].append(v)
            continue
        q = cloud[:, 2]
        pads = cloud[:, 0].astype(np.int64)
        obs["n_points"].append(len(cloud))
        obs["sum_charge"].append(q.sum())
        obs["tb_extent"].append(cloud[:, 1].max() - cloud[:, 1].min())
        obs["n_pads"].append(len(np.unique(pads)))
        obs["median_charge"].append(np.median(q))
        obs["max_charge"].append(q.max())
        obs["point_charge_sample"].append(q[picker.integers(len(q))])
        obs["pad_charge_sample"].append(q[pads == pads[picker.integers(len(pads))]].sum())
    return {k: np.asarray(v, dtype=np.float64) for k, v in obs.items()}, per_label


@pytest.mark.parametrize("name", WORKLOAD_NAMES)
def test_distributions_match_reference_unpaired(dist, name):
    import bench
    from attpc_engine_b200.detector import simulate_batch
    from attpc_engine_b200.detector.engine import engine_for
    from attpc_engine_b200.detector.simulator import _nuclei_for

    n = N_GPU[name]
    cfg, momenta, vertices, zs, as_, indices = bench.build_workload(name, n, seed_offset=1)  # not the reference's events
    batch = simulate_batch(momenta, vertices, zs, as_, cfg, 20261018, indices, columns=True)
    obs, per_label = _event_observables(batch, indices, 5)
    failures = []

    def check(what, ours, theirs):
        p = ks_2samp(ours, theirs).pvalue
        if not p > P_MIN:
            failures.append(f"{what}: p = {p:.3g} (n = {len(ours)} vs {len(theirs)})")

    for key, ours in obs.items():
        check(key, ours, dist[f"{name}/{key}"])
    charged = [k for k, idx in enumerate(indices) if zs[idx] != 0]
    for k in charged:
        check(f"points_per_track[{indices[k]}]", per_label[:, k], dist[f"{name}/points_per_track"][:, k])
    # trajectories: rows on the 0.1 ns grid and path length per track; freeze_ke_mev = 0 integrates a stalled ion to
    # 1 us like the reference, so that the row counts mean the same thing on both sides
    nuclei = _nuclei_for(zs, as_, indices, nuclear_map)
    eng = engine_for(cfg, nuclei, freeze_ke_mev=0.0)
    n_traj = min(n, 1500)
    max_rows = 10001 // STRIDE + 1
    for t, k in enumerate(charged):
        nucleus = nuclear_map.get_data(int(zs[indices[k]]), int(as_[indices[k]]))
        rows, length = [], []
        for a in range(0, n_traj, 250):
            b = min(a + 250, n_traj)
            pts, counts = eng.trajectories(momenta[a:b, indices[k]], vertices[a:b], [nucleus] * (b - a), stride=STRIDE,
                                           max_points=max_rows)  # fmt: skip
            rows.append(counts)
            for i, c in enumerate(counts):
                m = (int(c) - 1) // STRIDE + 1  # rows 0, 4, 8, ... < c
                seg = np.diff(pts[i, :m, :3], axis=0)
                length.append(np.sqrt((seg**2).sum(axis=1)).sum() if m > 1 else 0.0)
        check(f"traj_rows[{indices[k]}]", np.concatenate(rows).astype(np.float64), dist[f"{name}/traj_rows"][:, t])
        check(f"path_length[{indices[k]}]", np.asarray(length), dist[f"{name}/path_length"][:, t])
    assert not failures, f"{name}: " + "; ".join(failures)


@pytest.mark.parametrize("name", WORKLOAD_NAMES)
def test_freeze_changes_nothing(name):
    """Draws are keyed by grid step, so ending a stalled ion early (`inert_forever`, the default) and integrating it
    to 1 us like the reference (`freeze_ke_mev=0`) must give IDENTICAL clouds: 10 000 events per workload."""
    import bench
    from attpc_engine_b200.detector import simulate_batch

    total, per = 10000, 2500
    cfg, momenta, vertices, zs, as_, indices = bench.build_workload(name, total, seed_offset=2)
    rows = 0
    for a in range(0, total, per):
        sl = slice(a, a + per)
        fast = simulate_batch(momenta[sl], vertices[sl], zs, as_, cfg, 99, indices, first_event=a, columns=True)
        full = simulate_batch(momenta[sl], vertices[sl], zs, as_, cfg, 99, indices, first_event=a, columns=True,
                              freeze_ke_mev=0.0)  # fmt: skip
        assert full.stats["n_trajectory_points"] > fast.stats["n_trajectory_points"]  # the tails really were cut
        assert np.array_equal(fast.offsets, full.offsets), (name, a)
        assert fast.columns.keys() == full.columns.keys()
        for key in fast.columns:
            assert np.array_equal(fast.columns[key], full.columns[key]), (name, a, key)
        rows += len(fast.columns["pad"])
    assert rows > 0
