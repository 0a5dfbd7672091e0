"""GPU parity, part (b): trajectories of the CUDA Dormand-Prince integrator against scipy `solve_ivp` driving the
reference's own equation of motion and terminal events (tests/golden/make_golden.py: make_tight_trajectories).

Tolerance (north_star): 1e-4 relative in position and energy.  Position errors are taken relative to the path
length travelled so far (floored at 1 cm), energy errors relative to the initial kinetic energy."""

import numpy as np
import pytest

from attpc_engine_b200 import nuclear_map
from tests.common import make_config

pytestmark = pytest.mark.gpu

CASES = ["d_exit", "d_stop", "d_loop", "c16_fwd", "p_back", "alpha"]
RTOL = 1e-4


def _ke(rows, mass):
    g2 = np.sum(rows[:, 3:] ** 2, axis=1)
    return mass * g2 / (np.sqrt(1.0 + g2) + 1.0)


def _engine(nuclei, **kw):
    from attpc_engine_b200.detector.engine import engine_for

    return engine_for(make_config(), nuclei, **kw)


@pytest.mark.parametrize("name", CASES)
def test_trajectory_matches_converged_reference(golden_traj, name):
    g = golden_traj
    z, a = (int(v) for v in g[f"traj/{name}/za"])
    nucleus = nuclear_map.get_data(z, a)
    eng = _engine([nucleus], freeze_ke_mev=0.0)
    want = g[f"traj/{name}/tight_every8"]
    pts, counts = eng.trajectories(g[f"traj/{name}/momentum"], g[f"traj/{name}/vertex"], [nucleus], stride=8,
                                   max_points=len(want))  # fmt: skip
    n_ref = int(g[f"traj/{name}/npoints"][1])
    # same number of 0.1 ns grid points as scipy emits (a terminal event may fall within rounding of a grid point)
    assert abs(int(counts[0]) - n_ref) <= 1
    n = min(len(want), (int(counts[0]) - 1) // 8 + 1)
    got, want = pts[0, :n], want[:n]
    path = np.concatenate([[0.0], np.cumsum(np.linalg.norm(np.diff(want[:, :3], axis=0), axis=1))])
    pos_err = np.linalg.norm(got[:, :3] - want[:, :3], axis=1) / np.maximum(path, 0.01)
    ke_got, ke_want = _ke(got, nucleus.mass), _ke(want, nucleus.mass)
    ke_err = np.abs(ke_got - ke_want) / ke_want[0]
    assert pos_err.max() < RTOL, f"position error {pos_err.max():.2e}"
    assert ke_err.max() < RTOL, f"energy error {ke_err.max():.2e}"


@pytest.mark.parametrize("name", CASES)
def test_distance_to_default_radau_is_reported(golden_traj, name):
    """Not a pass/fail parity statement: the reference's DEFAULT tolerances (rtol 1e-3) are themselves ~1e-4..1e-3
    away from the converged solution; this pins that the CUDA path is at least as close to the
    converged solution as the reference's default run is."""
    g = golden_traj
    z, a = (int(v) for v in g[f"traj/{name}/za"])
    nucleus = nuclear_map.get_data(z, a)
    eng = _engine([nucleus], freeze_ke_mev=0.0)
    tight, default = g[f"traj/{name}/tight_every8"], g[f"traj/{name}/default_every8"]
    n = min(len(tight), len(default))
    pts, _ = eng.trajectories(g[f"traj/{name}/momentum"], g[f"traj/{name}/vertex"], [nucleus], stride=8, max_points=n)
    ours = np.linalg.norm(pts[0, :n, :3] - tight[:n, :3], axis=1).max()
    ref_default = np.linalg.norm(default[:n, :3] - tight[:n, :3], axis=1).max()
    assert ours <= max(ref_default, 1e-7)


def test_freeze_only_drops_inert_tail(golden_traj):
    """With the default freeze threshold a stopped ion ends early; the points before that are unchanged."""
    g = golden_traj
    nucleus = nuclear_map.get_data(1, 2)
    mom, vtx = g["traj/d_stop/momentum"], g["traj/d_stop/vertex"]
    full, n_full = _engine([nucleus], freeze_ke_mev=0.0).trajectories(mom, vtx, [nucleus], stride=1, max_points=4000)
    cut, n_cut = _engine([nucleus]).trajectories(mom, vtx, [nucleus], stride=1, max_points=4000)
    assert n_full[0] == 10001 and 10 < n_cut[0] < 4000
    m = int(n_cut[0])
    assert np.array_equal(full[0, :m], cut[0, :m])
    ke_tail = _ke(full[0, m - 1 : 4000], nucleus.mass)
    assert np.all(np.abs(np.diff(ke_tail)) * 1e6 / 34.0 < 0.059)  # < n* electrons per step: can never fire
