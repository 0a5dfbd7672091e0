"""CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A CPU restatement (scipy + numba + numpy, the reference's own tool chain) of the detector
hot path of ATTPC/attpc_engine v0.9.0.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this module; nothing under
`attpc_engine_b200/` does, and the product path raises when its CUDA library is missing.

Pinning: `tests/test_oracle_golden.py` checks every function here against fixtures under
`tests/golden/` that were produced by running the UNMODIFIED reference in the build
container (`tests/golden/make_golden.py`, via `tests/golden/ref_shim.py`), plus the
reference's own known-answer vectors for Szudzik pairing (`tests/test_pairing.py:4-26`).
The one thing that cannot be pinned is CATIMA's dE/dx (pycatima 1.96 via spyral-utils 2.0.0,
not vendored, not installed): oracle, reference-under-shim and CUDA path all read the same
`attpc_engine_b200.target.DedxTable`, see DESIGN.md.

Every function cites the reference lines it restates (paths relative to
`/root/reference/src/attpc_engine/`).
"""

from __future__ import annotations

import math

import numpy as np
from numba import njit, types
from numba.typed import Dict
from scipy.integrate import solve_ivp

# detector/constants.py:23-35
NUM_TB = 512
E_CHARGE = 1.602176634e-19
C_LIGHT = 299792458.0
MEV_2_JOULE = E_CHARGE * 1.0e6
MEV_2_KG = (E_CHARGE / (C_LIGHT * C_LIGHT)) * 1.0e6

KE_LIMIT = 1e-6  # detector/solver.py:14
TIME_GRID = np.linspace(0, 10e-7, 10001)  # detector/solver.py:16
MESH_N = 10  # detector/transporter.py:8

# detector/beam_pads.py:11-137 (run-length form of the 122 literal ids)
_BEAM_RUNS = (
    (134, 164), (166, 166), (435, 457), (459, 459), (733, 733), (735, 735), (738, 738),
    (740, 741), (5254, 5284), (5286, 5286), (5555, 5577), (5579, 5579), (5853, 5853),
    (5855, 5855), (5858, 5858), (5860, 5861),
)  # fmt: skip
BEAM_PAD_IDS = np.array([p for a, b in _BEAM_RUNS for p in range(a, b + 1)], dtype=np.int64)


# ------------------------------------------------------------------------------ pairing
@njit(cache=False)
def szudzik_pair(tb, pad):
    """detector/pairing.py:6-28."""
    if tb < 0 or pad < 0:
        return -1
    if tb >= pad:
        return tb * tb + tb + pad
    return pad * pad + tb


@njit(cache=False)
def szudzik_unpair(key):
    """detector/pairing.py:31-55 (returns floats, like the reference)."""
    if key < 0:
        return (-1.0, -1.0)
    root = np.floor(np.sqrt(key))
    rest = key - root**2
    if rest < root:
        return (rest, root)
    return (root, rest - root)


# --------------------------------------------------------------------------- trajectory
class TrackProblem:
    """Constants of one track's equation of motion (detector/solver.py:52-66, 285-303)."""

    def __init__(self, nucleus, det_params):
        self.mass = float(nucleus.mass)
        self.charge_number = int(nucleus.Z)
        self.nucleus = nucleus
        self.target = det_params.gas_target
        self.bfield = det_params.bfield * -1.0  # solver.py:298
        self.efield = det_params.efield * -1.0  # solver.py:299
        self.mass_kg = self.mass * MEV_2_KG
        self.q_over_m = self.charge_number * E_CHARGE / self.mass_kg

    def kinetic_energy(self, state):
        gv = math.sqrt(state[3] ** 2.0 + state[4] ** 2.0 + state[5] ** 2.0)
        beta = math.sqrt(gv**2.0 / (1.0 + gv**2.0))
        return self.mass * (gv / beta - 1.0)

    def rhs(self, t, state):
        """detector/solver.py:19-76."""
        gv = math.sqrt(state[3] ** 2.0 + state[4] ** 2.0 + state[5] ** 2.0)
        beta = math.sqrt(gv**2.0 / (1.0 + gv**2.0))
        gamma = gv / beta
        ux, uy, uz = state[3] / gv, state[4] / gv, state[5] / gv
        vx, vy, vz = ux * beta * C_LIGHT, uy * beta * C_LIGHT, uz * beta * C_LIGHT
        ke = self.mass * (gamma - 1.0)
        decel = (
            self.target.get_dedx(self.nucleus, ke) * MEV_2_JOULE * self.target.density * 100.0
        ) / self.mass_kg
        out = np.zeros(6)
        out[0], out[1], out[2] = vx, vy, vz
        out[3] = (self.q_over_m * vy * self.bfield - decel * ux) / C_LIGHT
        out[4] = (self.q_over_m * (-1.0 * vx * self.bfield) - decel * uy) / C_LIGHT
        out[5] = (self.q_over_m * self.efield - decel * uz) / C_LIGHT
        return out


def _terminal_events(problem: TrackProblem):
    """detector/solver.py:80-240 (functions) and :276-283 (terminal flags, directions)."""

    def stopped(t, y):
        return problem.kinetic_energy(y) - KE_LIMIT

    def past_window(t, y):
        return y[2] - 1.0

    def past_micromegas(t, y):
        return y[2]

    def past_cage(t, y):
        return float(np.linalg.norm(y[:2])) - 0.292

    for fn, direction in ((stopped, -1.0), (past_window, 1.0), (past_micromegas, -1.0), (past_cage, 1.0)):
        fn.terminal = True
        fn.direction = direction
    return [stopped, past_window, past_micromegas, past_cage]


def integrate_track(vertex, momentum, nucleus, det_params, rtol=None, atol=None, method="Radau"):
    """detector/solver.py:243-305.  ``rtol``/``atol`` None = scipy defaults, as the reference."""
    problem = TrackProblem(nucleus, det_params)
    y0 = np.zeros(6)
    y0[:3] = vertex
    y0[3:] = np.asarray(momentum)[:3] / nucleus.mass
    extra = {}
    if rtol is not None:
        extra["rtol"] = rtol
    if atol is not None:
        extra["atol"] = atol
    sol = solve_ivp(
        problem.rhs,
        (0.0, 1.0),
        y0,
        method=method,
        events=_terminal_events(problem),
        t_eval=TIME_GRID,
        **extra,
    )
    return sol.y.T


# ---------------------------------------------------------------------------- electrons
def kinetic_energy_of_track(track, mass):
    """detector/solver.py:332-335."""
    gv = np.linalg.norm(track[:, 3:], axis=1)
    beta = np.sqrt(gv**2.0 / (1.0 + gv**2.0))
    gamma = gv / beta
    return mass * (gamma - 1.0)


def mean_electrons(track, mass, w_value):
    """detector/solver.py:338-340: |dKE| per grid step over W, point 0 gets zero."""
    energy = kinetic_energy_of_track(track, mass)
    n = np.zeros_like(energy)
    n[1:] = abs(np.diff(energy))
    n *= 1.0e6 / w_value
    return n


def fano_electrons(track, mass, w_value, fano_factor, rng=None, normals=None):
    """detector/solver.py:343-346.

    With ``rng``: one scalar ``rng.normal(n_k, sqrt(F n_k))`` per point, exactly as the
    reference.  With ``normals`` (replay): ``n_k + sqrt(F n_k) * z_k``, which is what numpy's
    ``Generator.normal`` computes from ``standard_normal`` (pinned by the golden test).
    Cast to int64 truncates toward zero in both cases.
    """
    n = mean_electrons(track, mass, w_value)
    if normals is not None:
        draws = n + np.sqrt(fano_factor * n) * np.asarray(normals, dtype=np.float64)[: len(n)]
        return draws.astype(np.int64)
    return np.array([rng.normal(m, np.sqrt(fano_factor * m)) for m in n], dtype=np.int64)


# ------------------------------------------------------------------------------- drift
@njit(cache=False)
def _grid_index(edges, x_m, y_m):
    """detector/transporter.py:78-120."""
    fx = np.floor(x_m * 1000.0)
    fy = np.floor(y_m * 1000.0)
    if fx >= edges[1] or fy >= edges[1]:
        return -1, -1
    if fx < edges[0] or fy < edges[0]:
        return -1, -1
    return int((fx - edges[0]) / edges[2]), int((fy - edges[0]) / edges[2])


@njit(cache=False)
def _is_beam_pad(pad, beam_ids):
    for b in beam_ids:
        if b == pad:
            return True
    return False


@njit(cache=False)
def _deposit(cloud, key, electrons, label):
    """Insertion-ordered accumulate; zero deposits still create/relabel (transporter.py:166-169, 247-249)."""
    if key in cloud:
        old = cloud[key]
        cloud[key] = (old[0] + electrons, label)
    else:
        cloud[key] = (electrons, label)


@njit(cache=False)
def drift_track(pad_grid, edges, beam_ids, diffusion, efield, dv, xs, ys, times, electrons, cloud, label):
    """detector/transporter.py:252-317 with :123-169 (sigma == 0) and :172-249 (10x10 mesh)."""
    for k in range(len(times)):
        t = times[k]
        cx = xs[k]
        cy = ys[k]
        n_e = electrons[k]
        sigma = np.sqrt(2.0 * diffusion * dv * t / efield)
        if sigma == 0.0:
            ix, iy = _grid_index(edges, cx, cy)
            if ix == -1 or iy == -1:
                continue
            pad = int(pad_grid[ix, iy])
            if pad != -1 and not _is_beam_pad(pad, beam_ids):
                _deposit(cloud, szudzik_pair(int(t), pad), n_e, label)
            continue
        # numba's linspace: start + i*((stop-start)/(n-1)), last element forced to stop
        lo_x = cx - 3 * sigma
        hi_x = cx + 3 * sigma
        lo_y = cy - 3 * sigma
        hi_y = cy + 3 * sigma
        dx = (hi_x - lo_x) / (MESH_N - 1)
        dy = (hi_y - lo_y) / (MESH_N - 1)
        cell_x = 2 * 3 * sigma / (MESH_N - 1)
        cell_y = 2 * 3 * sigma / (MESH_N - 1)
        norm = 1 / 2 / np.pi / (sigma**2)
        for i in range(MESH_N):
            px = lo_x + i * dx
            if i == MESH_N - 1:
                px = hi_x
            for j in range(MESH_N):
                py = lo_y + j * dy
                if j == MESH_N - 1:
                    py = hi_y
                ix, iy = _grid_index(edges, px, py)
                if ix == -1 or iy == -1:
                    continue
                pad = int(pad_grid[ix, iy])
                if pad == -1 or _is_beam_pad(pad, beam_ids):
                    continue
                arg = (-1 / 2 / sigma**2) * (((px - cx) ** 2) + ((py - cy) ** 2))
                share = int((norm * np.exp(arg)) * (cell_x * cell_y) * n_e)
                _deposit(cloud, szudzik_pair(int(t), pad), share, label)


def new_cloud():
    """detector/simulator.py:93-95."""
    return Dict.empty(key_type=types.int64, value_type=types.UniTuple(types.int64, 2))


@njit(cache=False)
def _cloud_arrays(cloud):
    n = len(cloud)
    keys = np.empty(n, dtype=np.int64)
    charge = np.empty(n, dtype=np.int64)
    label = np.empty(n, dtype=np.int64)
    i = 0
    for k, v in cloud.items():
        keys[i] = k
        charge[i] = v[0]
        label[i] = v[1]
        i += 1
    return keys, charge, label


@njit(cache=False)
def _cloud_points(cloud):
    """detector/simulator.py:19-49."""
    n = len(cloud)
    pts = np.empty((n, 3), dtype=np.float64)
    lab = np.empty(n, dtype=np.int64)
    i = 0
    for k, v in cloud.items():
        tb, pad = szudzik_unpair(k)
        pts[i, 0] = pad
        pts[i, 1] = tb
        pts[i, 2] = v[0]
        lab[i] = v[1]
        i += 1
    return pts, lab


def track_to_cloud(track, electrons, config, cloud, label):
    """detector/solver.py:386-413: >=1 mask, gain, z -> time bucket, drift."""
    keep = electrons >= 1
    track = track[keep]
    electrons = electrons[keep] * int(config.det_params.mpgd_gain)
    dv = config.drift_velocity
    times = (config.det_params.length - track[:, 2]) / dv + config.elec_params.micromegas_edge
    drift_track(
        config.pad_grid,
        config.pad_grid_edges,
        BEAM_PAD_IDS,
        config.det_params.diffusion,
        config.det_params.efield,
        dv,
        np.ascontiguousarray(track[:, 0]),
        np.ascontiguousarray(track[:, 1]),
        times,
        electrons,
        cloud,
        label,
    )
    return int(keep.sum()), int((electrons // int(config.det_params.mpgd_gain)).sum())


# ----------------------------------------------------------------------------- simulate
def simulate_event(
    momenta,
    vertex,
    proton_numbers,
    mass_numbers,
    config,
    rng,
    indices,
    nuclear_map,
    record=None,
    tracks=None,
    normals=None,
    uniforms=None,
    solver_kwargs=None,
):
    """detector/simulator.py:52-115.

    ``record`` (a dict) receives intermediates.  ``tracks`` / ``normals`` (lists, one entry per
    simulated charged nucleus, in ``indices`` order) and ``uniforms`` replay pre-computed
    trajectories and random numbers instead of integrating / drawing.
    """
    cloud = new_cloud()
    slot = 0
    stats = {"active_points": 0, "primary_electrons": 0, "trajectory_points": 0}
    rec_tracks, rec_electrons = [], []
    for idx in indices:
        if proton_numbers[idx] == 0:
            continue
        nucleus = nuclear_map.get_data(proton_numbers[idx], mass_numbers[idx])
        if tracks is not None:
            track = np.array(tracks[slot], dtype=np.float64)
        else:
            track = integrate_track(vertex, momenta[idx], nucleus, config.det_params, **(solver_kwargs or {}))
        electrons = fano_electrons(
            track,
            nucleus.mass,
            config.det_params.w_value,
            config.det_params.fano_factor,
            rng=rng,
            normals=None if normals is None else normals[slot],
        )
        rec_tracks.append(track)
        rec_electrons.append(electrons)
        act, prim = track_to_cloud(track, electrons, config, cloud, idx)
        stats["active_points"] += act
        stats["primary_electrons"] += prim
        stats["trajectory_points"] += len(track)
        slot += 1
    points, labels = _cloud_points(cloud)
    if uniforms is None:
        wiggle = rng.uniform(low=0.0, high=1.0, size=len(points))
    else:
        wiggle = np.asarray(uniforms, dtype=np.float64)[: len(points)]
    if record is not None:
        keys, charge, lab = _cloud_arrays(cloud)
        record.update(
            tracks=rec_tracks, electrons=rec_electrons, keys=keys, charges=charge,
            key_labels=lab, uniforms=wiggle, stats=stats,
        )  # fmt: skip
    points[:, 1] += wiggle
    keep = np.logical_and(0 <= points[:, 1], points[:, 1] < NUM_TB)
    return points[keep], labels[keep]


# ------------------------------------------------------------------- electronics response
def get_response(config):
    """detector/response.py:8-32."""
    c1 = 4095 * E_CHARGE / config.elec_params.amp_gain / 1e-15
    tbs = np.linspace(0.0, NUM_TB, NUM_TB)
    c2 = tbs / (config.elec_params.shaping_time * config.elec_params.clock_freq * 0.001)
    resp = c1 * np.exp(-3.0 * c2) * (c2**3) * np.sin(c2)
    resp[resp < 0] = 0
    return resp


@njit(cache=False)
def shaped_amplitude(response, electrons):
    """detector/response.py:35-57: scale, clip at 4095, (max, sum)."""
    sig = response * electrons
    for i in range(len(sig)):
        if sig[i] > 4095:
            sig[i] = 4095
    return sig.max(), sig.sum()


@njit(cache=False)
def spyral_rows(points, window_edge, mm_edge, length, response, pad_centers, pad_sizes):
    """detector/writer.py:61-112."""
    out = np.empty((len(points), 8))
    for i in range(len(points)):
        pad = int(points[i, 0])
        amp, integral = shaped_amplitude(response, points[i, 2])
        out[i, 0] = pad_centers[pad, 0]
        out[i, 1] = pad_centers[pad, 1]
        out[i, 2] = (window_edge - points[i, 1]) / (window_edge - mm_edge) * length * 1000.0
        out[i, 3] = amp
        out[i, 4] = integral
        out[i, 5] = points[i, 0]
        out[i, 6] = points[i, 1]
        out[i, 7] = pad_sizes[pad]
    return out


def spyral_event(points, labels, config, response):
    """detector/writer.py:220-238: rows, ADC threshold, z-sort."""
    rows = spyral_rows(
        points,
        config.elec_params.windows_edge,
        config.elec_params.micromegas_edge,
        config.det_params.length,
        response,
        config.pad_centers,
        config.pad_sizes,
    )
    keep = rows[:, 3] > config.elec_params.adc_threshold
    rows, labels = rows[keep], labels[keep]
    order = np.argsort(rows[:, 2])
    return rows[order], labels[order]
