"""Generate golden fixtures by running the UNMODIFIED reference in the build container.

    python tests/golden/make_golden.py        # writes tests/golden/*.npz + golden_meta.json

Needs `/root/reference` (read-only mount) -- it therefore runs only where the reference is
present; the fixtures it writes are committed and are what the tests read everywhere else.
For every case the reference's own functions are driven twice from the same seed:

1. `attpc_engine.detector.simulator.simulate(...)` end to end  -> `cloud`, `labels`;
2. staged, calling the reference's `generate_trajectory`, `generate_electrons`,
   `transport_track`, `dict_to_points` in the order `simulate` does, to capture trajectories,
   electrons, the insertion-ordered dict and (from a shadow generator on the same seed) the
   standard normals / uniforms the reference consumed.  The staged result must equal (1).

Then the reference's `SpyralWriter.write` runs on the cloud through an in-memory h5py
stand-in to capture the 8-column rows.
"""

import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent.parent))

import ref_shim  # noqa: E402

ae = ref_shim.install()

from attpc_engine.detector import Config, DetectorParams, ElectronicsParams, PadParams  # noqa: E402
from attpc_engine.detector import simulator as ref_sim  # noqa: E402
from attpc_engine.detector import solver as ref_solver  # noqa: E402
from attpc_engine.detector.pairing import pair as ref_pair  # noqa: E402
from attpc_engine.detector.pairing import unpair as ref_unpair  # noqa: E402
from attpc_engine.detector.response import get_response as ref_get_response  # noqa: E402
from attpc_engine.detector.transporter import position_to_index as ref_position_to_index  # noqa: E402
from attpc_engine.detector.transporter import transport_track as ref_transport_track  # noqa: E402
from attpc_engine.detector.beam_pads import BEAM_PADS_ARRAY  # noqa: E402
from attpc_engine.detector.writer import SpyralWriter  # noqa: E402
from numba.core import types  # noqa: E402
from numba.typed import Dict  # noqa: E402
from spyral_utils.nuclear.target import GasTarget  # noqa: E402

nuclear_map = ae.nuclear_map

GASES = {"D2_600": ([(1, 2, 2)], 600.0), "He4_600": ([(2, 4, 1)], 600.0), "H2_600": ([(1, 1, 2)], 600.0)}
_gas_cache = {}


def gas(name):
    if name not in _gas_cache:
        comp, p = GASES[name]
        _gas_cache[name] = GasTarget(comp, p, nuclear_map)
    return _gas_cache[name]


def make_config(gas_name="D2_600", bfield=3.0, diffusion=0.277, threshold=40):
    det = DetectorParams(
        length=1.0, efield=45000.0, bfield=bfield, mpgd_gain=175000, gas_target=gas(gas_name),
        diffusion=diffusion, fano_factor=0.2, w_value=34.0,
    )  # fmt: skip
    elec = ElectronicsParams(
        clock_freq=6.25, amp_gain=900, shaping_time=1000, micromegas_edge=10, windows_edge=560,
        adc_threshold=threshold,
    )  # fmt: skip
    return Config(det, elec, PadParams())


def p4(z, a, ke, theta, phi):
    m = nuclear_map.get_data(z, a).mass
    e = ke + m
    p = np.sqrt(e * e - m * m)
    return [p * np.sin(theta) * np.cos(phi), p * np.sin(theta) * np.sin(phi), p * np.cos(theta), e]


# name -> (config kwargs, Z, A, [(ke, theta, phi) per nucleus], vertex, indices, seed)
CASES = {
    "dd_exit": (
        {}, [1, 6, 1, 6], [2, 16, 2, 16],
        [(0, 0, 0), (180, 0, 0), (2.0, 1.1, 0.3), (170.0, 0.04, 3.44)], [0.002, -0.001, 0.55], [2, 3], 11,
    ),
    "dd_stop": (
        {}, [1, 6, 1, 6], [2, 16, 2, 16],
        [(0, 0, 0), (180, 0, 0), (0.6, 1.3, 2.0), (178.0, 0.02, 5.14)], [-0.003, 0.004, 0.30], [2, 3], 12,
    ),
    "dd_back": (
        {}, [1, 6, 1, 6], [2, 16, 2, 16],
        [(0, 0, 0), (180, 0, 0), (1.5, 2.4, 4.0), (176.0, 0.03, 0.86)], [0.001, 0.002, 0.15], [2, 3], 13,
    ),
    "dp_decay": (
        {}, [1, 6, 1, 6, 6, 0], [2, 14, 1, 15, 14, 1],
        [(0, 0, 0), (150, 0, 0), (6.0, 0.9, 1.0), (140.0, 0.05, 4.1), (130.0, 0.06, 4.0), (5.0, 0.3, 1.0)],
        [0.0, 0.003, 0.60], [2, 4, 5], 14,
    ),
    "alpha_breakup": (
        {"gas_name": "He4_600"}, [2, 6, 2, 6, 2, 4, 2, 2], [4, 12, 4, 12, 4, 8, 4, 4],
        [(0, 0, 0), (60, 0, 0), (8.0, 0.7, 0.5), (45.0, 0.2, 3.6), (12.0, 0.35, 2.0), (30.0, 0.1, 5.0),
         (14.0, 0.25, 5.3), (15.0, 0.15, 4.2)],
        [0.002, 0.002, 0.70], [2, 4, 6, 7], 15,
    ),
    "outside": (
        {"bfield": 2.85}, [1, 1, 1, 1], [1, 1, 1, 1],
        [(0.0533, 0, 0)] * 4, [1.0, 1.0, 1.0], [0], 16,
    ),
    "nodiff": (
        {"diffusion": 0.0}, [1, 6, 1, 6], [2, 16, 2, 16],
        [(0, 0, 0), (180, 0, 0), (3.0, 0.8, 1.0), (172.0, 0.05, 4.14)], [0.0, 0.0, 0.75], [2, 3], 17,
    ),
    "sn_dp": (
        {"threshold": 10}, [1, 50, 1, 50], [2, 132, 1, 133],
        [(0, 0, 0), (1300, 0, 0), (4.0, 2.2, 0.7), (1280.0, 0.012, 3.84)], [0.004, 0.0, 0.80], [2, 3], 18,
    ),
}  # fmt: skip


def staged_reference_event(cfg, momenta, vertex, zs, as_, indices, seed):
    """Drive the reference's stage functions in `simulate`'s order, recording intermediates."""
    rng = np.random.default_rng(seed)
    shadow = np.random.default_rng(seed)
    points = Dict.empty(key_type=types.int64, value_type=types.Tuple(types=[types.int64, types.int64]))
    tracks, normals, electrons_all, full_len = [], [], [], []
    for idx in indices:
        if zs[idx] == 0:
            continue
        nucleus = nuclear_map.get_data(zs[idx], as_[idx])
        track = ref_solver.generate_trajectory(vertex, momenta[idx], nucleus, cfg.det_params)
        z = shadow.standard_normal(len(track))
        electrons = ref_solver.generate_electrons(track, nucleus, cfg.det_params, rng)
        # trim the inert tail of stopped tracks (all < 1 electron) to keep fixtures small
        live = np.nonzero(electrons >= 1)[0]
        keep = min(len(track), (int(live[-1]) + 17) if len(live) else 8)
        tracks.append(track[:keep].copy())
        normals.append(z[:keep].copy())
        electrons_all.append(electrons[:keep].copy())
        full_len.append(len(track))
        mask = electrons >= 1
        tr, el = track[mask], electrons[mask]
        el *= cfg.det_params.mpgd_gain
        dv = cfg.drift_velocity
        tr[:, 2] = (cfg.det_params.length - tr[:, 2]) / dv + cfg.elec_params.micromegas_edge
        ref_transport_track(
            cfg.pad_grid, cfg.pad_grid_edges, cfg.det_params.diffusion, cfg.det_params.efield, dv, tr, el,
            points, idx,
        )  # fmt: skip
    keys = np.array(list(points.keys()), dtype=np.int64)
    vals = np.array(list(points.values()), dtype=np.int64).reshape(-1, 2)
    pts, labs = ref_sim.dict_to_points(points)
    u_ref = rng.uniform(low=0.0, high=1.0, size=len(pts))
    u = shadow.random(len(pts))
    assert np.array_equal(u, u_ref), "uniform replay contract broken"
    pts[:, 1] += u_ref
    m = np.logical_and(0 <= pts[:, 1], pts[:, 1] < 512)
    return dict(
        tracks=tracks, normals=normals, electrons=electrons_all, full_len=np.array(full_len),
        keys=keys, charges=vals[:, 0] if len(vals) else np.zeros(0, np.int64),
        key_labels=vals[:, 1] if len(vals) else np.zeros(0, np.int64), uniforms=u,
        cloud=pts[m], labels=labs[m],
    )  # fmt: skip


def spyral_through_reference_writer(cfg, cloud, labels, event_number):
    ref_shim.MemFile.opened.clear()
    writer = SpyralWriter(Path("/nonexistent"), cfg, 5000)
    writer.write(cloud.copy(), labels.copy(), cfg, event_number)
    writer.close()
    grp = ref_shim.MemFile.opened[0]["cloud"]
    return grp[f"cloud_{event_number}"].data, grp[f"labels_{event_number}"].data


def make_events(out):
    meta = {}
    for name, (ckw, zs, as_, kin, vertex, indices, seed) in CASES.items():
        cfg = make_config(**ckw)
        zs, as_ = np.array(zs), np.array(as_)
        momenta = np.array([p4(z, a, *k) for z, a, k in zip(zs, as_, kin)])
        vertex = np.array(vertex, dtype=np.float64)
        cloud_ref, labels_ref = ref_sim.simulate(
            momenta.copy(), vertex, zs, as_, cfg, np.random.default_rng(seed), indices
        )
        st = staged_reference_event(cfg, momenta, vertex, zs, as_, indices, seed)
        assert np.array_equal(st["cloud"], cloud_ref) and np.array_equal(st["labels"], labels_ref), name
        # the Fano replay contract: electrons == int64(n + sqrt(F n) z)
        for tr, zn, el, idx in zip(st["tracks"], st["normals"], st["electrons"], [i for i in indices if zs[i] != 0]):
            mass = nuclear_map.get_data(zs[idx], as_[idx]).mass
            gv = np.linalg.norm(tr[:, 3:], axis=1)
            ke = mass * (gv / np.sqrt(gv**2.0 / (1.0 + gv**2.0)) - 1.0)
            n = np.zeros_like(ke)
            n[1:] = abs(np.diff(ke))
            n *= 1.0e6 / cfg.det_params.w_value
            assert np.array_equal((n + np.sqrt(cfg.det_params.fano_factor * n) * zn).astype(np.int64), el), name
        if len(cloud_ref):
            rows, row_labels = spyral_through_reference_writer(cfg, cloud_ref, labels_ref, 7)
        else:
            rows, row_labels = np.zeros((0, 8)), np.zeros(0, np.int64)
        out[f"{name}/momenta"] = momenta
        out[f"{name}/vertex"] = vertex
        out[f"{name}/Z"] = zs
        out[f"{name}/A"] = as_
        out[f"{name}/indices"] = np.array(indices)
        for t, (tr, zn, el) in enumerate(zip(st["tracks"], st["normals"], st["electrons"])):
            out[f"{name}/track{t}"] = tr
            out[f"{name}/normals{t}"] = zn
            out[f"{name}/electrons{t}"] = el
        for k in ("full_len", "keys", "charges", "key_labels", "uniforms", "cloud", "labels"):
            out[f"{name}/{k}"] = st[k]
        out[f"{name}/spyral_rows"] = rows
        out[f"{name}/spyral_labels"] = row_labels
        meta[name] = dict(
            config=dict(ckw), seed=seed, n_tracks=len(st["tracks"]), n_keys=int(len(st["keys"])),
            n_cloud=int(len(cloud_ref)), n_spyral=int(len(rows)),
            track_points=[int(x) for x in st["full_len"]],
        )  # fmt: skip
        print(name, meta[name])
    return meta


def make_pad_lookup(out):
    """G1: the reference's position -> pad chain (position_to_index, grid read, beam veto)."""
    cfg = make_config()
    rng = np.random.default_rng(20260101)
    n = 60000
    xy = rng.uniform(-0.2805, 0.2805, size=(n, 2))
    edge = np.array([-0.2800, -0.28000001, -0.2799999, 0.2789999, 0.279, 0.27900001, 0.2795, 0.28, 0.0, -0.0,
                     1e-9, -1e-9, 0.001, -0.001, 0.0009999999, 0.0195, -0.0195])  # fmt: skip
    ex, ey = np.meshgrid(edge, edge)
    xy = np.concatenate([xy, np.stack([ex.ravel(), ey.ravel()], axis=1)])
    # positions over the beam region, where the veto matters
    xy = np.concatenate([xy, rng.uniform(-0.025, 0.025, size=(8000, 2))])
    pads = np.empty(len(xy), dtype=np.int16)
    for i, (x, y) in enumerate(xy):
        ix, iy = ref_position_to_index(cfg.pad_grid_edges, (x, y))
        if ix == -1 or iy == -1:
            pads[i] = -1
            continue
        pad = int(cfg.pad_grid[ix, iy])
        pads[i] = -1 if (pad == -1 or pad in BEAM_PADS_ARRAY) else pad
    out["pad_lookup/xy"] = xy
    out["pad_lookup/pad"] = pads
    print("pad_lookup", len(xy), "vetoed", int((pads < 0).sum()))


def make_misc(out):
    cfg = make_config()
    out["response/default"] = ref_get_response(cfg)
    tb = np.array([56, 937, 0, 511, 10, 560, 10239, 0, 3, 3], dtype=np.int64)
    pad = np.array([937, 56, 0, 10239, 10, 560, 10239, 1, 3, 4], dtype=np.int64)
    out["pairing/tb"] = tb
    out["pairing/pad"] = pad
    out["pairing/key"] = np.array([ref_pair(int(a), int(b)) for a, b in zip(tb, pad)], dtype=np.int64)
    out["pairing/unpaired"] = np.array([ref_unpair(int(k)) for k in out["pairing/key"]], dtype=np.float64)


def make_tight_trajectories(out):
    """G6: the reference's RHS/events integrated by scipy at tight tolerance (every 8th grid point)."""
    cfg = make_config()
    from scipy.integrate import solve_ivp

    kin = {
        "d_exit": ((1, 2), 2.0, 1.1, 0.3, [0.002, -0.001, 0.55]),
        "d_stop": ((1, 2), 0.6, 1.3, 2.0, [-0.003, 0.004, 0.30]),
        "d_loop": ((1, 2), 1.2, 1.5, 0.0, [0.0, 0.0, 0.50]),
        "c16_fwd": ((6, 16), 170.0, 0.04, 3.44, [0.002, -0.001, 0.55]),
        "p_back": ((1, 1), 4.0, 2.2, 0.7, [0.004, 0.0, 0.80]),
        "alpha": ((2, 4), 8.0, 0.7, 0.5, [0.002, 0.002, 0.70]),
    }
    orig = solve_ivp

    def tight(fun, t_span, *a, **k):
        # same RHS / events / t_eval as the reference's own call.  Only the solver settings change:
        # tolerances tightened to converge the solution, the span cut at the last grid point (nothing
        # beyond 1 us is ever emitted, solver.py:16), and an explicit high-order method, because
        # Radau's finite-difference Jacobian on the piecewise-linear dE/dx table makes it crawl at
        # 1e-11 (it did not finish one stopping track in 10 minutes).
        k.update(rtol=1e-11, atol=1e-14, method="DOP853")
        return orig(fun, (0.0, float(ref_solver.TIME_STEPS[-1])), *a, **k)

    for name, ((z, a), ke, th, ph, vtx) in kin.items():
        nucleus = nuclear_map.get_data(z, a)
        mom = np.array(p4(z, a, ke, th, ph))
        default = ref_solver.generate_trajectory(np.array(vtx), mom, nucleus, cfg.det_params)
        ref_solver.solve_ivp = tight
        try:
            fine = ref_solver.generate_trajectory(np.array(vtx), mom, nucleus, cfg.det_params)
        finally:
            ref_solver.solve_ivp = orig
        out[f"traj/{name}/za"] = np.array([z, a])
        out[f"traj/{name}/momentum"] = mom
        out[f"traj/{name}/vertex"] = np.array(vtx)
        out[f"traj/{name}/npoints"] = np.array([len(default), len(fine)])
        out[f"traj/{name}/tight_every8"] = fine[: min(len(fine), 4001) : 8].copy()
        out[f"traj/{name}/default_every8"] = default[: min(len(default), 4001) : 8].copy()
        print("traj", name, len(default), len(fine))


def main():
    import numba
    import scipy

    events, misc, traj = {}, {}, {}
    meta = {"cases": make_events(events)}
    np.savez_compressed(HERE / "events.npz", **events)
    make_pad_lookup(misc)
    make_misc(misc)
    np.savez_compressed(HERE / "misc.npz", **misc)
    make_tight_trajectories(traj)
    np.savez_compressed(HERE / "trajectories.npz", **traj)
    meta["versions"] = dict(numpy=np.__version__, scipy=scipy.__version__, numba=numba.__version__)
    meta["gases"] = {k: dict(compound=v[0], pressure=v[1]) for k, v in GASES.items()}
    meta["reference"] = "ATTPC/attpc_engine 0.9.0, unmodified, via tests/golden/ref_shim.py"
    (HERE / "golden_meta.json").write_text(json.dumps(meta, indent=1))
    for f in ("events.npz", "misc.npz", "trajectories.npz"):
        print(f, (HERE / f).stat().st_size)


if __name__ == "__main__":
    main()
