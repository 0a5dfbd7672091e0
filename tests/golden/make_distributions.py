"""Distribution samples of the UNMODIFIED reference for the KS tests of parity part (b).

    python tests/golden/make_distributions.py [name ...]     # writes tests/golden/distributions.npz

Runs `attpc_engine.detector.simulator.simulate` (via tests/golden/ref_shim.py) on synthetic kinematics of all four
workloads of bench.py, in parallel over the host cores (every worker owns a disjoint event range and its own
generator), and stores, per event, summary observables of the returned cloud and, per simulated track, the
number of trajectory rows and the path length of `generate_trajectory` (recorded by wrapping the reference's
function, not by changing it).  The first PAIRED events also keep their kinematics so that a test can feed the CUDA
path the very same events; the KS tests proper use kinematics drawn with a DIFFERENT seed (unpaired).
"""

import multiprocessing as mp
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent.parent))

N_EVENTS = {"c16dd": 6000, "c14dp": 4000, "c12aa": 2400, "sn132dp": 2400}
PAIRED = {"c16dd": 400, "c14dp": 300, "c12aa": 150, "sn132dp": 300}  # events whose kinematics are stored
_WORKLOADS = {}
STRIDE = 4  # the path length is the chord sum over every 4th trajectory row (the CUDA side uses the same rows)


def event_observables(cloud, labels, indices, picker):
    """Per-event observables (independent across events, which is what a KS test needs)."""
    per_label = [int((labels == i).sum()) for i in indices]
    if len(cloud) == 0:
        return (0, 0.0, 0.0, 0, 0.0, 0.0, np.nan, np.nan), per_label
    q = cloud[:, 2]
    pads = cloud[:, 0].astype(np.int64)
    one_pad = pads[picker.integers(len(pads))]  # a random hit pad: its charge summed over time = "charge per pad"
    return (len(cloud), float(q.sum()), float(cloud[:, 1].max() - cloud[:, 1].min()), len(np.unique(pads)),
            float(np.median(q)), float(q.max()), float(q[picker.integers(len(q))]), float(q[pads == one_pad].sum())), per_label  # fmt: skip


def path_length(track, stride=STRIDE):
    pts = track[::stride, :3]
    return float(np.sqrt((np.diff(pts, axis=0) ** 2).sum(axis=1)).sum()) if len(pts) > 1 else 0.0


def _worker(args):
    name, first, count = args
    from ref_workloads import reference_config

    import bench
    from attpc_engine.detector import solver as ref_solver
    from attpc_engine.detector.simulator import simulate

    cfg = reference_config(name)
    if name not in _WORKLOADS:  # once per worker process
        _WORKLOADS[name] = bench.build_workload(name, N_EVENTS[name])
    _, momenta, vertices, zs, as_, indices = _WORKLOADS[name]
    rng = np.random.default_rng([424242, first])
    picker = np.random.default_rng([1, first])
    traj = []
    original = ref_solver.generate_trajectory

    def recording(*a, **k):
        track = original(*a, **k)
        traj.append((len(track), path_length(track)))
        return track

    ref_solver.generate_trajectory = recording
    obs, per_label, per_track = [], [], []
    try:
        for i in range(first, first + count):
            traj.clear()
            cloud, labels = simulate(momenta[i].copy(), vertices[i], zs, as_, cfg, rng, indices)
            o, pl = event_observables(cloud, labels, indices, picker)
            obs.append(o)
            per_label.append(pl)
            per_track.append(list(traj))
    finally:
        ref_solver.generate_trajectory = original
    return first, np.array(obs, dtype=np.float64), np.array(per_label, dtype=np.int64), np.array(per_track, dtype=np.float64)


def main():
    import bench

    path = HERE / "distributions.npz"
    out = {}
    if path.exists() and len(sys.argv) > 1:  # partial regeneration keeps the other workloads
        with np.load(path) as f:
            out = {k: f[k] for k in f.files}
    names = sys.argv[1:] or list(N_EVENTS)
    cores = len(__import__("os").sched_getaffinity(0))
    with mp.get_context("spawn").Pool(cores) as pool:
        for name in names:
            n = N_EVENTS[name]
            per = 25
            jobs = [(name, a, min(per, n - a)) for a in range(0, n, per)]
            parts = sorted(pool.imap_unordered(_worker, jobs), key=lambda p: p[0])
            obs = np.concatenate([p[1] for p in parts])
            per_label = np.concatenate([p[2] for p in parts])
            per_track = np.concatenate([p[3] for p in parts])
            _, momenta, vertices, zs, as_, indices = bench.build_workload(name, n)
            for k in [k for k in out if k.startswith(name + "/")]:
                del out[k]
            keep = PAIRED[name]
            out[f"{name}/momenta"] = momenta[:keep]
            out[f"{name}/vertices"] = vertices[:keep]
            out[f"{name}/Z"] = np.asarray(zs)
            out[f"{name}/A"] = np.asarray(as_)
            out[f"{name}/indices"] = np.array(indices)
            cols = ("n_points", "sum_charge", "tb_extent", "n_pads", "median_charge", "max_charge", "point_charge_sample",
                    "pad_charge_sample")  # fmt: skip
            for c, col in enumerate(cols):
                v = obs[:, c]
                out[f"{name}/{col}"] = v[~np.isnan(v)] if col.endswith("_sample") else v
            out[f"{name}/points_per_track"] = per_label  # [n, len(indices)] cloud rows per label, in `indices` order
            out[f"{name}/traj_rows"] = per_track[:, :, 0]  # [n, charged tracks] len(generate_trajectory(...))
            out[f"{name}/path_length"] = per_track[:, :, 1]  # [n, charged tracks] metres, chords over every 4th row
            print(name, n, "events; mean points", obs[:, 0].mean(), "mean charge", obs[:, 1].mean(), "mean rows",
                  per_track[:, :, 0].mean(axis=0), "mean path", per_track[:, :, 1].mean(axis=0), flush=True)  # fmt: skip
            np.savez_compressed(path, **out)
    print(path.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
