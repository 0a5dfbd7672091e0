"""Replay fixtures drawn from the BENCHMARK's own kinematic distributions ("next round" item 1a).

    python tests/golden/make_workload_golden.py     # writes tests/golden/workload_<name>.npz for the four workloads

For each workload of bench.py, the first N events of ``bench.build_workload(name, N)`` go through the UNMODIFIED
reference (tests/golden/ref_shim.py) exactly as in tests/golden/make_golden.py: once through
`attpc_engine.detector.simulator.simulate` end to end, once staged (trajectory, normals, electrons, the
insertion-ordered dict, uniforms); the staged result must equal the end-to-end one.  Stored per workload:

* the kinematics fed in, the trajectories (inert tails trimmed), the standard normals and the electrons per row;
* the dict in insertion order (Szudzik keys, charges, labels) and the uniforms the reference drew;
* SHA-256 digests of the reference's final cloud / labels and of its `SpyralWriter.write` rows / labels.  The full
  arrays follow from the stored dict by the three lines of `simulator.py:104-113` and by `writer.py:61-112,232-238`;
  the CPU test `tests/test_oracle_golden.py::test_workload_fixture_digests` rebuilds them with the oracle and checks
  the digests, and the GPU test compares the CUDA path with the rebuilt arrays.  (Storing the arrays themselves
  would be 25 MB of incompressible float64.)

Needs /root/reference; the fixtures it writes are committed.
"""

import hashlib
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent.parent))

import make_golden as mg  # noqa: E402  (installs the shim, imports the reference)
from ref_workloads import reference_config  # noqa: E402

import bench  # noqa: E402

N_EVENTS = 32
WORKLOADS = ("c16dd", "c14dp", "c12aa", "sn132dp")


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode() + str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()


def make(name):
    _, momenta, vertices, zs, as_, indices = bench.build_workload(name, N_EVENTS)
    cfg = reference_config(name)
    zs, as_ = np.asarray(zs), np.asarray(as_)
    out = dict(momenta=momenta, vertices=vertices, Z=zs, A=as_, indices=np.array(indices))
    rows, normals, electrons = [], [], []
    t_event, t_rank, t_idx, t_full = [], [], [], []
    keys, charges, key_labels, uniforms = [], [], [], []
    key_off = [0]
    digests = []
    for e in range(N_EVENTS):
        seed = 900000 + 1000 * bench.WORKLOADS[name]["config_id"] + e
        cloud_ref, labels_ref = mg.ref_sim.simulate(
            momenta[e].copy(), vertices[e], zs, as_, cfg, np.random.default_rng(seed), indices
        )
        st = mg.staged_reference_event(cfg, momenta[e], vertices[e], zs, as_, indices, seed)
        assert np.array_equal(st["cloud"], cloud_ref) and np.array_equal(st["labels"], labels_ref), (name, e)
        charged = [(rank, idx) for rank, idx in enumerate(indices) if zs[idx] != 0]
        for (rank, idx), tr, zn, el, full in zip(charged, st["tracks"], st["normals"], st["electrons"], st["full_len"]):
            rows.append(tr)
            normals.append(zn)
            electrons.append(el)
            t_event.append(e)
            t_rank.append(rank)
            t_idx.append(idx)
            t_full.append(full)
        keys.append(st["keys"])
        charges.append(st["charges"])
        key_labels.append(st["key_labels"])
        uniforms.append(st["uniforms"])
        key_off.append(key_off[-1] + len(st["keys"]))
        if len(cloud_ref):
            srows, slabels = mg.spyral_through_reference_writer(cfg, cloud_ref, labels_ref, e)
        else:
            srows, slabels = np.zeros((0, 8)), np.zeros(0, np.int64)
        digests.append(dict(n_cloud=int(len(cloud_ref)), cloud=digest(cloud_ref, labels_ref), n_spyral=int(len(srows)),
                            spyral=digest(srows, slabels)))  # fmt: skip
        print(name, e, "tracks", [len(t) for t in st["tracks"]], "keys", len(st["keys"]), "cloud", len(cloud_ref),
              "spyral", len(srows), flush=True)  # fmt: skip
    lens = np.array([len(r) for r in rows], dtype=np.int64)
    out.update(
        track_offsets=np.concatenate([[0], np.cumsum(lens)]), track_rows=np.concatenate(rows),
        track_normals=np.concatenate(normals), track_electrons=np.concatenate(electrons),
        track_event=np.array(t_event, dtype=np.int32), track_rank=np.array(t_rank, dtype=np.int32),
        track_idx=np.array(t_idx, dtype=np.int32), track_full_len=np.array(t_full, dtype=np.int64),
        key_offsets=np.array(key_off, dtype=np.int64), keys=np.concatenate(keys), charges=np.concatenate(charges),
        key_labels=np.concatenate(key_labels), uniforms=np.concatenate(uniforms),
    )  # fmt: skip
    np.savez_compressed(HERE / f"workload_{name}.npz", **out)
    return digests


def main():
    meta = {}
    for name in WORKLOADS if len(sys.argv) < 2 else sys.argv[1:]:
        meta[name] = make(name)
        print(name, (HERE / f"workload_{name}.npz").stat().st_size, "bytes", flush=True)
    path = HERE / "workload_digests.json"
    old = json.loads(path.read_text()) if path.exists() else {}
    old.update(meta)
    path.write_text(json.dumps(old, indent=1))


if __name__ == "__main__":
    main()
