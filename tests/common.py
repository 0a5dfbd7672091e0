"""Shared helpers of the test-suite: configurations of the golden cases, fixture loading."""

import json
from pathlib import Path

import numpy as np

from attpc_engine_b200 import nuclear_map
from attpc_engine_b200.detector import Config, DetectorParams, ElectronicsParams, PadParams
from attpc_engine_b200.detector.pairing import pair
from attpc_engine_b200.target import AnalyticGasTarget, TableGasTarget

GOLDEN = Path(__file__).resolve().parent / "golden"
META = json.loads((GOLDEN / "golden_meta.json").read_text())
GASES = {k: (v["compound"], v["pressure"]) for k, v in META["gases"].items()}
_gas_cache: dict = {}


def load_golden(name):
    with np.load(GOLDEN / name) as f:
        return {k: f[k] for k in f.files}


def gas(name="D2_600"):
    """Same dE/dx tables the fixtures were generated with (tests/golden/ref_shim.py: GasTarget)."""
    if name not in _gas_cache:
        compound, pressure = GASES[name]
        _gas_cache[name] = TableGasTarget(AnalyticGasTarget([tuple(c) for c in compound], pressure))
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


def case_names():
    return list(META["cases"].keys())


def case_config(name):
    return make_config(**META["cases"][name]["config"])


def case_tracks(ev, name):
    """Charged tracks of a golden case in `indices` order: list of (nucleus idx, rank, (Z, A), rows, normals, electrons)."""
    zs, as_, indices = ev[f"{name}/Z"], ev[f"{name}/A"], ev[f"{name}/indices"]
    out, slot = [], 0
    for rank, idx in enumerate(indices):
        if zs[idx] == 0:
            continue
        out.append(
            dict(idx=int(idx), rank=rank, za=(int(zs[idx]), int(as_[idx])), rows=ev[f"{name}/track{slot}"],
                 normals=ev[f"{name}/normals{slot}"], electrons=ev[f"{name}/electrons{slot}"])
        )  # fmt: skip
        slot += 1
    return out


def nuclei_of(tracks):
    return [nuclear_map.get_data(z, a) for z, a in dict.fromkeys(t["za"] for t in tracks)]


def cloud_keys(cloud):
    """Szudzik key of every cloud row ([pad, tb(float), e])."""
    return np.asarray(pair(np.floor(cloud[:, 1]).astype(np.int64), cloud[:, 0].astype(np.int64)), dtype=np.int64)


def canonical_order(cloud):
    """Row order of the CUDA path: ascending (time bucket, pad)."""
    return np.lexsort((cloud[:, 0].astype(np.int64), np.floor(cloud[:, 1]).astype(np.int64)))


def sort_cloud(cloud, labels):
    order = canonical_order(cloud)
    return cloud[order], labels[order]


# ------------------------------------------------------------------ fixtures drawn from the bench workloads
WORKLOAD_NAMES = ("c16dd", "c14dp", "c12aa", "sn132dp")
WORKLOAD_GAS = {"c16dd": "D2_600", "c14dp": "D2_600", "c12aa": "He4_600", "sn132dp": "D2_600"}


def digest(*arrays):
    """SHA-256 over dtype, shape and bytes (tests/golden/make_workload_golden.py: digest)."""
    import hashlib

    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode() + str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()


def workload_config(name):
    """The Config of bench.build_workload(name) (same numbers for all four; only the gas differs)."""
    return make_config(WORKLOAD_GAS[name])


def load_workload(name):
    """Fixture of tests/golden/make_workload_golden.py as a dict + per-event digests of the reference's outputs."""
    fx = load_golden(f"workload_{name}.npz")
    fx["digests"] = json.loads((GOLDEN / "workload_digests.json").read_text())[name]
    return fx


def workload_tracks(fx):
    """Per charged track: dict(event, rank, idx, za, rows, normals, electrons), in the order the reference ran them."""
    off = fx["track_offsets"]
    out = []
    for t in range(len(off) - 1):
        idx = int(fx["track_idx"][t])
        a, b = off[t], off[t + 1]
        out.append(dict(event=int(fx["track_event"][t]), rank=int(fx["track_rank"][t]), idx=idx,
                        za=(int(fx["Z"][idx]), int(fx["A"][idx])), rows=fx["track_rows"][a:b],
                        normals=fx["track_normals"][a:b], electrons=fx["track_electrons"][a:b]))  # fmt: skip
    return out


def workload_event_dict(fx, e):
    """(keys, charges, labels, uniforms) of event e in the reference's dict insertion order."""
    a, b = fx["key_offsets"][e], fx["key_offsets"][e + 1]
    return fx["keys"][a:b], fx["charges"][a:b], fx["key_labels"][a:b], fx["uniforms"][a:b]


def reference_cloud_from_dict(keys, charges, labels, uniforms, num_tb=512):
    """`dict_to_points` + wiggle + time-bucket mask (detector/simulator.py:19-49, 104-113) on a recorded dict."""
    from attpc_engine_b200.detector.pairing import unpair

    tb, pad = unpair(keys)
    cloud = np.empty((len(keys), 3), dtype=np.float64)
    cloud[:, 0] = pad
    cloud[:, 1] = tb
    cloud[:, 2] = charges
    cloud[:, 1] += uniforms
    keep = np.logical_and(0 <= cloud[:, 1], cloud[:, 1] < num_tb)
    return cloud[keep], np.asarray(labels, dtype=np.int64)[keep]
