"""Distribution samples of the UNMODIFIED reference for the KS tests of parity part (b).

    python tests/golden/make_distributions.py     # writes tests/golden/distributions.npz

Runs `attpc_engine.detector.simulator.simulate` (via tests/golden/ref_shim.py) on synthetic kinematics of two
workloads of bench.py and stores, per event, the kinematics fed in and summary observables of the returned cloud.
"""

import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent.parent))

import ref_shim  # noqa: E402

ae = ref_shim.install()

from attpc_engine.detector import Config, DetectorParams, ElectronicsParams, PadParams  # noqa: E402
from attpc_engine.detector.simulator import simulate  # noqa: E402
from spyral_utils.nuclear.target import GasTarget  # noqa: E402

import bench  # noqa: E402

N_EVENTS = {"c16dd": 400, "c12aa": 150}


def event_observables(cloud):
    """Per-event observables (independent across events, which is what a KS test needs)."""
    if len(cloud) == 0:
        return 0, 0.0, 0.0, 0, 0.0, 0.0
    q = cloud[:, 2]
    return (len(cloud), float(q.sum()), float(cloud[:, 1].max() - cloud[:, 1].min()), len(np.unique(cloud[:, 0])),
            float(np.median(q)), float(q.max()))


def main():
    out = {}
    for name, n in N_EVENTS.items():
        config_b200, momenta, vertices, zs, as_, indices = bench.build_workload(name, n)
        compound, pressure = bench.WORKLOADS[name]["gas"]
        gas = GasTarget(compound, pressure, ae.nuclear_map)
        d, e = config_b200.det_params, config_b200.elec_params
        det = DetectorParams(d.length, d.efield, d.bfield, d.mpgd_gain, gas, d.diffusion, d.fano_factor, d.w_value)
        elec = ElectronicsParams(e.clock_freq, e.amp_gain, e.shaping_time, e.micromegas_edge, e.windows_edge, e.adc_threshold)
        cfg = Config(det, elec, PadParams())
        rng = np.random.default_rng(424242)
        obs, one_point = [], []
        picker = np.random.default_rng(1)
        for i in range(n):
            cloud, _ = simulate(momenta[i].copy(), vertices[i], zs, as_, cfg, rng, indices)
            obs.append(event_observables(cloud))
            if len(cloud):  # one random point per event: an independent sample of the charge-per-point law
                one_point.append(cloud[picker.integers(len(cloud)), 2])
        obs = np.array(obs, dtype=np.float64)
        out[f"{name}/momenta"] = momenta
        out[f"{name}/vertices"] = vertices
        out[f"{name}/Z"] = zs
        out[f"{name}/A"] = as_
        out[f"{name}/indices"] = np.array(indices)
        out[f"{name}/n_points"] = obs[:, 0]
        out[f"{name}/sum_charge"] = obs[:, 1]
        out[f"{name}/tb_extent"] = obs[:, 2]
        out[f"{name}/n_pads"] = obs[:, 3]
        out[f"{name}/median_charge"] = obs[:, 4]
        out[f"{name}/max_charge"] = obs[:, 5]
        out[f"{name}/point_charge_sample"] = np.array(one_point)
        print(name, n, "events; mean points", obs[:, 0].mean(), "mean charge", obs[:, 1].mean())
    np.savez_compressed(HERE / "distributions.npz", **out)
    print((HERE / "distributions.npz").stat().st_size, "bytes")


if __name__ == "__main__":
    main()
