"""The REFERENCE's own `Config` for a bench.py workload (build container only, see ref_shim.py)."""

import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent.parent))

import ref_shim  # noqa: E402

ae = ref_shim.install()

from attpc_engine.detector import Config, DetectorParams, ElectronicsParams, PadParams  # noqa: E402
from spyral_utils.nuclear.target import GasTarget  # noqa: E402

import bench  # noqa: E402

_cache = {}


def reference_config(name):
    """Same numbers as ``bench.build_workload(name, ...)[0]``, but the reference's classes and its GasTarget (shim)."""
    if name not in _cache:
        ours = bench.build_workload(name, 1)[0]
        compound, pressure = bench.WORKLOADS[name]["gas"]
        gas = GasTarget(compound, pressure, ae.nuclear_map)
        d, e = ours.det_params, ours.elec_params
        det = DetectorParams(d.length, d.efield, d.bfield, d.mpgd_gain, gas, d.diffusion, d.fano_factor, d.w_value)
        elec = ElectronicsParams(e.clock_freq, e.amp_gain, e.shaping_time, e.micromegas_edge, e.windows_edge,
                                 e.adc_threshold)  # fmt: skip
        _cache[name] = Config(det, elec, PadParams())
    return _cache[name]
