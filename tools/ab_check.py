"""A/B check of two builds of the library on the same events: prints a digest of the full result (run under gpurun).

    python tools/ab_check.py [path/to/libattpc_b200.so [abi]]
"""
import hashlib
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np

from attpc_engine_b200 import _lib

if len(sys.argv) > 1:
    _lib.LIB_PATH = Path(sys.argv[1]).resolve()
if len(sys.argv) > 2:
    _lib.ABI_VERSION = int(sys.argv[2])
import os

import bench
from attpc_engine_b200 import nuclear_map
from attpc_engine_b200.detector.engine import engine_for
from attpc_engine_b200.detector.simulator import _nuclei_for

for name, n in (("c16dd", 6000), ("c14dp", 3000), ("c12aa", 1500), ("sn132dp", 2000)):
    config, momenta, vertices, zs, as_, indices = bench.build_workload(name, n)
    tune = {}
    if os.environ.get("AB_RTOL"):
        tune = dict(ode_rtol=float(os.environ["AB_RTOL"]), ode_atol=float(os.environ["AB_RTOL"]) * 1e-4)
    eng = engine_for(config, _nuclei_for(zs, as_, indices, nuclear_map), **tune)
    b = eng.simulate_batch(momenta, vertices, zs, as_, indices, seed=77, first_event=1000)
    hsh = hashlib.sha256()
    for arr in (b.offsets, b.cloud, b.labels):
        hsh.update(np.ascontiguousarray(arr).tobytes())
    st = b.stats
    print(name, n, "rows", len(b.labels), "traj", st["n_trajectory_points"], "active", st["n_active_points"],
          "prim", st["n_primary_electrons"], "ms_tracks", round(st["ms_tracks"], 3), hsh.hexdigest()[:16], flush=True)
