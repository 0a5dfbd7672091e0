"""Track-kernel latency vs throughput: device time and integrator statistics for growing batches (run under gpurun)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import bench
from attpc_engine_b200 import nuclear_map
from attpc_engine_b200.detector.engine import engine_for
from attpc_engine_b200.detector.simulator import _nuclei_for

name = sys.argv[1] if len(sys.argv) > 1 else "c16dd"
config, momenta, vertices, zs, as_, indices = bench.build_workload(name, 65536)
eng = engine_for(config, _nuclei_for(zs, as_, indices, nuclear_map), max_events_per_launch=65536)
K = momenta.shape[1]
for n in (512, 2048, 8192, 32768, 65536):
    m = torch.from_numpy(momenta[:n]).cuda()
    v = torch.from_numpy(vertices[:n]).cuda()
    for rep in range(3):
        st = eng.simulate_device(m.data_ptr(), v.data_ptr(), n, K, zs, as_, indices, seed=1 + rep).stats
    print(name, "events", n, "ms_tracks", round(st["ms_tracks"], 3), "ms_deposit", round(st["ms_deposit"], 3),
          "ms_finalize", round(st["ms_finalize"], 3), "ms_total", round(st["ms_total"], 3),
          "steps/track", round(st["n_rk_steps"] / st["n_tracks"], 1), "rejects/track", round(st["n_rk_rejects"] / st["n_tracks"], 2),
          "max passes", st["max_track_passes"], "traj pts/track", round(st["n_trajectory_points"] / st["n_tracks"], 1), flush=True)
