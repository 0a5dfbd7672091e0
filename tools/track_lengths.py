import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, bench
from attpc_engine_b200.detector.engine import engine_for
from attpc_engine_b200.detector.simulator import _nuclei_for
from attpc_engine_b200 import nuclear_map
n = 16384
config, momenta, vertices, zs, as_, indices = bench.build_workload('c16dd', n)
nuclei = _nuclei_for(zs, as_, indices, nuclear_map)
eng = engine_for(config, nuclei)
for idx, nuc in zip(indices, nuclei):
    t0 = time.perf_counter()
    pts, counts = eng.trajectories(momenta[:, idx], vertices, [nuc] * n, stride=1, max_points=1)
    dt = time.perf_counter() - t0
    q = np.quantile(counts, [0.5, 0.9, 0.99, 0.999, 1.0])
    print(nuc.isotopic_symbol, 'mean', counts.mean(), 'quantiles 50/90/99/99.9/max', q, 'frac>=3000', (counts >= 3000).mean(), 'frac==10001', (counts == 10001).mean(), 'wall ms', round(dt * 1e3, 1))
