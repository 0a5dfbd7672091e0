import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, bench
from attpc_engine_b200.detector.engine import engine_for
from attpc_engine_b200.detector.simulator import _nuclei_for
from attpc_engine_b200 import nuclear_map
name = sys.argv[1] if len(sys.argv) > 1 else 'c16dd'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
config, momenta, vertices, zs, as_, indices = bench.build_workload(name, B)
eng = engine_for(config, _nuclei_for(zs, as_, indices, nuclear_map))
for i in range(2):
    b = eng.simulate_batch(momenta, vertices, zs, as_, indices, seed=i, copy=False)
st = b.stats
n = np.diff(b.offsets)
print(name, {k: st[k] for k in ('n_keys', 'n_points', 'n_table_flushes', 'n_retries', 'hash_capacity', 'n_deposits', 'n_hash_probes', 'n_active_points')})
print('points/event quantiles 50/90/99/99.9/max', np.quantile(n, [0.5, 0.9, 0.99, 0.999, 1.0]), 'frac > 6080:', (n > 6080).mean(), 'frac > 4700', (n > 4700).mean())
