import sys; sys.path.insert(0,'/root/repo')
import numpy as np, bench, torch
from attpc_engine_b200.detector.engine import engine_for
from attpc_engine_b200.detector.simulator import _nuclei_for
from attpc_engine_b200 import nuclear_map
config, momenta, vertices, zs, as_, indices = bench.build_workload('c16dd', 8192)
eng = engine_for(config, _nuclei_for(zs, as_, indices, nuclear_map))
for i in range(3):
    st = eng.simulate_batch(momenta, vertices, zs, as_, indices, seed=i, copy=False).stats
    print({k: st[k] for k in ('n_deposits','n_hash_probes','hash_capacity','n_retries','n_keys','n_points','ms_tracks','ms_deposit','ms_finalize','ms_d2h','ms_total')})
