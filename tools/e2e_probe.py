import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch, bench
from attpc_engine_b200.detector.engine import engine_for
from attpc_engine_b200.detector.simulator import _nuclei_for
from attpc_engine_b200 import nuclear_map
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
L = int(sys.argv[2]) if len(sys.argv) > 2 else 0
config, momenta, vertices, zs, as_, indices = bench.build_workload('c16dd', B)
eng = engine_for(config, _nuclei_for(zs, as_, indices, nuclear_map), copy_events_per_launch=L)
mp = torch.from_numpy(momenta).pin_memory(); vp = torch.from_numpy(vertices).pin_memory()
md = mp.cuda(); vd = vp.cuda()
keys = ('ms_h2d','ms_tracks','ms_deposit','ms_finalize','ms_d2h','ms_total','n_retries','n_points','n_kernel_launches')
for i in range(4):
    t0 = time.perf_counter(); st = eng.simulate_batch(mp.numpy(), vp.numpy(), zs, as_, indices, seed=i, copy=False, columns=True).stats; w = time.perf_counter() - t0
    print('e2e ', round(w * 1e3, 1), 'ms wall', {k: round(st[k], 2) if isinstance(st[k], float) else st[k] for k in keys})
for i in range(3):
    t0 = time.perf_counter(); st = eng.simulate_device(md.data_ptr(), vd.data_ptr(), B, 4, zs, as_, indices, seed=i).stats; w = time.perf_counter() - t0
    print('dev ', round(w * 1e3, 1), 'ms wall', {k: round(st[k], 2) if isinstance(st[k], float) else st[k] for k in keys})
