#!/bin/bash
set -u
mkdir -p gpurun_out
T=${TAG:-r3d}
timeout 900 python bench.py --workload c14dp --events 32768 --steps 31 --warmup 3 > gpurun_out/${T}_bench_c14dp_1M.log 2>&1; echo "c14dp rc=$?"
timeout 900 python bench.py --workload c12aa --events 16384 --steps 62 --warmup 3 > gpurun_out/${T}_bench_c12aa_1M.log 2>&1; echo "c12aa rc=$?"
timeout 900 python bench.py --workload c16dd_sweep --events 32768 --steps 10 --warmup 3 > gpurun_out/${T}_bench_c16dd_sweep.log 2>&1; echo "sweep rc=$?"
timeout 900 python bench.py --pipeline --events-total 10000000 > gpurun_out/${T}_bench_pipeline_10M_1gpu.log 2>&1; echo "pipeline rc=$?"; tail -1 gpurun_out/${T}_bench_pipeline_10M_1gpu.log | cut -c1-300
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
