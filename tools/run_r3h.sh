#!/bin/bash
set -u
mkdir -p gpurun_out
T=${TAG:-r3h}
timeout 600 python -m pytest tests -m gpu -x -q -k "e2e or parity" --timeout 900 --timeout-method=thread 2>&1 | tail -3 > gpurun_out/${T}_tests.log; tail -1 gpurun_out/${T}_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_c16dd.log 2>&1; echo "c16dd rc=$?"
timeout 600 python bench.py --spyral --steps 20 --warmup 5 --no-cpu > gpurun_out/${T}_bench_c16dd_spyral.log 2>&1; echo "spyral rc=$?"
timeout 900 python bench.py --workload c12aa --events 16384 --steps 62 --warmup 3 > gpurun_out/${T}_bench_c12aa_1M.log 2>&1; echo "c12aa rc=$?"
timeout 900 python bench.py --workload c16dd_sweep --events 32768 --steps 10 --warmup 3 > gpurun_out/${T}_bench_c16dd_sweep.log 2>&1; echo "sweep rc=$?"
