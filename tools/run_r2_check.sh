#!/bin/bash
set -u
mkdir -p gpurun_out
T=${TAG:-r2check}
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 --timeout-method=thread 2>&1 | tail -8 > gpurun_out/${T}_tests.log; tail -1 gpurun_out/${T}_tests.log
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/${T}_bench_c16dd.log 2>&1; echo "c16dd rc=$?"
timeout 900 python bench.py --workload c12aa --events 16384 --steps 20 --warmup 3 --no-cpu > gpurun_out/${T}_bench_c12aa.log 2>&1; echo "c12aa rc=$?"
timeout 600 python bench.py --spyral --steps 20 --warmup 5 --no-cpu > gpurun_out/${T}_bench_c16dd_spyral.log 2>&1; echo "spyral rc=$?"
