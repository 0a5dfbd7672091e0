#!/bin/bash
# full GPU tests + default bench (with CPU baseline) + Spyral + float64 rows
set -u
mkdir -p gpurun_out
T=${TAG:-r3a}
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 --timeout-method=thread 2>&1 | tail -15 > gpurun_out/${T}_tests.log; tail -1 gpurun_out/${T}_tests.log
timeout 600 python bench.py --steps 6 --warmup 3 ${NOCPU:-} > gpurun_out/${T}_bench_c16dd.log 2>&1; echo "bench rc=$?"
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu --spyral > gpurun_out/${T}_bench_c16dd_spyral.log 2>&1; echo "spyral rc=$?"
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu --float64-rows > gpurun_out/${T}_bench_c16dd_float64rows.log 2>&1; echo "f64 rc=$?"
