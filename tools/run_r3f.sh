#!/bin/bash
set -u
mkdir -p gpurun_out
T=${TAG:-r3f}
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 --timeout-method=thread 2>&1 | tail -8 > gpurun_out/${T}_tests.log; tail -1 gpurun_out/${T}_tests.log
cp attpc_engine_b200/libattpc_b200.so /tmp/default.so
for so in build/variants/*.so; do
  name=$(basename $so .so)
  cp $so attpc_engine_b200/libattpc_b200.so
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/${T}_${name}_c16dd.log 2>&1; echo "$name rc=$?"
  timeout 300 python bench.py --workload c12aa --events 16384 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/${T}_${name}_c12aa.log 2>&1; echo "$name c12aa rc=$?"
done
cp /tmp/default.so attpc_engine_b200/libattpc_b200.so
