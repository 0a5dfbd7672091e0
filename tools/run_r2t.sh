#!/bin/bash
# finalize_kernel: GPU tests, then the device-resident step for every build/variants/*.so, without and with Spyral rows
set -u
mkdir -p gpurun_out
T=${TAG:-r2t}
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 --timeout-method=thread 2>&1 | tail -15 > gpurun_out/${T}_tests.log; tail -1 gpurun_out/${T}_tests.log
cp attpc_engine_b200/libattpc_b200.so /tmp/default.so
for so in build/variants/*.so; do
  name=$(basename $so .so)
  cp $so attpc_engine_b200/libattpc_b200.so
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/${T}_${name}.log 2>&1; echo "$name rc=$?"
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --spyral > gpurun_out/${T}_${name}_spyral.log 2>&1; echo "$name spyral rc=$?"
done
cp /tmp/default.so attpc_engine_b200/libattpc_b200.so
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/${T}_bench.log 2>&1; echo "bench rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --spyral > gpurun_out/${T}_bench_spyral.log 2>&1; echo "bench spyral rc=$?"
timeout 600 python bench.py --workload c12aa --events 16384 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/${T}_bench_c12aa.log 2>&1; echo "c12aa rc=$?"
