#!/bin/bash
set -u
mkdir -p gpurun_out
T=${TAG:-r3k}
cp attpc_engine_b200/libattpc_b200.so /tmp/default.so
for so in build/variants/*.so; do
  name=$(basename $so .so)
  cp $so attpc_engine_b200/libattpc_b200.so
  timeout 300 python bench.py --steps 12 --warmup 4 --no-cpu > gpurun_out/${T}_${name}_raw.log 2>&1; echo "$name rc=$?"
  timeout 300 python bench.py --steps 12 --warmup 4 --no-cpu --spyral > gpurun_out/${T}_${name}_spyral.log 2>&1; echo "$name spyral rc=$?"
done
cp /tmp/default.so attpc_engine_b200/libattpc_b200.so
