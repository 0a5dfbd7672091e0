#!/bin/bash
# e2e against the host-copy chunk size, raw typed cloud and Spyral typed rows
set -u
mkdir -p gpurun_out
T=${TAG:-r2x}
for ce in 2048 4096 8192 16384; do
  timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu --copy-events $ce > gpurun_out/${T}_raw_c$ce.log 2>&1; echo "raw $ce rc=$?"
  timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu --spyral --copy-events $ce > gpurun_out/${T}_spyral_c$ce.log 2>&1; echo "spyral $ce rc=$?"
done
