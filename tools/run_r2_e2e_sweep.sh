#!/bin/bash
# pipelined e2e (simulate_stream): engines per GPU x copy chunk, raw typed cloud and Spyral typed rows; e2e GPU tests
set -u
mkdir -p gpurun_out
T=${TAG:-r2e2e}
timeout 900 python -m pytest tests/test_gpu_e2e.py -x -q --timeout 600 --timeout-method=thread 2>&1 | tail -5 > gpurun_out/${T}_tests.log; tail -1 gpurun_out/${T}_tests.log
for eng in 2 3; do
 for ce in 2048 4096 8192; do
  timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --e2e-engines $eng --copy-events $ce > gpurun_out/${T}_raw_e${eng}_c$ce.log 2>&1; echo "raw $eng $ce rc=$?"
  timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --spyral --e2e-engines $eng --copy-events $ce > gpurun_out/${T}_spyral_e${eng}_c$ce.log 2>&1; echo "spyral $eng $ce rc=$?"
 done
done
