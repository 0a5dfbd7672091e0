#!/bin/bash
set -u
mkdir -p gpurun_out
T=${TAG:-r3b}
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 --timeout-method=thread 2>&1 | tail -15 > gpurun_out/${T}_tests.log; tail -1 gpurun_out/${T}_tests.log
for eng in 2 3; do
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --e2e-engines $eng > gpurun_out/${T}_raw_e$eng.log 2>&1; echo "raw $eng rc=$?"
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --spyral --e2e-engines $eng > gpurun_out/${T}_spyral_e$eng.log 2>&1; echo "spyral $eng rc=$?"
done
