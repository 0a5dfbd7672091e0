#!/bin/bash
# ncu captures (source-level) of deposit_kernel and collect_kernel at steady state.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 2 --events 8192 --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
for k in deposit_kernel collect_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 20 -c 1 -f -o gpurun_out/prof_$k $CMD > gpurun_out/ncu_$k.log 2>&1
  echo "$k rc=$?"
done
