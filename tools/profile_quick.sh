#!/bin/bash
# ncu --set full of selected kernels of the timed device-resident step (KERNELS="deposit_kernel collect_kernel")
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 2 --no-cpu --no-e2e"
for k in ${KERNELS:-deposit_kernel}; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof_$k $CMD > gpurun_out/ncu_$k.log 2>&1
  echo "$k rc=$?"
done
