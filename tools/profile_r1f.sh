#!/bin/bash
# Round-1 session-3 ncu pass (one GPU): plain run, launch list, full captures of the main kernels of the timed step
# (chunked launches: one launch of every group kernel per 32768-event step).
set -u
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep
CMD="python bench.py --steps 1 --warmup 2 --no-cpu --no-e2e"
$CMD > gpurun_out/plain_launches.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_launches.log; exit 1; }
tail -1 gpurun_out/plain_launches.log | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
for k in ${KERNELS:-deposit_kernel collect_kernel emit_kernel track_kernel point_order_kernel}; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof_$k $CMD > gpurun_out/ncu_$k.log 2>&1
  echo "$k rc=$?"
done
