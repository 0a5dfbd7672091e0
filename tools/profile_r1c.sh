#!/bin/bash
# Round-1 final ncu captures (run under gpurun, one GPU): launch list + full sets of the four main kernels.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep
CMD="python bench.py --steps 1 --warmup 2 --events 16384 --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:track_kernel -s 3 -c 1 -f -o gpurun_out/prof_track_kernel $CMD > gpurun_out/ncu_track.log 2>&1; echo "track rc=$?"
for k in deposit_kernel collect_kernel emit_kernel point_order_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 40 -c 1 -f -o gpurun_out/prof_$k $CMD > gpurun_out/ncu_$k.log 2>&1
  echo "$k rc=$?"
done
ls -la gpurun_out | head -30
