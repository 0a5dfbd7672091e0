#!/bin/bash
# ncu captures of the three stages (run under gpurun, one GPU).  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 2 --events 8192 --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
for k in track_kernel deposit_kernel collect_kernel emit_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s $( [ $k = track_kernel ] && echo 4 || echo 70 ) -c 2 -f -o gpurun_out/prof_$k $CMD > gpurun_out/ncu_$k.log 2>&1
  echo "$k rc=$?"
done
ls -la gpurun_out
