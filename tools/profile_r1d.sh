#!/bin/bash
# Round-1 (session 2) baseline: plain bench, launch list, full captures of deposit / collect / track with source.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e"
$CMD > gpurun_out/plain_launches.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_launches.log; exit 1; }
tail -1 gpurun_out/plain_launches.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -s 345 -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
for k in ${KERNELS:-deposit_kernel collect_kernel track_kernel}; do
  skip=40; [ $k = track_kernel ] && skip=3
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/prof_$k $CMD > gpurun_out/ncu_$k.log 2>&1
  echo "$k rc=$?"
done
