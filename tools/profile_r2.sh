#!/bin/bash
# Round-2 ncu pass (one GPU): plain run, launch list, full captures of the kernels named in $KERNELS.
set -u
mkdir -p gpurun_out
TAG=${TAG:-r2}
CMD="python bench.py --steps 1 --warmup 2 --no-cpu --no-e2e ${BENCH_ARGS:-}"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
tail -1 gpurun_out/${TAG}_plain.log | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
for k in ${KERNELS:-deposit_kernel emit_kernel track_kernel point_order_kernel event_sort_kernel}; do
  n=${k//[^a-z_]/}
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_$n $CMD > gpurun_out/${TAG}_ncu_$n.log 2>&1
  echo "$k rc=$?"
done
