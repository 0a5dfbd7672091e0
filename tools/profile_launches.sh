#!/bin/bash
# ncu launch list (every launch with its device time) of the device-resident bench step.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e"
$CMD > gpurun_out/plain_launches.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_launches.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 345 -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
