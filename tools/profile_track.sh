#!/bin/bash
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 2 --no-cpu --no-e2e"
$CMD > gpurun_out/plain_track.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_track.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:track_kernel -s 2 -c 1 -f -o gpurun_out/prof_track_kernel $CMD > gpurun_out/ncu_track.log 2>&1; echo "track rc=$?"
