#!/bin/bash
# Round-1 final measurement pass (one GPU): GPU tests, ncu pass, full default bench, the other workloads, Spyral rows,
# float64 rows, reference arm.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
bash tools/profile_r1f.sh
python bench.py --steps 5 --warmup 3 > gpurun_out/r1g_bench_c16dd.log 2>&1; echo "c16dd rc=$?"; tail -1 gpurun_out/r1g_bench_c16dd.log | cut -c1-160
for w in c14dp c12aa sn132dp; do
  timeout 600 python bench.py --workload $w --events 16384 --steps 3 --warmup 3 --no-cpu > gpurun_out/r1g_bench_$w.log 2>&1
  echo "$w rc=$?"; tail -1 gpurun_out/r1g_bench_$w.log | cut -c1-160
done
timeout 600 python bench.py --spyral --steps 3 --warmup 3 --no-cpu > gpurun_out/r1g_bench_c16dd_spyral.log 2>&1; echo "spyral rc=$?"
timeout 600 python bench.py --float64-rows --steps 3 --warmup 3 --no-cpu > gpurun_out/r1g_bench_c16dd_float64rows.log 2>&1; echo "float64 rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1g_bench_reference.log 2>&1; echo "reference rc=$?"; tail -1 gpurun_out/r1g_bench_reference.log | cut -c1-200
