#!/bin/bash
# Round-2 measurement pass (one GPU): GPU tests, ncu pass, default bench with CPU baseline, the other workloads with
# CPU baseline, Spyral rows (typed and float64), float64 rows, reference arm, the config-5 pipeline.
set -u
mkdir -p gpurun_out
T=${TAG:-r2p}
python -m pytest tests -m gpu -q --timeout 600 --timeout-method=thread 2>&1 | tail -4 > gpurun_out/${T}_tests.log; tail -1 gpurun_out/${T}_tests.log
TAG=$T KERNELS="deposit_kernel collect_kernel emit_kernel track_kernel point_order_kernel spyral_rows_kernel spyral_count_kernel" BENCH_ARGS="--spyral" bash tools/profile_r2.sh
python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench_c16dd.log 2>&1; echo "c16dd rc=$?"; tail -1 gpurun_out/${T}_bench_c16dd.log | cut -c1-160
for w in c14dp c12aa sn132dp c16dd_sweep; do
  timeout 900 python bench.py --workload $w --events 16384 --steps 3 --warmup 3 > gpurun_out/${T}_bench_$w.log 2>&1
  echo "$w rc=$?"; tail -1 gpurun_out/${T}_bench_$w.log | cut -c1-160
done
timeout 600 python bench.py --spyral --steps 3 --warmup 3 --no-cpu > gpurun_out/${T}_bench_c16dd_spyral.log 2>&1; echo "spyral rc=$?"
timeout 600 python bench.py --spyral --float64-rows --steps 3 --warmup 3 --no-cpu > gpurun_out/${T}_bench_c16dd_spyral_float64rows.log 2>&1; echo "spyral f64 rc=$?"
timeout 600 python bench.py --float64-rows --steps 3 --warmup 3 --no-cpu > gpurun_out/${T}_bench_c16dd_float64rows.log 2>&1; echo "float64 rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.log 2>&1; echo "reference rc=$?"; tail -1 gpurun_out/${T}_bench_reference.log | cut -c1-200
timeout 900 python bench.py --pipeline --events-total 10000000 > gpurun_out/${T}_bench_pipeline_10M_1gpu.log 2>&1; echo "pipeline rc=$?"; tail -1 gpurun_out/${T}_bench_pipeline_10M_1gpu.log | cut -c1-300
