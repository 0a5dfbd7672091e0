#!/bin/bash
# Round-2 evidence pass on one GPU: tests, ncu, bench lines of BASELINE configs 1-5 with CPU baselines, pipeline
set -u
mkdir -p gpurun_out
T=${TAG:-r2final}
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --timeout-method=thread 2>&1 | tail -6 > gpurun_out/${T}_tests.log; tail -1 gpurun_out/${T}_tests.log
TAG=$T KERNELS="deposit_kernel ^order_kernel order_queue_kernel ^emit_kernel track_kernel point_order_kernel" BENCH_ARGS="" bash tools/profile_r2.sh
python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_c16dd.log 2>&1; echo "c16dd rc=$?"; tail -1 gpurun_out/${T}_bench_c16dd.log | cut -c1-160
timeout 600 python bench.py --spyral --steps 20 --warmup 5 --no-cpu > gpurun_out/${T}_bench_c16dd_spyral.log 2>&1; echo "spyral rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.log 2>&1; echo "reference rc=$?"
timeout 900 python bench.py --workload c14dp --events 32768 --steps 31 --warmup 3 > gpurun_out/${T}_bench_c14dp_1M.log 2>&1; echo "c14dp rc=$?"
timeout 900 python bench.py --workload c12aa --events 16384 --steps 62 --warmup 3 > gpurun_out/${T}_bench_c12aa_1M.log 2>&1; echo "c12aa rc=$?"
timeout 900 python bench.py --workload sn132dp --events 16384 --steps 7 --warmup 3 > gpurun_out/${T}_bench_sn132dp_100k.log 2>&1; echo "sn132dp rc=$?"
timeout 900 python bench.py --workload c16dd_sweep --events 32768 --steps 10 --warmup 3 > gpurun_out/${T}_bench_c16dd_sweep.log 2>&1; echo "sweep rc=$?"
timeout 900 python bench.py --pipeline --events-total 10000000 > gpurun_out/${T}_bench_pipeline_10M_1gpu.log 2>&1; echo "pipeline rc=$?"
