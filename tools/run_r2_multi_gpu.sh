#!/bin/bash
# N-GPU pass: default bench (raw cloud, packed columns), Spyral typed rows, config-5 pipeline (10 M events)
set -u
mkdir -p gpurun_out
T=${TAG:-r2multi}
N=${N:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${T}_bench_${N}gpu.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/${T}_bench_${N}gpu.log | cut -c1-200
timeout 400 $TR bench.py --gpus $N --steps 20 --warmup 5 --spyral > gpurun_out/${T}_bench_spyral_${N}gpu.log 2>&1; echo "spyral rc=$?"
timeout 600 $TR bench.py --gpus $N --pipeline --events-total 10000000 > gpurun_out/${T}_pipeline_10M_${N}gpu.log 2>&1; echo "pipeline rc=$?"; tail -1 gpurun_out/${T}_pipeline_10M_${N}gpu.log | cut -c1-300
