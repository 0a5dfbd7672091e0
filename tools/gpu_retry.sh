#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3 / status=transient): tools/gpu_retry.sh <timeout_s> '<command>'
t=$1; shift
for attempt in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if ! grep -q "status=transient" /tmp/gpurun_last.log; then cat /tmp/gpurun_last.log | tail -4; exit $rc; fi
  sleep 45
done
tail -3 /tmp/gpurun_last.log; exit 3
