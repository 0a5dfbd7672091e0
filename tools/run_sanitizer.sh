#!/bin/bash
# compute-sanitizer over the replay-parity and work-splitting tests (one GPU).  memcheck: out-of-bounds / misaligned
# accesses in any kernel; synccheck: barrier misuse; racecheck: shared-memory hazards (the deposit kernel's table is
# written with atomics and read with volatile loads on purpose: see profiles/r02_sanitizer_README.md for the reading).
set -u
mkdir -p gpurun_out
SEL='test_simulate_cloud or test_three_events or (test_workload_replay_dict and c16dd) or (test_workload_replay_cloud_and_spyral and c14dp)'
for tool in memcheck synccheck racecheck; do
  extra=""
  [ "$tool" = racecheck ] && extra="--racecheck-report analysis"
  timeout 1500 compute-sanitizer --tool $tool $extra --print-limit 40 --error-exitcode 0 \
      --log-file gpurun_out/r02_sanitizer_$tool.txt \
      python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" -p no:cacheprovider > gpurun_out/r02_sanitizer_${tool}_pytest.log 2>&1
  echo "$tool rc=$? $(tail -1 gpurun_out/r02_sanitizer_${tool}_pytest.log)"
  grep -c "=========" gpurun_out/r02_sanitizer_$tool.txt
  tail -3 gpurun_out/r02_sanitizer_$tool.txt
done
# the production path (integrator + Philox + chunked copies) under memcheck
timeout 900 compute-sanitizer --tool memcheck --print-limit 40 --error-exitcode 0 --log-file gpurun_out/r02_sanitizer_memcheck_e2e.txt \
    python -m pytest tests/test_gpu_e2e.py -m gpu -q -x -k "work_splitting and c16dd or spyral_columns and c16dd or small_launches" -p no:cacheprovider > gpurun_out/r02_sanitizer_memcheck_e2e_pytest.log 2>&1
echo "memcheck e2e rc=$? $(tail -1 gpurun_out/r02_sanitizer_memcheck_e2e_pytest.log)"; tail -3 gpurun_out/r02_sanitizer_memcheck_e2e.txt
