#!/bin/bash
set -u
mkdir -p gpurun_out
T=${TAG:-r3i}
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --timeout-method=thread 2>&1 | tail -8 > gpurun_out/${T}_tests.log; tail -1 gpurun_out/${T}_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log
