#!/bin/bash
# device-resident step time for different numbers of groups per kernel launch
for c in 1 2 4 8 16; do
  echo "chunk_groups=$c"
  ATTPC_CHUNK_GROUPS=$c python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['stage_ms_per_step'], d['gpu_launches'])"
done
ATTPC_CHUNK_GROUPS=1 python tools/ab_check.py | cut -c1-200
ATTPC_CHUNK_GROUPS=16 python tools/ab_check.py | cut -c1-200
