#!/bin/bash
for k in 1500 2000 2500 3000 3488 4200; do
  python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --tune table_spill_keys=$k 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('spill $k', d['value'], d['ms_per_step'], d['roofline']['stage_ms_per_step'], d['per_event']['table_flushes'])"
done
for u in 256 512; do
  python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --tune unit_points=$u 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('unit $u', d['value'], d['ms_per_step'], d['roofline']['stage_ms_per_step'], d['per_event']['table_flushes'])"
done
