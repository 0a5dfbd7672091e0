#!/bin/bash
# e2e throughput against host-copy chunk size (and launch size)
for ce in 0 4096 8192; do
  python bench.py --steps 4 --warmup 3 --no-cpu --copy-events $ce 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('copy $ce', 'device', d['value'], 'e2e', d['e2e']['value'])"
done
