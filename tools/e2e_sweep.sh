#!/bin/bash
# e2e throughput against launch size and host-copy chunk size
for le in 0 8192 4096; do for ce in 0 4096; do
  python bench.py --steps 3 --warmup 3 --no-cpu --launch-events $le --copy-events $ce 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('launch $le copy $ce', 'device', d['value'], 'e2e', d['e2e']['value'])"
done; done
