#!/bin/bash
# usage (GPU box): bash tools/bench_e2e.sh slabs...   -- e2e (host-buffer) rate of the headline workload per slab count
cd "$(dirname "$0")/.."
for s in "$@"; do
  PSKD_SLABS=$s python bench.py --steps 3 --warmup 2 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('slabs $s', round(e['value']/1e3,3), 'Gsamples/s', round(e['ms_per_step'],2), 'ms', round(e['h2d_bytes_per_step']/e['ms_per_step']/1e6,1), 'GB/s H2D')"
done
