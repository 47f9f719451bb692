#!/bin/bash
# usage (GPU box): bash tools/bench_packets.sh samples...   -- headline bank, one call per `samples` complex samples per channel
cd "$(dirname "$0")/.."
for n in "$@"; do
  python bench.py --samples $n --steps 20 --warmup 5 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('samples/call', d['config']['samples_per_channel'], round(d['ms_per_step'],3), 'ms', round(d['value']/1e3,1), 'Gsamples/s', {k:round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()})"
done
