#!/bin/bash
# usage (GPU box): bash tools/bench_channels.sh n1 n2 ...  -- step time of the headline workload at other bank sizes
cd "$(dirname "$0")/.."
for n in "$@"; do
  python bench.py --channels $n --steps 6 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['config']['channels_per_gpu'], 'channels', round(d['ms_per_step'],3), 'ms', round(d['value']/1e3,1), 'Gsamples/s', 'frac', round(d['roofline']['whole_path']['frac'],3))"
done
