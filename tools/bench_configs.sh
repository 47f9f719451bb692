#!/bin/bash
# usage (GPU box): bash tools/bench_configs.sh [workload ...]   -- device-resident step time of the parity-case configs
cd "$(dirname "$0")/.."
for w in "${@:-config4 config3 config2 config1}"; do
  for ww in $w; do
    python bench.py --workload $ww --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['config']['workload'], round(d['ms_per_step'],3), 'ms', round(d['value']/1e3,1), 'Gsamples/s', {k:round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items()}, 'whole-path frac', round(d['roofline']['whole_path']['frac'],3))"
  done
done
