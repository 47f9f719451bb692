#!/bin/bash
# usage: bash tools/make_profiles.sh <full-set .ncu-rep> <launch-list csv> <bench n1 json> <bench reference json>
# copies / condenses the evidence of one measurement round into profiles/ (tracked)
set -e
cd "$(dirname "$0")/.."
REP=$1; LAUNCHES=$2; N1=$3; REF=$4
grep -v "^==" "$LAUNCHES" > profiles/r01_ncu_launches.csv
ncu -i "$REP" --page raw --csv > profiles/r01_ncu_k_fused_full.csv 2>/dev/null
python tools/sass_mix.py "$REP" 16000000 > profiles/r01_k_fused_sass_mix.txt
python tools/hot_code.py "$REP" >> profiles/r01_k_fused_sass_mix.txt
tail -1 "$N1" > profiles/r01_bench_n1.json
tail -1 "$REF" > profiles/r01_bench_reference.json
python - "$REP" <<'PY'
import csv, json, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); h, u, r = rows[0], rows[1], rows[2]
def val(n):
    v = float(r[h.index(n)]); un = u[h.index(n)]
    return v * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'Tbyte': 1e12}[un]
rd, wr = val('dram__bytes_read.sum'), val('dram__bytes_write.sum')
nch, n = 4096, 1000000
json.dump({"source": "ncu --set full --clock-control none, " + r[h.index('Kernel Name')] + ", profiles/r01_ncu_k_fused_full.csv (python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e)",
           "workload": "bank8psk", "channels": nch, "samples_per_channel": n, "dram_bytes_read": rd, "dram_bytes_write": wr,
           "dram_bytes_per_sample": (rd + wr) / (nch * n), "algorithmic_bytes_per_sample": 10.5}, open('profiles/r01_traffic.json', 'w'), indent=1)
print('traffic', rd, wr, (rd + wr) / (nch * n))
PY
