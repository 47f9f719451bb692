#!/bin/bash
# usage: bash tools/make_profiles.sh   (after tools/r2_profile_run.sh has run on the GPU box and gpurun_out/ came back)
# condenses the evidence of the round into profiles/ (tracked)
set -e
cd "$(dirname "$0")/.."
R=r02
grep -v "^==" gpurun_out/${R}_launches_n1.csv > profiles/${R}_ncu_launches.csv
grep -v "^==" gpurun_out/${R}_launches_512.csv > profiles/${R}_ncu_launches_512ch.csv
ncu -i gpurun_out/prof_${R}_fused.ncu-rep --page raw --csv > profiles/${R}_ncu_k_fused_full.csv 2>/dev/null
ncu -i gpurun_out/prof_${R}_fzs512.ncu-rep --page raw --csv > profiles/${R}_ncu_k_fzs_512ch_full.csv 2>/dev/null
python tools/sass_mix.py gpurun_out/prof_${R}_fused.ncu-rep 16000000 > profiles/${R}_k_fused_sass_mix.txt
python tools/hot_code.py gpurun_out/prof_${R}_fused.ncu-rep >> profiles/${R}_k_fused_sass_mix.txt
tail -1 gpurun_out/${R}_bench_n1.json > profiles/${R}_bench_n1.json
tail -1 gpurun_out/${R}_bench_n1_steps20.json > profiles/${R}_bench_n1_steps20.json
tail -1 gpurun_out/${R}_bench_reference.json > profiles/${R}_bench_reference.json
python - <<'PY'
import csv, json, subprocess
def load(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines())); return rows[0], rows[1], rows[2:]
def val(h, u, r, n):
    v = float(r[h.index(n)]); un = u[h.index(n)]
    return v * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'Tbyte': 1e12, 'ms': 1e-3, 'us': 1e-6, 'ns': 1e-9, 's': 1.0, 'msecond': 1e-3, 'usecond': 1e-6, 'nsecond': 1e-9, 'second': 1.0}.get(un, 1.0)
h, u, rs = load('gpurun_out/prof_r02_fused.ncu-rep'); r = rs[0]
rd, wr = val(h, u, r, 'dram__bytes_read.sum'), val(h, u, r, 'dram__bytes_write.sum')
nch, n = 4096, 1000000
json.dump({"source": "ncu --set full --clock-control none, " + r[h.index('Kernel Name')] + ", profiles/r02_ncu_k_fused_full.csv (python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e)",
           "kernel": "k_fused", "workload": "bank8psk", "channels": nch, "samples_per_channel": n, "dram_bytes_read": rd, "dram_bytes_write": wr,
           "dram_bytes_per_sample": (rd + wr) / (nch * n), "algorithmic_bytes_per_sample": 10.5}, open('profiles/r02_traffic.json', 'w'), indent=1)
print('k_fused traffic', rd, wr, (rd + wr) / (nch * n))
h, u, rs = load('gpurun_out/prof_r02_fzs512.ncu-rep')
best = {}
for r in rs:
    name = r[h.index('Kernel Name')].split('(')[0]
    t = val(h, u, r, 'gpu__time_duration.sum')
    if name not in best or t > best[name][0]: best[name] = (t, r)
out = {}
for name, (t, r) in best.items():
    out[name] = {"duration_ms": t * 1e3, "dram_bytes_read": val(h, u, r, 'dram__bytes_read.sum'), "dram_bytes_write": val(h, u, r, 'dram__bytes_write.sum'),
                 "warp_instructions": float(r[h.index('smsp__inst_executed.sum')]), "issue_active_pct": float(r[h.index('smsp__issue_active.avg.pct_of_peak_sustained_active')]),
                 "registers": int(float(r[h.index('launch__registers_per_thread')]))}
json.dump({"source": "ncu --set full --clock-control none -k regex:k_fzs, largest launch of each kernel, profiles/r02_ncu_k_fzs_512ch_full.csv (python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --channels 512)",
           "workload": "bank8psk, 512 channels x 1M (the per-GPU shard of the N=8 strong-scaling run)", "kernels": out}, open('profiles/r02_fzs_512ch.json', 'w'), indent=1)
print(json.dumps(out, indent=1))
PY
