#!/usr/bin/env python
"""warp instructions per source line of the biggest launch in an ncu report (needs -lineinfo + --import-source on):
   src_lines.py report.ncu-rep [units (e.g. 32-symbol chunks) to normalise by] [top N]"""
import csv, collections, subprocess, sys
rep = sys.argv[1]; nch = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0; top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass,cuda'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
num = lambda x: int(x) if x.strip().lstrip('-').isdigit() else 0
launches = []; cur = None; fpath = None; hdr = None
for r in rows:
    if not r: continue
    if r[0] == 'File Path':
        fpath = r[1]
        if cur is None or fpath in cur['files']:
            cur = {'files': set(), 'rows': []}; launches.append(cur)
        cur['files'].add(fpath); continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr and r[0].isdigit(): cur['rows'].append((fpath, r))
iE = hdr.index('Instructions Executed'); iS = hdr.index('# Samples')
big = max(launches, key=lambda L: sum(num(r[iE]) for f, r in L['rows']))
agg = collections.Counter(); samp = collections.Counter(); src = {}
for f, r in big['rows']:
    k = (f.split('/')[-1], int(r[0])); agg[k] += num(r[iE]); samp[k] += num(r[iS]); src[k] = r[1].strip()[:100]
tot = sum(agg.values()); stot = max(sum(samp.values()), 1)
print(f'total warp instructions {tot} = {tot / nch:.1f} per unit')
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
    print(f'{k[0]:18s}{k[1]:5d} {v / nch:7.2f} {100 * samp[k] / stot:5.1f}%  {src[k]}')
