#!/usr/bin/env python
"""hot-code footprint of the first kernel of an ncu report: hot_code.py rep [nchunks]"""
import csv, subprocess, sys
rep = sys.argv[1]; nch = float(sys.argv[2]) if len(sys.argv) > 2 else 16e6
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; data = []
for r in rows:
    if r and r[0] == 'Address': hdr = r; continue
    if hdr and r and r[0].startswith('0x'):
        g = lambda k: int(r[hdr.index(k)] or 0)
        data.append((int(r[0], 16), r[1].strip(), g('Instructions Executed'), g('stall_no_inst'), g('# Samples'), g('stall_long_sb'), g('stall_short_sb'), g('stall_wait')))
base = data[0][0]
print('total instrs', len(data), 'bytes', len(data) * 16)
thr = nch / 16
hot = [d for d in data if d[2] >= thr]
lines = set(d[0] // 128 for d in hot)
print(f'exec >= {thr:.0f}: {len(hot)} instrs = {len(hot)*16/1024:.1f} KB; 128-B lines touched: {len(lines)} = {len(lines)/8:.1f} KB')
runs = []; cur = None
for d in data:
    if d[2] >= thr:
        if cur is None: cur = [d[0], d[0], 0, 0, 0, 0, 0, 0, 0]
        cur[1] = d[0]; cur[2] += 1; cur[3] += d[2]; cur[4] += d[3]; cur[5] += d[4]; cur[6] += d[5]; cur[7] += d[6]; cur[8] += d[7]
    else:
        if cur: runs.append(cur); cur = None
if cur: runs.append(cur)
tot_s = sum(d[4] for d in data); tot_n = sum(d[3] for d in data)
print(len(runs), 'hot runs; total samples', tot_s, 'no_inst', tot_n)
print('   start     end  instrs  exec/chunk  samples%  no_inst%  long_sb short_sb wait (% of all samples)')
for r in runs:
    if r[2] >= 6 or r[5] > tot_s * 0.003:
        print(f'{r[0]-base:8x} {r[1]-base:8x} {r[2]:6d} {r[3]/nch:10.1f} {100*r[5]/tot_s:8.1f} {100*r[4]/tot_s:8.1f} {100*r[6]/tot_s:8.1f} {100*r[7]/tot_s:8.1f} {100*r[8]/tot_s:8.1f}')
