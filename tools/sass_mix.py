#!/usr/bin/env python
"""Opcode mix / hot source lines of every kernel in an ncu report (source page).
usage: sass_mix.py report.ncu-rep [--lines N]"""
import csv, re, collections, subprocess, sys
rep = sys.argv[1]
nlines = int(sys.argv[sys.argv.index('--lines') + 1]) if '--lines' in sys.argv else 0
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'] + (['--print-source', 'sass,cuda'] if False else []),
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kern = None; hdr = None; data = collections.OrderedDict()
for r in rows:
    if r and r[0] == 'Kernel Name': kern = r[1][:60]; data.setdefault(kern, []); continue
    if r and r[0] == 'Address': hdr = r; continue
    if kern and hdr and len(r) >= len(hdr) - 2: data[kern].append(r)
for k, v in data.items():
    iS = hdr.index('Source'); iE = hdr.index('Instructions Executed'); iT = hdr.index('# Samples')
    iTh = hdr.index('Thread Instructions Executed')
    tot = sum(int(r[iE] or 0) for r in v); tots = sum(int(r[iT] or 0) for r in v)
    print('=====', k, 'warp-instr', tot, 'thread-instr', sum(int(r[iTh] or 0) for r in v), 'samples', tots, 'n sass', len(v))
    ops = collections.Counter(); ops_s = collections.Counter()
    for r in v:
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[iS]); op = m.group(2) if m else '?'
        op = '.'.join(op.split('.')[:2]) if op.startswith(('F2F', 'I2F', 'F2I', 'LDS', 'STS', 'LDG', 'STG', 'MUFU')) else op.split('.')[0]
        ops[op] += int(r[iE] or 0); ops_s[op] += int(r[iT] or 0)
    for op, c in ops.most_common(30): print(f'  {op:14s} {c:12d} {100*c/max(tot,1):5.1f}%  stall-samples {100*ops_s[op]/max(tots,1):5.1f}%')
