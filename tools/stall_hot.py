#!/usr/bin/env python
"""SASS instructions of a kernel with the most stall samples (ncu source page).
usage: stall_hot.py report.ncu-rep [N] [stall_column]"""
import csv, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
col = sys.argv[3] if len(sys.argv) > 3 else '# Samples'
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = None; hdr = None; cur = None; seen = {}
def I(x):
    try: return int(x)
    except ValueError: return 0
for r in rows:
    if not r: continue
    if r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None: continue
    if r[0].isdigit(): cur = (fname, int(r[0])); continue
    if r[0] == '' and cur:
        addr = r[2]
        # the same SASS row shows up under every inlined file/line: keep the innermost-first occurrence, list all lines
        ent = seen.setdefault(addr, [r[3].strip(), I(r[hdr.index(col)]), I(r[hdr.index('Instructions Executed')]), []])
        ent[3].append(f'{cur[0]}:{cur[1]}')
tot = sum(e[1] for e in seen.values()); toti = sum(e[2] for e in seen.values())
print(f'{col}: total {tot}; warp-instr {toti}')
for e in sorted(seen.values(), key=lambda e: -e[1])[:N]:
    print(f'{100*e[1]/max(tot,1):5.1f}%  x{e[2]:>10d}  {e[0][:60]:60s} {" < ".join(e[3][:3])}')
