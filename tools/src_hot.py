#!/usr/bin/env python
"""Hot CUDA source lines of a kernel in an ncu report: warp-instructions executed and stall samples per line.
usage: src_hot.py report.ncu-rep [N]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = None; hdr = None; lines = []
for r in rows:
    if not r: continue
    if r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr and r[0] not in ('', '-') and r[0].isdigit():
        iE = hdr.index('Instructions Executed'); iS = hdr.index('# Samples')
        lines.append((fname, int(r[0]), r[1].strip(), int(r[iE]) if r[iE].isdigit() else 0, int(r[iS]) if r[iS].isdigit() else 0))
tot = sum(l[3] for l in lines); tots = sum(l[4] for l in lines)
print(f'total warp-instr {tot}  samples {tots}')
for l in sorted(lines, key=lambda l: -l[3])[:N]:
    print(f'{100*l[3]/tot:5.1f}% instr {100*l[4]/max(tots,1):5.1f}% stall  {l[0]}:{l[1]}  {l[2][:110]}')

# region summary for pskd_fused.cu: pass --regions a:b,c:d,...
if '--regions' in sys.argv:
    regs = [tuple(map(int, r.split(':'))) for r in sys.argv[sys.argv.index('--regions') + 1].split(',')]
    nsym = float(sys.argv[sys.argv.index('--symbols') + 1]) if '--symbols' in sys.argv else 1.0
    byfile = collections.Counter()
    for l in lines:
        if l[0] != 'pskd_fused.cu': byfile[l[0]] += l[3]
    for a, b in regs:
        t = sum(l[3] for l in lines if l[0] == 'pskd_fused.cu' and a <= l[1] <= b)
        print(f'pskd_fused.cu:{a}-{b}: {100*t/tot:5.1f}%  {32*t/nsym:7.1f} thread-instr/symbol')
    for f, t in byfile.most_common(): print(f'{f}: {100*t/tot:5.1f}%  {32*t/nsym:7.1f} thread-instr/symbol')
