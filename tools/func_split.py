#!/usr/bin/env python
"""warp instructions per function (split at unconditional RET/EXIT) of an ncu report; usage: func_split.py rep nchunks"""
import csv, subprocess, sys
rep=sys.argv[1]; nchunk=float(sys.argv[2]) if len(sys.argv)>2 else 1.0
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr=None; seg=[]; cur=0; curs=0; first=None
for r in rows:
    if r and r[0]=='Address': hdr=r; iE=hdr.index('Instructions Executed'); iS=hdr.index('# Samples'); iSrc=hdr.index('Source'); continue
    if hdr and len(r)>iE and r[0].startswith('0x'):
        if first is None: first=r[0]
        n=int(r[iE]) if r[iE].isdigit() else 0
        s=int(r[iS]) if r[iS].isdigit() else 0
        cur+=n; curs+=s
        txt=r[iSrc].strip()
        if (txt.startswith('RET.') or txt.startswith('EXIT')):
            seg.append((first,r[0],cur,curs)); cur=0; curs=0; first=None
seg.append((first,'end',cur,curs))
for a,b,n,s in seg:
    if n/nchunk>0.5: print(a,b, f'{n/nchunk:8.1f} instr/chunk  samples {s}')
