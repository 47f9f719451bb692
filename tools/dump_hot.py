#!/usr/bin/env python
"""hot SASS of an ncu report in address order: dump_hot.py rep nchunks lo_addr_suffix hi_addr_suffix [minfrac]"""
import csv, subprocess, collections, sys
rep=sys.argv[1]; nchunk=float(sys.argv[2]); lo=int(sys.argv[3],16); hi=int(sys.argv[4],16); minf=float(sys.argv[5]) if len(sys.argv)>5 else 0.2
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname=None; hdr=None; cur=None; seen=collections.OrderedDict()
def I(x):
    try: return int(x)
    except: return 0
for r in rows:
    if not r: continue
    if r[0]=='File Path': fname=r[1].split('/')[-1]; continue
    if r[0]=='Function Name': continue
    if r[0]=='Line No': hdr=r; continue
    if hdr is None: continue
    if r[0].isdigit(): cur=(fname,int(r[0])); continue
    if r[0]=='' and cur and r[2].startswith('0x'):
        a=int(r[2],16)
        if not (lo <= (a & 0xfffff) <= hi): continue
        e=seen.setdefault(a,[r[3].strip(),I(r[hdr.index('Instructions Executed')]),I(r[hdr.index('# Samples')]),[]])
        e[3].append(f'{cur[0].replace("pskd_","").replace(".cuh","").replace(".cu","").replace("sm_30_intrinsics.hpp","i30").replace("sm_32_intrinsics.hpp","i32")}:{cur[1]}')
for a in sorted(seen):
    e=seen[a]
    if e[1]/nchunk>=minf:
        print(f'{a&0xfffff:05x} {e[1]/nchunk:5.2f} {e[2]:5d}  {e[0][:58]:58s} {"<".join(e[3][:3])}')
