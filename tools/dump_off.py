#!/usr/bin/env python
"""SASS of an ncu report by offset range from the kernel start: dump_off.py rep lo hi [per] (hex offsets; per = divisor for exec counts)"""
import csv, subprocess, sys
rep = sys.argv[1]; lo = int(sys.argv[2], 16); hi = int(sys.argv[3], 16); per = float(sys.argv[4]) if len(sys.argv) > 4 else 16e6
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; base = None
for r in rows:
    if r and r[0] == 'Address': hdr = r; continue
    if hdr and r and r[0].startswith('0x'):
        a = int(r[0], 16)
        if base is None: base = a
        o = a - base
        if lo <= o <= hi:
            g = lambda k: int(r[hdr.index(k)] or 0)
            print(f'{o:6x} {g("Instructions Executed")/per:6.2f} {g("# Samples"):6d} ni{g("stall_no_inst"):5d} w{g("stall_wait"):5d} ss{g("stall_short_sb"):5d} ls{g("stall_long_sb"):5d}  {r[1].strip()[:90]}')
