#!/usr/bin/env python
"""key metrics of the first kernel in an ncu report: ncu_key.py report.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u, r = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'sm__warps_active.avg.per_cycle_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct', 'launch__shared_mem_per_block_dynamic',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'sm__cycles_elapsed.avg.per_second']
for w in want:
    if w in h: print(f'{w:75s} {r[h.index(w)]:>20s} {u[h.index(w)]}')
tot = 0.0; st = []
for i, n in enumerate(h):
    if n.startswith('smsp__average_warps_issue_stalled_') and n.endswith('_per_issue_active.ratio') and 'not_issued' not in n:
        try: st.append((float(r[i]), n[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]))
        except ValueError: pass
for v, n in sorted(st, reverse=True)[:12]: print(f'  stall {n:30s} {v:6.2f} warp-cycles per issue')
