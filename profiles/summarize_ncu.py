#!/usr/bin/env python
"""Condense ncu output into the small text summaries committed under profiles/.

    python profiles/summarize_ncu.py launches gpurun_out/launches.csv         # per-kernel time shares of a launch list
    python profiles/summarize_ncu.py full gpurun_out/prof.ncu-rep             # key metrics of a --set full capture
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed.sum',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'lts__t_sectors.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__average_warps_issue_stalled', 'sm__inst_executed_pipe_lsu', 'smsp__inst_executed_op_global_atom', 'lts__t_sectors_op_atom.sum',
        'lts__t_sectors_op_red.sum', 'derived__smsp__inst_executed_per_warp']


def launches(path):
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith('==')]
    hdr = rows[0]
    iname, ival, imet = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name')
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for r in rows[1:]:
        if r[imet] != 'gpu__time_duration.sum':
            continue
        name = r[iname].split('(')[0].replace('void ', '')
        tot[name] += float(r[ival].replace(',', ''))
        cnt[name] += 1
    unit = rows[1][hdr.index('Metric Unit')]
    allt = sum(tot.values())
    print('# ncu launch list: %s (cold-cache, serialised; compare SHARES)' % path)
    print('%-60s %8s %14s %7s' % ('kernel', 'launches', 'time [%s]' % unit, 'share'))
    for k in sorted(tot, key=lambda k: -tot[k]):
        print('%-60s %8d %14.1f %6.1f%%' % (k[:60], cnt[k], tot[k], 100 * tot[k] / allt))


def full(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print('# ncu --set full: %s' % path)
    for r in rows[2:]:
        print('## kernel: %s' % r[hdr.index('Kernel Name')][:100])
        for h, u, v in zip(hdr, units, r):
            if any(k in h for k in KEYS):
                print('%-90s %16s %s' % (h, v, u))


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2])
