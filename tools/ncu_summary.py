#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics and stall reasons per kernel, optionally the hottest SASS lines.

    python tools/ncu_summary.py report.ncu-rep [--hot N]
"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__average_warp_latency_per_inst_issued.ratio',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor', 'launch__grid_size', 'launch__block_size']


def page(rep, name, extra=()):
    out = subprocess.run(['ncu', '-i', rep, '--page', name, '--csv', *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    hot = int(sys.argv[sys.argv.index('--hot') + 1]) if '--hot' in sys.argv else 0
    rows = page(rep, 'raw')
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print('==', d.get('Kernel Name', '?')[:80])
        for k in KEYS:
            if k in d:
                print(f'  {k:70s} {d[k]} {units[hdr.index(k)]}')
        st = {k: float(v.replace(',', '')) for k, v in d.items() if 'smsp__average_warps_issue_stalled' in k and k.endswith('_per_issue_active.ratio')}
        for k, v in sorted(st.items(), key=lambda x: -x[1])[:7]:
            print('    stall', k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), round(v, 3))
    if hot:
        rows = page(rep, 'source', ('--print-source', 'sass'))
        hdr = rows[1]
        isrc, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
        data = [r for r in rows[2:] if len(r) > max(isrc, isamp, iex)]      # (a report with several kernels repeats the two header rows)
        tot = sum(int(r[isamp]) for r in data if r[isamp].isdigit())
        top = sorted(((int(r[isamp]), i) for i, r in enumerate(data) if r[isamp].isdigit()), reverse=True)[:hot]
        print('total samples', tot)
        for smp, i in sorted(top, key=lambda x: x[1]):
            print(f'  {i:5d} {smp:7d} {100.0 * smp / tot:5.1f}%  exec {data[i][iex]:>10s}  {data[i][isrc].strip()[:90]}')


if __name__ == '__main__':
    main()
