#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics per kernel + hottest SASS lines (needs ncu on PATH)."""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__grid_size',
        'launch__block_size', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__inst_executed.sum',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'sm__cycles_active.avg', 'sm__cycles_elapsed.max']


def main():
    rep = sys.argv[1]
    top = float(sys.argv[2]) if len(sys.argv) > 2 else 0.015
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('== kernel:', r[hdr.index('Kernel Name')])
        for w in WANT:
            if w in hdr:
                print(f'   {w:70s} {r[hdr.index(w)]:>16s} {units[hdr.index(w)]}')
        for i, h in enumerate(hdr):
            if 'warp_issue_stalled' in h and h.endswith('_per_warp_active.pct'):
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v > 3:
                    print(f'   stall {h.split("stalled_")[1].split("_per_warp")[0]:40s} {v:8.1f} %')
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    k, hd, data = None, None, {}
    for r in csv.reader(src.splitlines()):
        if r and r[0] == 'Kernel Name':
            k = r[1]
            data[k] = []
            hd = None
            continue
        if r and r[0] == 'Address':
            hd = r
            continue
        if k and hd and len(r) > 5:
            data[k].append(r)
    for k, rs in data.items():
        si, so, ie = hd.index('# Samples'), hd.index('Source'), hd.index('Instructions Executed')
        tot = sum(int(r[si]) for r in rs) or 1
        print(f'== hottest SASS lines: {k}  (samples {tot}, warp-instructions {sum(int(r[ie]) for r in rs)})')
        for i, r in enumerate(rs):
            s = int(r[si])
            if s > tot * top:
                print(f'   {i:5d} {100 * s / tot:5.1f}%  exec={r[ie]:>10s}  {r[so].strip()[:100]}')


if __name__ == '__main__':
    main()
