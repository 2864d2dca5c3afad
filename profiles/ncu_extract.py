#!/usr/bin/env python
"""Pull the metrics we judge a sweep kernel by out of `ncu -i X.ncu-rep --page raw --csv`."""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed_pipe_fp64.sum',
        'sm__inst_executed_pipe_fp64.sum', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'lts__t_sectors_op_atom.sum', 'lts__t_sectors_op_red.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.max']


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
    for r in rows[2:]:
        print('---', r[idx['Kernel Name']][:60], 'grid', r[idx['Grid Size']], 'block', r[idx['Block Size']])
        for w in WANT:
            if w in idx:
                print('   %-62s %s %s' % (w, r[idx[w]], units[idx[w]]))
        st = sorted(((float(r[idx[h]].replace(',', '')), h) for h in stall if r[idx[h]]), reverse=True)[:6]
        for v, h in st:
            print('   stall %-56s %.2f' % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v))


if __name__ == '__main__':
    main(sys.argv[1])
