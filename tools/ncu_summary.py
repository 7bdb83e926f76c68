#!/usr/bin/env python
"""Summarise an `ncu --set full` report: one block of metrics per profiled launch (raw page) -- the text committed under profiles/.
usage: tools/ncu_summary.py report.ncu-rep [> profiles/ncu_xxx.txt]"""
import csv, io, subprocess, sys

WANT = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers'] + [
        'smsp__average_warps_issue_stalled_%s_per_issue_active.ratio' % s for s in
        ('long_scoreboard', 'short_scoreboard', 'barrier', 'math_pipe_throttle', 'wait', 'mio_throttle', 'lg_throttle', 'membar',
         'dispatch_stall', 'not_selected', 'no_instruction', 'branch_resolving')]


def main():
    out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    kn = hdr.index('Kernel Name')
    for r in rows[2:]:
        print('== ' + r[kn])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print('   %-86s %s %s' % (w, r[i], units[i]))


if __name__ == '__main__':
    main()
