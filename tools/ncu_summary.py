#!/usr/bin/env python
"""Prints the key metrics of every kernel in an .ncu-rep (read on the CPU box): python tools/ncu_summary.py file.ncu-rep"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__m_xbar2l1tex_read_bytes.sum', 'lts__t_sector_hit_rate.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
        'sm__cycles_elapsed.max', 'smsp__cycles_active.avg', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("---- %s  grid=%s block=%s" % (r[idx['Kernel Name']][:90], r[idx['Grid Size']], r[idx['Block Size']]))
        for w in WANT:
            if w in idx:
                print("   %-75s %s %s" % (w, r[idx[w]], units[idx[w]]))


if __name__ == "__main__":
    main(sys.argv[1])
