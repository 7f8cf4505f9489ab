#!/usr/bin/env python
"""Developer aid: condense an .ncu-rep into the markdown table kept under profiles/."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor', 'smsp__inst_executed.sum',
        'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.max', 'launch__grid_size',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']


def main(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    n = len(rows) - 2
    print("| metric | unit | " + " | ".join("launch %d" % (i + 1) for i in range(n)) + " |")
    print("|---|---|" + "---|" * n)
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print("| %s | %s | %s |" % (k, rows[1][i], " | ".join(r[i] for r in rows[2:])))


if __name__ == "__main__":
    main(sys.argv[1])
