#!/usr/bin/env python
"""Print the metrics we track from an `ncu --page raw --csv` export (one block per kernel launch)."""
import csv
import sys

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum']
STALL = 'smsp__average_warps_issue_stalled_'
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print('=====')
    for w in WANT:
        if w in idx:
            print("%-72s %s %s" % (w, r[idx[w]], units[idx[w]]))
    st = [(float(r[i].replace(',', '')), h[len(STALL):-len('_per_issue_active.ratio')]) for h, i in idx.items()
          if h.startswith(STALL) and h.endswith('_per_issue_active.ratio') and r[i]]
    st.sort(reverse=True)
    print("stalls (warps per issue):", ", ".join("%s=%.2f" % (n, v) for v, n in st[:7]))
