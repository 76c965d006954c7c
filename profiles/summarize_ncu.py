"""Summarise an `ncu --page raw --csv` export into the per-kernel text blocks kept in profiles/.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv
    python profiles/summarize_ncu.py raw.csv > profiles/rNN_..._summary.txt
"""
import csv
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
]
STALL = "smsp__average_warps_issue_stalled_"


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("=====")
        print("%-72s %s" % ("Kernel Name", r[ki]))
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("%-72s %s %s" % (k, r[i], units[i]))
        st = []
        for i, h in enumerate(hdr):
            if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and r[i]:
                st.append((float(r[i]), h[len(STALL):-len("_per_issue_active.ratio")]))
        st.sort(reverse=True)
        print("stalls (warps per issue): " + ", ".join("%s=%.2f" % (n, v) for v, n in st[:7]))


if __name__ == "__main__":
    main(sys.argv[1])
