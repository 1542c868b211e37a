#!/usr/bin/env python
"""Condense an .ncu-rep into the few numbers the roofline discussion needs.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [> profiles/summary.txt]"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "dmma_pipe%"),
    ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor_active_elapsed%"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__t_sector_hit_rate.pct", "l2hit%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
    ("smsp__inst_executed.sum", "warp_insts"),
]
STALLS = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
STALL_NAMES = ["barrier", "long_scoreboard", "short_scoreboard", "math_pipe_throttle", "wait", "not_selected",
               "mio_throttle", "lg_throttle", "branch_resolving", "no_instruction", "dispatch_stall", "membar"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    name_i = col.get("Kernel Name")
    for d in data:
        print("== %s  (id %s)" % (d[name_i][:60], d[col["ID"]]))
        for k, short in KEYS:
            hit = [h for h in hdr if h.endswith(k)]
            if hit:
                i = col[hit[0]]
                print("   %-26s %14s %s" % (short, d[i][:14], units[i]))
        st = []
        for s in STALL_NAMES:
            h = STALLS % s
            if h in col:
                st.append((float(d[col[h]] or 0), s))
        st.sort(reverse=True)
        print("   stalls/issue: " + ", ".join("%s %.2f" % (s, v) for v, s in st[:6]))


if __name__ == "__main__":
    main()
