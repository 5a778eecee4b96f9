#!/usr/bin/env python
"""Prints the metrics we track from an .ncu-rep (raw page) as `name unit value` lines.
Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [extra-substring ...]"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread ",
    "launch__occupancy_limit", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum ", "sm__inst_executed.avg.per_cycle_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
    "smsp__average_warp_latency_issue_stalled", "smsp__average_warps_issue_stalled",
    "lts__t_bytes.sum ", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_active.avg ", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__pcsamp_warps_issue_stalled",
]


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print("== kernel:", name[:100])
        for h, u, v in zip(hdr, units, r):
            hh = h + " "
            if any(k in hh for k in KEYS) or any(e in h for e in extra):
                if v not in ("", "0", "0.000000"):
                    print("%-90s %-14s %s" % (h, u, v))


if __name__ == "__main__":
    main()
