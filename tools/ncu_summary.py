#!/usr/bin/env python
"""Condense an ncu report (.ncu-rep, read here without a GPU) into the handful of counters the
roofline claims rest on.  Usage:  python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_x.txt
"""

import csv
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "launch__shared_mem_per_block_dynamic",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "dram__bytes_read.sum.per_second",
    "dram__bytes_write.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum",
    "smsp__average_warp_latency_issue_stalled",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    lines = [ln for ln in out.splitlines() if ln.startswith('"')]
    rows = list(csv.reader(lines))
    hdr, units = rows[0], rows[1]
    print(f"# {rep}")
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")]
        print(f"kernel: {name}")
        print(f"  grid {vals[hdr.index('Grid Size')]} block {vals[hdr.index('Block Size')]}")
        for i, h in enumerate(hdr):
            if h in KEEP:
                print(f"  {h:85s} {vals[i]:>18s} {units[i]}")


if __name__ == "__main__":
    main()
