#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of numbers we track per kernel.
usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep [--stalls]"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed_op_shared_atom.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_fma.sum", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
        "smsp__warp_issue_stalled_not_selected_per_warp_active.pct", "smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct",
        "smsp__warp_issue_stalled_membar_per_warp_active.pct", "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct"]
ki = hdr.index("Kernel Name")
for r in rows[2:]:
    print("==", r[ki][:70])
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"   {w:75s} {r[i]:>16s} {units[i]}")
