#!/usr/bin/env python3
"""Prints the handful of ncu metrics we track from a .ncu-rep (run where ncu is installed)."""
import csv
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ",
        "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "lts__t_bytes.sum ", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum ", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum ", "launch__registers_per_thread ",
        "launch__occupancy_limit", "smsp__warps_eligible.avg.per_cycle_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard ", "smsp__pcsamp_warps_issue_stalled_wait ",
        "smsp__pcsamp_warps_issue_stalled_short_scoreboard ",
        "smsp__pcsamp_warps_issue_stalled_branch_resolving ",
        "smsp__pcsamp_warps_issue_stalled_selected ", "smsp__pcsamp_warps_issue_stalled_not_selected ",
        "smsp__pcsamp_warps_issue_stalled_no_instructions ",
        "smsp__pcsamp_warps_issue_stalled_lg_throttle ", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle ",
        "smsp__pcsamp_warps_issue_stalled_mio_throttle ", "smsp__pcsamp_warps_issue_stalled_dispatch_stall ",
        "sm__inst_executed_pipe_fp64", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum ",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum ", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum ",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum ",
        # which pipe bounds the kernel
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active ",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active ",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active ",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active ",
        # atomics (the raycast kernel): RED instructions, sectors at L2, share of the RED unit's peak
        "smsp__inst_executed_op_global_red.sum ", "lts__t_sectors_srcunit_tex_op_red.sum ",
        "lts__t_sectors_srcunit_tex_op_red.avg.pct_of_peak_sustained_elapsed ",
        "lts__t_sectors_srcunit_tex_op_red_lookup_hit.sum ",
        "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed ",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed ",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed "]

report = sys.argv[1]
raw = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(raw.splitlines()))
header, units = rows[0], rows[1]
for row in rows[2:]:
    print("==", row[header.index("Kernel Name")][:90])
    for i, name in enumerate(header):
        if any((name + " ").startswith(k) or name == k.strip() for k in KEEP):
            print(f"   {name} [{units[i]}] = {row[i]}")
