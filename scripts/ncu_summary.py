#!/usr/bin/env python
"""Prints the key metrics of an ncu report (first kernel): used to write the summaries under profiles/."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct",
        "gpu__dram_throughput", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit", "launch__grid_size", "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fp64",
        "sm__pipe_fp64_cycles_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "smsp__issue_active.avg.pct", "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_elapsed.max",
        "sass__inst_executed_local", "lts__throughput.avg.pct", "l1tex__throughput.avg.pct", "smsp__average_warp",
        "smsp__average_warps_issue_stalled", "sm__inst_executed_pipe_alu", "sm__inst_executed_pipe_fma", "sm__inst_executed_pipe_lsu",
        "sm__inst_executed_pipe_xu", "smsp__inst_executed_op_shared", "smsp__inst_executed_op_global", "launch__shared_mem_per_block",
        "lts__t_sectors_op_atom", "lts__t_sectors_op_red", "sm__sass_thread_inst_executed_op_dfma", "sm__sass_thread_inst_executed_op_dmul",
        "sm__sass_thread_inst_executed_op_dadd", "smsp__sass_thread_inst_executed_op_fp64"]
for h, u, v in zip(hdr, units, vals):
    if any(w in h for w in want) and "per_second" not in h:
        print(f"{h} [{u}] = {v}")
