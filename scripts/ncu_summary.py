#!/usr/bin/env python
"""Summarise one .ncu-rep (raw page) into a text file for profiles/."""
import csv
import subprocess
import sys

rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0        # index of the launch inside the report
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2 + which]
title += "\nkernel: " + vals[hdr.index("Kernel Name")]
keep = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct",
        "dram__throughput.avg.pct", "lts__t_bytes.sum", "lts__t_sector_hit_rate", "l1tex__t_sector_hit_rate", "l1tex__t_bytes.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit", "sm__warps_active.avg.pct", "sm__throughput.avg.pct", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_", "smsp__pipe_", "sm__inst_executed_pipe_", "smsp__inst_executed_pipe_", "sm__cycles_elapsed.avg",
        "smsp__average_warps_issue_stalled", "l1tex__data_bank_conflicts", "smsp__sass_average_data_bytes_per_sector",
        "l1tex__average_t_sectors_per_request", "sm__sass_inst_executed_op_shared", "launch__shared_mem_per_block")
with open(out, "w") as f:
    f.write(title + "\n\n")
    for h, u, v in sorted(zip(hdr, units, vals)):
        if any(h.startswith(k) for k in keep):
            if h.startswith("smsp__average_warps_issue_stalled") and not h.endswith("per_issue_active.ratio"):
                continue
            f.write("%-95s %s %s\n" % (h, v, u))
print(open(out).read())
