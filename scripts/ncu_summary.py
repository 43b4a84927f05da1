#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into the CSV columns the roofline
discussion needs.  usage: ncu_summary.py in.ncu-rep out.csv"""
import csv
import subprocess
import sys

KEEP = ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput",
        "sm__throughput.avg.pct", "sm__warps_active.avg.pct", "launch__registers", "launch__grid_size", "launch__block_size",
        "launch__waves", "launch__occupancy_limit", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_bytes.sum", "sm__cycles_elapsed.max",
        "smsp__average_warp", "smsp__warp_issue_stalled", "sm__inst_executed_pipe", "launch__shared_mem")
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
keep = [i for i, h in enumerate(hdr) if any(s in h for s in KEEP)]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    for r in rows:
        w.writerow([r[i] for i in keep])
