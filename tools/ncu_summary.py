#!/usr/bin/env python3
"""Text summary of an ncu --set full report: one block per kernel launch with the metrics the
roofline discussion uses.  usage: tools/ncu_summary.py <report.ncu-rep> [header text]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "launch__grid_size", "launch__block_size", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__waves_per_multiprocessor", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
idx = [hdr.index(w) if w in hdr else None for w in want]
if len(sys.argv) > 2:
    print(sys.argv[2])
for r in rows[2:]:
    print("---")
    for w, i in zip(want, idx):
        if i is not None:
            print("  %-62s %s %s" % (w, r[i][:90], units[i]))
