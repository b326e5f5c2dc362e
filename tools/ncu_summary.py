#!/usr/bin/env python3
"""Condense an .ncu-rep into the two text artifacts kept under profiles/:
<name>_ncu_details.csv (ncu --page details) and <name>_ncu_summary.txt (selected lines + raw counters).
usage: python tools/ncu_summary.py gpurun_out/<name>.ncu-rep [out_dir]"""
import csv, subprocess, sys
from pathlib import Path

rep = Path(sys.argv[1])
out_dir = Path(sys.argv[2] if len(sys.argv) > 2 else "profiles")
name = rep.stem
details = subprocess.run(["ncu", "-i", str(rep), "--page", "details", "--csv"], capture_output=True, text=True).stdout
(out_dir / f"{name}_ncu_details.csv").write_text(details)
keep = [("GPU Speed Of Light Throughput", "Memory Throughput"), ("GPU Speed Of Light Throughput", "DRAM Throughput"),
        ("GPU Speed Of Light Throughput", "Duration"), ("GPU Speed Of Light Throughput", "SM Frequency"),
        ("Compute Workload Analysis", "Executed Ipc Active"),
        ("Compute Workload Analysis", "Issue Slots Busy"), ("Memory Workload Analysis", "Memory Throughput"),
        ("Memory Workload Analysis", "Mem Busy"), ("Memory Workload Analysis", "L2 Hit Rate"),
        ("Scheduler Statistics", "No Eligible"), ("Scheduler Statistics", "Eligible Warps Per Scheduler"),
        ("Launch Statistics", "Registers Per Thread"), ("Launch Statistics", "Block Size"), ("Launch Statistics", "Grid Size"),
        ("Launch Statistics", "Dynamic Shared Memory Per Block"),
        ("Occupancy", "Theoretical Occupancy"), ("Occupancy", "Achieved Occupancy")]
lines = []
rows = list(csv.reader(details.splitlines()))
hdr = rows[0]
si, mi, ui, vi, ki = (hdr.index(x) for x in ("Section Name", "Metric Name", "Metric Unit", "Metric Value", "Kernel Name"))
lines.append(f"# {rows[1][ki]}")
for r in rows[1:]:
    if (r[si], r[mi]) in keep:
        lines.append(f"{r[si]:35s} {r[mi]:40s} {r[vi]:>14s} {r[ui]}")
raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
want = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]
for i, h in enumerate(rr[0]):
    if h in want or (h.startswith("sm__inst_executed_pipe_") and h.endswith(".avg.pct_of_peak_sustained_active")
                     and float(rr[2][i] or 0) > 0.5) or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")
                                                          and float(rr[2][i] or 0) > 0.05):
        lines.append(f"{h:95s} {rr[2][i]} {rr[1][i]}")
(out_dir / f"{name}_ncu_summary.txt").write_text("\n".join(lines) + "\n")
print("\n".join(lines))
