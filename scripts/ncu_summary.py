#!/usr/bin/env python
"""Turn an .ncu-rep (one kernel launch, `ncu --set full`) into the 3-column summary kept under profiles/:
metric,unit,value for the first captured launch.  Usage: ncu_summary.py report.ncu-rep out.csv"""
import csv
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ("Kernel Name", "gpu__time", "launch__", "dram__", "sm__inst_executed_pipe", "sm__pipe_fp64", "sm__issue", "smsp__issue",
        "smsp__inst_executed", "smsp__average_warps", "smsp__pcsamp", "l1tex__data_pipe_lsu_wavefronts_mem_shared",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "sm__warps_active", "sm__throughput", "lts__t_bytes", "lts__throughput",
        "device__attribute_max_registers", "SM_A.TriageCompute", "TPC.TriageCompute")
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit", "value"])
    for h, u, v in zip(hdr, units, vals):
        if h.startswith(keep):
            w.writerow([h, u, v])
print("wrote", out)
