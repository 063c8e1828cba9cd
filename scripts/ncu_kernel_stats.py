#!/usr/bin/env python
"""Record what bench.py needs from one `ncu --set full` capture of an element kernel in profiles/kernel_stats.json:
DRAM bytes per launch (roofline.traffic), FP64 pipe utilisation and the FP64 thread-instruction count per launch (the
instruction count is a property of the kernel, so bench.py can turn its own live timing into a fraction of the FP64 peak).
The entry is keyed by the kernel LABEL the library reports (jots_ctx_last_kernel) and the vector length; bench.py refuses an
entry whose label differs from the kernel it timed.

Usage: ncu_kernel_stats.py report.ncu-rep "<label>" <n_dof> "<source note>" """
import csv
import json
import os
import subprocess
import sys

rep, label, n_dof, note = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
m = dict(zip(rows[0], rows[2]))
u = dict(zip(rows[0], rows[1]))


def val(k):
    return float(m[k].replace(",", ""))


def nbytes(k):
    return val(k) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[k]]


cyc = val("smsp__cycles_elapsed.avg") if "smsp__cycles_elapsed.avg" in m else val("smsp__cycles_elapsed.sum") / (4 * 148)
fp64 = sum(val("smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % op) for op in ("dfma", "dmul", "dadd")) * cyc
entry = {"kernel": label, "ncu_name": m["Kernel Name"], "n_dof": n_dof,
         "dram_bytes_per_launch": nbytes("dram__bytes_read.sum") + nbytes("dram__bytes_write.sum"),
         "fp64_pipe_active_pct": val("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
         "fp64_thread_instr_per_launch": fp64, "fp64_thread_instr_per_dof": fp64 / n_dof,
         "registers": int(val("launch__registers_per_thread")), "ncu_duration_us": val("gpu__time_duration.sum") * (1e3 if u["gpu__time_duration.sum"] == "ms" else 1.0),
         "source": note}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "kernel_stats.json")
db = json.load(open(path)) if os.path.exists(path) else []
db = [e for e in db if not (e["kernel"] == label and e["n_dof"] == n_dof)] + [entry]
json.dump(db, open(path, "w"), indent=1)
print(json.dumps(entry, indent=1))
