#!/usr/bin/env python
"""Per-phase breakdown of one captured launch (`ncu --set full --import-source on`): the SASS of the kernel is cut at its
BAR.SYNC instructions and, per segment, the share of warp-stall samples, instructions, FP64 instructions, shared-memory
wavefronts and the dominant stall reasons are printed (barrier stalls are attributed to the instruction after the barrier).
Usage: ncu_phases.py report.ncu-rep"""
import collections
import csv
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
print(rows[0][1][:200])
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]


def f(r, k):
    try:
        return float(r[ix[k]])
    except (ValueError, KeyError, IndexError):
        return 0.0


tot_s = sum(f(r, "# Samples") for r in data)
tot_i = sum(f(r, "Instructions Executed") for r in data)
print("SASS instructions %d, samples %d, warp instructions %.1f M" % (len(data), tot_s, tot_i / 1e6))
bars = [n for n, r in enumerate(data) if "BAR.SYNC" in r[ix["Source"]] or "BAR.ARV" in r[ix["Source"]]]
bounds = [0] + [b + 1 for b in bars] + [len(data)]
for s in range(len(bounds) - 1):
    seg = data[bounds[s]:bounds[s + 1]]
    if not seg:
        continue
    smp = sum(f(r, "# Samples") for r in seg)
    ins = sum(f(r, "Instructions Executed") for r in seg)
    fp = sum(f(r, "Instructions Executed") for r in seg if r[ix["Source"]].split()[-0 if not r[ix["Source"]].strip().startswith("@") else 1].startswith(("DFMA", "DMUL", "DADD")))
    exc = sum(f(r, "L1 Wavefronts Shared Excessive") for r in seg)
    wf = sum(f(r, "L1 Wavefronts Shared") for r in seg)
    d = {k: sum(f(r, k) for r in seg) for k in stalls}
    top = sorted(d.items(), key=lambda kv: -kv[1])[:6]
    print("seg %d rows %4d-%4d: samples %5.1f%% instr %5.1f%% fp64 %6.1fM smem wavefronts %6.1fM (excess %5.1fM) | %s"
          % (s, bounds[s], bounds[s + 1], 100 * smp / tot_s, 100 * ins / tot_i, fp / 1e6, wf / 1e6, exc / 1e6,
             " ".join("%s=%.1f%%" % (k[6:], 100 * v / tot_s) for k, v in top)))
print("instructions with the most excessive shared-memory wavefronts:")
for r in sorted(data, key=lambda r: -f(r, "L1 Wavefronts Shared Excessive"))[:8]:
    print("  %-64s excess %.2fM of %.2fM" % (r[ix["Source"]].strip()[:64], f(r, "L1 Wavefronts Shared Excessive") / 1e6, f(r, "L1 Wavefronts Shared") / 1e6))
