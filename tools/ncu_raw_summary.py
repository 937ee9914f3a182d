#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` dump: per kernel launch the speed-of-light numbers that
matter for this path (FP64 pipe, issue, L1/L2/DRAM throughput, occupancy, DRAM bytes).
Usage: python tools/ncu_raw_summary.py raw.csv [top]"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def f(r, k):
    try:
        return float(r[ix[k]].replace(",", ""))
    except Exception:
        return float("nan")


M = [("dur_us", "gpu__time_duration.sum"), ("regs", "launch__registers_per_thread"),
     ("occ%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
     ("fp64%", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
     ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
     ("l1tex%", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
     ("lts%", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
     ("dram%", "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
     ("l1hit%", "l1tex__t_sector_hit_rate.pct"), ("l2hit%", "lts__t_sector_hit_rate.pct"),
     ("dramR_MB", "dram__bytes_read.sum"), ("dramW_MB", "dram__bytes_write.sum")]
unit_t = units[ix["gpu__time_duration.sum"]]
scale_t = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit_t, 1.0)
out = []
for r in data:
    m = re.search(r"(eri_\w+_kernel|jk_\w+_kernel)<?\(?(?:int\))?(\d)?", r[ix["Kernel Name"]])
    name = (m.group(1) + ("<%s>" % m.group(2) if m.group(2) else "")) if m else r[ix["Kernel Name"]][:40]
    vals = {}
    for short, key in M:
        v = f(r, key) if key in ix else float("nan")
        if short == "dur_us":
            v *= scale_t
        if short.endswith("_MB"):
            u = units[ix[key]] if key in ix else ""
            v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
        vals[short] = v
    out.append((vals["dur_us"], name, r[ix["Grid Size"]] if "Grid Size" in ix else "", vals))
out.sort(key=lambda x: -x[0])
print("launches captured:", len(out), " total us: %.1f" % sum(o[0] for o in out))
print("%-26s %-14s " % ("kernel", "grid") + " ".join("%8s" % s for s, _ in M))
for dur, name, grid, vals in out[:top]:
    print("%-26s %-14s " % (name, grid) + " ".join("%8.1f" % vals[s] for s, _ in M))
