#!/usr/bin/env python
"""Run on the GPU box after an `ncu -o REP` capture: for each requested kernel-name substring pick
the longest captured launch and dump its SASS source page (per-instruction samples/stalls) to
gpurun_out/src_<name>.csv.gz.   Usage: python tools/ncu_source_dump.py REP name1 name2 ..."""
import csv
import gzip
import io
import subprocess
import sys

rep = sys.argv[1]
names = sys.argv[2:]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", "gpu__time_duration.sum"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
i_n, i_t = hdr.index("Kernel Name"), hdr.index("gpu__time_duration.sum")
for nm in names:
    best = None
    count = 0
    for r in rows[2:]:
        if nm in r[i_n]:
            count += 1
            t = float(r[i_t].replace(",", ""))
            if best is None or t > best[0]:
                best = (t, count, r[i_n])
    if best is None:
        print("no launch matching", nm)
        continue
    t, inv, full = best
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass",
                          "--kernel-id", "::regex:%s:%d" % (nm, inv)], capture_output=True, text=True)
    safe = "".join(ch if ch.isalnum() else "_" for ch in nm).strip("_")
    with gzip.open("gpurun_out/src_%s.csv.gz" % safe, "wt") as fh:
        fh.write("# %s invocation %d duration %s\n" % (full, inv, t))
        fh.write(out.stdout)
    print(nm, "invocation", inv, "duration", t, "bytes", len(out.stdout), out.stderr[-200:])
