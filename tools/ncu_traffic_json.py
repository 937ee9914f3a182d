#!/usr/bin/env python
"""profiles/ncu_traffic.json from `ncu --set full ... --page raw --csv` dumps: per kernel the DRAM
bytes (dram__bytes_read.sum + dram__bytes_write.sum) of its launch, the figure bench.py reports as
roofline.traffic.  Usage: python tools/ncu_traffic_json.py out.json source-label raw1.csv [raw2.csv ...]"""
import csv
import json
import re
import sys

out_path, label, files = sys.argv[1], sys.argv[2], sys.argv[3:]
MODES = {"0": "BLOCKS", "1": "TENSOR", "2": "JK_RHF", "3": "JK_UHF", "4": "JK_GEN", "5": "NULL"}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
tab = {}
for f in files:
    rows = list(csv.reader(open(f)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        m = re.search(r"(eri_\w+_kernel)<(?:\(int\))?(\d)>", r[ix["Kernel Name"]]) or re.search(r"(jk_\w+_kernel)", r[ix["Kernel Name"]])
        if not m:
            continue
        name = m.group(1) + ("<%s>" % MODES.get(m.group(2), m.group(2)) if m.lastindex and m.lastindex > 1 else "")
        tot = 0.0
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(r[ix[key]].replace(",", "")) * SCALE.get(units[ix[key]], 1.0)
        dur = float(r[ix["gpu__time_duration.sum"]].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(units[ix["gpu__time_duration.sum"]], 1.0)
        e = tab.get(name)
        if e is None or dur > e["duration_ms_under_ncu"]:
            tab[name] = {"dram_bytes": tot, "duration_ms_under_ncu": dur, "source": "%s (%s)" % (label, f.split("/")[-1])}
json.dump(tab, open(out_path, "w"), indent=1, sort_keys=True)
print(json.dumps(tab, indent=1, sort_keys=True))
