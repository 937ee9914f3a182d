#!/usr/bin/env python
"""Registers / stack / shared memory of every class kernel (modes JK_RHF = 2 and NULL = 5) from
cuobjdump --dump-resource-usage of a built library.  Usage: python tools/resource_usage.py [lib.so]"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "pychem_b200", "libpychem_b200.so")
out = subprocess.run(["cuobjdump", "--dump-resource-usage", lib], capture_output=True, text=True).stdout
rows = {}
name = None
for line in out.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        name = m.group(1)
        continue
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
    if m and name:
        k = re.search(r"\d+eri_([A-Za-z]{4})_kernelILi(\d+)EE", name)
        if k and k.group(2) in ("2", "5"):
            rows.setdefault(k.group(1), {})[k.group(2)] = tuple(int(x) for x in m.groups())
        name = None
print("# class   JK_RHF: regs stack smem | NULL: regs stack smem")
for c in sorted(rows):
    a, b = rows[c].get("2", (0, 0, 0, 0)), rows[c].get("5", (0, 0, 0, 0))
    print("%-6s %5d %6d %6d | %5d %6d %6d" % (c, a[0], a[1], a[2], b[0], b[1], b[2]))
