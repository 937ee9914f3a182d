#!/usr/bin/env python
"""Summarise a SASS source page dumped by tools/ncu_source_dump.py: stall breakdown and the
instructions carrying the most stall samples.  Usage: python tools/ncu_top_stalls.py FILE.csv.gz [top]"""
import csv
import gzip
import io
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
text = gzip.open(path, "rt").read() if path.endswith(".gz") else open(path).read()
lines = text.splitlines()
print(lines[0])
rows = list(csv.reader(io.StringIO("\n".join(l for l in lines if not l.startswith("#")))))
h = next(k for k, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
data = [r for r in rows[h + 1:] if len(r) >= len(hdr) - 2]
ci = {c: i for i, c in enumerate(hdr)}
cs, csrc, cex = ci["Warp Stall Sampling (All Samples)"], ci["Source"], ci["Instructions Executed"]


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return 0.0


tot = sum(num(r[cs]) for r in data) or 1.0
stall_cols = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
agg = {}
for r in data:
    for i in stall_cols:
        agg[hdr[i]] = agg.get(hdr[i], 0.0) + num(r[i])
print("instructions: %d   samples: %d   warp-instructions executed: %d" % (len(data), tot, sum(num(r[cex]) for r in data)))
print("stall breakdown (%% of samples):", ", ".join("%s=%.1f" % (k.replace("stall_", ""), 100 * v / tot)
                                                   for v, k in sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:9]))
# instruction mix
mix = {}
for r in data:
    op = r[csrc].split()[0] if r[csrc].split() else "?"
    if op.startswith("@"):
        op = r[csrc].split()[1]
    op = op.split(".")[0]
    mix[op] = mix.get(op, 0) + num(r[cex])
tex = sum(mix.values()) or 1
print("executed mix:", ", ".join("%s=%.1f%%" % (k, 100 * v / tex) for v, k in sorted(((v, k) for k, v in mix.items()), reverse=True)[:14]))
order = sorted(range(len(data)), key=lambda k: -num(data[k][cs]))[:top]
for k in sorted(order):
    r = data[k]
    why = sorted(((num(r[i]), hdr[i]) for i in stall_cols if num(r[i])), reverse=True)[:2]
    print("%5d %6.2f%% ex=%-9s %-70s %s" % (k, 100 * num(r[cs]) / tot, r[cex], r[csrc].strip()[:70],
                                           " ".join("%s=%d" % (b.replace("stall_", ""), a) for a, b in why)))
