#!/usr/bin/env python
"""Extract the small subset of Gaussian basis-set DATA the benchmarks/tests need.

The reference keeps 62 basis sets in one 2 MB Python literal (Data/basis.py).  The GPU box has
no /root/reference, so the (public, EMSL-derived) numbers for the handful of sets/elements the
BASELINE configs use are written once to pychem_b200/data/basis_subset.json in the reference's
own record format ``[l, [exponent, coefficient], ...]`` (Util/structures.py:836-843).

Usage: python oracle/extract_basis.py   (needs oracle/_ref built, i.e. /root/reference present)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref", "pychem_py3"))
from Data import basis  # noqa: E402

SETS = ["STO3G", "321G", "631G", "631GS", "631GSS", "6311G", "6311GSS", "CCPVDZ", "CCPVTZ"]
ELEMENTS = ["H", "HE", "LI", "BE", "B", "C", "N", "O", "F", "NE"]

out = {}
for s in SETS:
    out[s] = {}
    for e in ELEMENTS:
        if e in basis.get[s]:
            out[s][e] = basis.get[s][e]
path = os.path.join(ROOT, "pychem_b200", "data", "basis_subset.json")
with open(path, "w") as fh:
    json.dump(out, fh, separators=(",", ":"))
print("wrote", path, os.path.getsize(path), "bytes")
