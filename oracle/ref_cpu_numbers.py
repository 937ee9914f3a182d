#!/usr/bin/env python
"""CPU-baseline numbers of the REFERENCE itself on this host (SURVEY 8(d) "CPU baseline beside
it"): evaluate_2e_ints wall time and make_coulomb_exchange_matrices ms/call for the small
BASELINE configs, 1 core (the reference is single-threaded).  Test/measurement infrastructure:
drives oracle/_ref.  Usage: python oracle/ref_cpu_numbers.py > profiles/...json"""
import json
import os
import platform
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_driver  # noqa: E402
from pychem_b200 import structures as S  # noqa: E402

ns = ref_driver.modules()
H2 = [["H", 1.0, 0.0, 0.0, 0.0], ["H", 1.0, 0.0, 0.0, 1.0]]
LIH = [["Li", 3.0, 0.0, 0.0, 0.0], ["H", 1.0, 2.2, 0.0, 0.0]]
out = {"host": platform.processor() or platform.machine(), "cores_used": 1, "cases": {}}
try:
    out["cpu_model"] = [l.split(":")[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
except Exception:
    pass


class St:
    class M:
        pass

    def __init__(self, D):
        self.Total, self.Alpha, self.Beta = St.M(), St.M(), St.M()
        self.Total.Density, self.Alpha.Density, self.Beta.Density = 2 * D, D, D


for name, coords, basis in (("H2 6-311G", H2, "6-311G"), ("LiH 6-31G", LIH, "6-31G"),
                            ("H2O 6-31G**", S.H2O_MONOMER, "6-31G**"),
                            ("(H2O)2 6-31G**", S.water_cluster(2), "6-31G**")):
    mol, _ = ref_driver.build_molecule(coords, basis)
    t0 = time.perf_counter()
    with np.errstate(all="ignore"):
        ns.hartree_fock.evaluate_2e_ints(mol)
    t_eri = time.perf_counter() - t0
    N = mol.NOrbitals
    rng = np.random.default_rng(1)
    X = rng.uniform(-1, 1, (N, N))
    st = St(0.5 * (X + X.T))
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        ns.hartree_fock.make_coulomb_exchange_matrices(mol, st)
    t_jk = (time.perf_counter() - t0) / reps
    n = mol.NCgtf
    npair = n * (n + 1) // 2
    out["cases"][name] = {"N": N, "shells": n, "unique_quartets": npair * (npair + 1) // 2,
                          "evaluate_2e_ints_s": t_eri, "tensor_elements_per_s": N ** 4 / t_eri,
                          "make_coulomb_exchange_ms": t_jk * 1e3}
print(json.dumps(out, indent=1))
