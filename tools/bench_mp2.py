#!/usr/bin/env python
"""MP2 transform numbers: DMMA GEMM throughput and pc_mp2_energy time (run on the GPU box).
Usage: python tools/bench_mp2.py [n_waters_for_mp2]"""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pychem_b200 import _lib, engine, hartree_fock as hf_gpu, mp2 as mp2_gpu, structures as S  # noqa: E402

lib = _lib.load()
out = {}
P = lambda t: ctypes.c_void_p(t.data_ptr())   # noqa: E731
for (M, N, K) in ((4096, 4096, 4096), (40, 192 ** 3, 192), (21, 96 ** 3, 96)):
    A = torch.randn(M, K, dtype=torch.float64, device="cuda")
    B = torch.randn(K, N, dtype=torch.float64, device="cuda")
    C = torch.empty(M, N, dtype=torch.float64, device="cuda")
    _lib.check(lib.pc_dgemm_dmma(0, M, N, K, P(A), P(B), P(C)), mp2=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        _lib.check(lib.pc_dgemm_dmma(0, M, N, K, P(A), P(B), P(C)), mp2=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    out["dgemm_%dx%dx%d" % (M, N, K)] = {"ms": dt * 1e3, "tflops": 2.0 * M * N * K / dt / 1e12}
    del A, B, C
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
for name, mol in (("benzene_631gs", S.Molecule(S.benzene(), "6-31G*")),
                  ("water%d_631gss" % n, S.Molecule(S.water_cluster(n), "6-31G**"))):
    os.environ["PYCHEM_B200_MODE"] = "stored"
    hf_gpu.STORED_LIMIT_BYTES = 1 << 62
    st = {"db": engine.DeviceBasis(mol)}
    db = st["db"]
    db.schwarz()
    G_dev, _ = db.eri_tensor(1e-8, to_host=False)
    hf_gpu._STATE[id(mol)] = {"mode": "stored", "db": db, "G_dev": G_dev, "molecule": mol}
    N = mol.NOrbitals
    rng = np.random.default_rng(1)
    C, _ = np.linalg.qr(rng.uniform(-1, 1, (N, N)))
    E = np.sort(rng.uniform(-2, 2, N)); E[mol.NAlphaElectrons:] += 3.0

    class M_:
        pass
    s = M_(); s.Alpha = M_(); s.Beta = M_()
    s.Alpha.MOs = s.Beta.MOs = C
    s.Alpha.Energies = s.Beta.Energies = E
    mp2_gpu.mp2_sums(mol, s)
    t0 = time.perf_counter()
    e = mp2_gpu.mp2_sums(mol, s)
    dt = time.perf_counter() - t0
    no = mol.NAlphaElectrons
    nv = N - no
    # restricted orbitals (Ca is Cb here): ONE transform serves the three sums (csrc/pc_mp2.cu);
    # unrestricted jobs run three
    flop = 2.0 * (no * N ** 4 + no * nv * N ** 3 + no * nv * no * N ** 2 + no * nv * no * nv * N)
    out["mp2_" + name] = {"N": N, "nocc": no, "seconds": dt, "transforms": 1, "gflop_executed": flop / 1e9,
                          "tflops": flop / dt / 1e12, "sums": e}
    del G_dev
    db.close()
print(json.dumps(out))
