#!/usr/bin/env python
"""Stored-tensor mode numbers: dense-tensor build time and the HBM roofline of
jk_stored_kernel (8*N^4 bytes streamed once per Fock build) at a given cluster size.
Usage: python tools/bench_stored.py [n_waters]      (run on the GPU box)"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pychem_b200 import engine, structures as S  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
mol = S.Molecule(S.water_cluster(n), "6-31G**")
db = engine.DeviceBasis(mol)
db.schwarz()
stream = db.torch_stream()
N = db.nbf


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


G = {}


def build():
    G["dev"], _ = db.eri_tensor(1.0e-8, to_host=False)


t_build = timed(build, 2)
counts = db.counts
rng = np.random.default_rng(1234)
X = rng.uniform(-1, 1, (N, N))
Da = torch.from_numpy(0.5 * (X + X.T)).cuda()
Dt = 2 * Da
t_jk = timed(lambda: db.jk_stored(G["dev"], Dt, Da, Da), 5)
peaks = {}
try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    pass
hbm = peaks.get("hbm_gbs", 6650.0)
gbs = 8.0 * N ** 4 / (t_jk * 1e-3) / 1e9
print(json.dumps({"workload": "(H2O)%d 6-31G** stored-tensor mode, N=%d" % (n, N),
                  "tensor_build_ms": t_build, "tensor_bytes": 8 * N ** 4,
                  "eris_per_sec_tensor_build": counts["all_eris"] / (t_build * 1e-3),
                  "jk_stored_ms": t_jk,
                  "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                               "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650"}}))
