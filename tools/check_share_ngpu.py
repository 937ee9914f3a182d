#!/usr/bin/env python
"""N-rank check of the sharded host path of DeviceBasis.jk_direct (run under torchrun, one rank per
GPU): host (numpy) densities -> row-sliced upload + all-gather, row-sliced download into the shared
host buffer -- against the same call with device tensors (whole matrices, no sharing), for closed-
shell, open-shell and general densities, several calls in a row (buffer rotation), and against the
one-rank result computed by rank 0 on its own plan."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as tdist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pychem_b200 import dist, engine, structures as S  # noqa: E402


def main():
    rank, world, local = dist.init("nccl")
    dev = torch.device("cuda", local)
    mol = S.Molecule(S.water_cluster(int(os.environ.get("WATERS", "4"))), "6-31G**")
    db = engine.DeviceBasis(mol, device=local)
    db.schwarz()
    db.plan(1.0e-8, rank, world)
    N = db.nbf
    rng = np.random.default_rng(7)
    sym = lambda: (lambda X: 0.5 * (X + X.T))(rng.uniform(-1, 1, (N, N)))  # noqa: E731
    Da, Db = sym(), sym()
    A, B = rng.uniform(-1, 1, (N, N)), rng.uniform(-1, 1, (N, N))
    cases = {"rhf": (2 * Da, Da, Da.copy()), "uhf": (Da + Db, Da, Db), "gen": (A + B, A, B)}
    worst = {}
    for rep in range(3):
        for name, (Dt, D1, D2) in cases.items():
            J, Xa, Xb = db.jk_direct(Dt, D1, D2)                     # host arrays: sharded path
            Jh, Xah, Xbh = np.array(J), np.array(Xa), np.array(Xb)
            t = [torch.from_numpy(x).to(dev) for x in (Dt, D1, D2)]
            Jd, Xad, Xbd = db.jk_direct(*t)                          # device tensors: plain path
            err = max(float(np.abs(Jh - Jd.cpu().numpy()).max()), float(np.abs(Xah - Xad.cpu().numpy()).max()),
                      float(np.abs(Xbh - Xbd.cpu().numpy()).max()))
            worst[name] = max(worst.get(name, 0.0), err)
    shared = bool(db._share and db._share[1] is not None)
    # one-rank reference on rank 0 (its own basis object and plan)
    ref_err = None
    if rank == 0:
        db1 = engine.DeviceBasis(mol, device=local)
        db1.schwarz()
        db1.plan(1.0e-8, 0, 1)
        J1, Xa1, _ = db1.jk_direct(*cases["rhf"])
        ref = (np.array(J1), np.array(Xa1))
    J, Xa, _ = db.jk_direct(*cases["rhf"])
    if rank == 0:
        ref_err = max(float(np.abs(np.array(J) - ref[0]).max()), float(np.abs(np.array(Xa) - ref[1]).max()))
    errs = [None] * world
    tdist.all_gather_object(errs, worst)
    if rank == 0:
        print(json.dumps({"world": world, "nbf": N, "shared_host_buffer": shared,
                          "max_abs_diff_sharded_vs_device_inputs_per_rank": errs,
                          "max_abs_diff_vs_one_rank": ref_err}))
    db.close()
    tdist.destroy_process_group()


if __name__ == "__main__":
    main()
