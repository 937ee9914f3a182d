#!/usr/bin/env python
"""Phase timing of the sharded host path of DeviceBasis.jk_direct at N ranks (torchrun): a
synchronise after every phase, host clock, median over repetitions.  Diagnostic only."""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as tdist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pychem_b200 import _lib, dist, engine, structures as S  # noqa: E402
from pychem_b200.engine import _ptr  # noqa: E402


def main():
    rank, world, local = dist.init("nccl")
    mol = S.Molecule(S.water_cluster(int(os.environ.get("WATERS", "32"))), "6-31G**")
    db = engine.DeviceBasis(mol, device=local)
    db.schwarz()
    db.plan(1.0e-8, rank, world)
    N = db.nbf
    rng = np.random.default_rng(7)
    X = rng.uniform(-1, 1, (N, N))
    Da_t = torch.from_numpy(0.5 * (X + X.T)).pin_memory()
    Dt_t = (2 * Da_t).pin_memory()
    Db_t = Da_t.clone().pin_memory()
    Dt, Da, Db = Dt_t.numpy(), Da_t.numpy(), Db_t.numpy()
    for _ in range(3):
        db.jk_direct(Dt, Da, Db)
    share = db._share[1]
    acc = db.accumulator()
    sync = lambda: torch.cuda.synchronize(local)  # noqa: E731
    rows = []
    for rep in range(15):
        tdist.barrier()
        sync()
        t = [time.perf_counter()]
        d = db._gather_densities(share, Dt, Da, Db, None)
        sync(); t.append(time.perf_counter())
        db._order_after_torch(d[0])
        v = ctypes.c_int()
        _lib.check(db.lib.pc_jk_direct_accumulate_auto(db.h, _ptr(d[0]), _ptr(d[1]), _ptr(d[2]), _ptr(acc), ctypes.byref(v)))
        sync(); t.append(time.perf_counter())
        nn = N * N
        tdist.all_reduce(acc[:2 * nn])
        sync(); t.append(time.perf_counter())
        out = db._finalize_shared(share, v.value, acc)
        t.append(time.perf_counter())
        rows.append([1e3 * (b - a) for a, b in zip(t, t[1:])])
    med = np.median(np.array(rows), axis=0)
    # the whole call, unsplit
    tdist.barrier(); sync()
    t0 = time.perf_counter()
    for _ in range(20):
        db.jk_direct(Dt, Da, Db)
    whole = 1e3 * (time.perf_counter() - t0) / 20
    # the same call with whole matrices per rank (no sharing), and through the reference-facing mirror
    os.environ["PYCHEM_B200_SHARED_RESULTS"] = "0"
    db.drop_share()
    for _ in range(3):
        db.jk_direct(Dt, Da, Db)
    tdist.barrier(); sync()
    t0 = time.perf_counter()
    for _ in range(20):
        db.jk_direct(Dt, Da, Db)
    whole_unshared = 1e3 * (time.perf_counter() - t0) / 20
    if rank == 0:
        print(json.dumps({"whole_call_unshared_ms": round(whole_unshared, 3)}))
    # through the reference-facing mirror (what bench.py times), both modes
    from pychem_b200 import hartree_fock as hf_gpu

    class _Spin:
        pass

    class _State:
        def __init__(self, Dt, Da, Db):
            self.Total, self.Alpha, self.Beta = _Spin(), _Spin(), _Spin()
            self.Total.Density, self.Alpha.Density, self.Beta.Density = Dt, Da, Db
    state = _State(Dt, Da, Db)
    hf_gpu.evaluate_2e_ints(mol)
    for mode in ("1", "0", "1"):
        os.environ["PYCHEM_B200_SHARED_RESULTS"] = mode
        hf_gpu._state_for(mol)["db"].drop_share()
        for _ in range(3):
            hf_gpu.make_coulomb_exchange_matrices(mol, state)
        tdist.barrier(); sync()
        t0 = time.perf_counter()
        for _ in range(20):
            hf_gpu.make_coulomb_exchange_matrices(mol, state)
        tdist.barrier(); sync()
        tm = 1e3 * (time.perf_counter() - t0) / 20
        if rank == 0:
            print(json.dumps({"mirror_ms": round(tm, 3), "shared": mode, "same_db": hf_gpu._state_for(mol)["db"] is db}))
    if rank == 0:
        print(json.dumps({"world": world, "phases_ms": dict(zip(["upload_slices+all_gather+permute", "accumulate(+classify)", "all_reduce", "finalize+download_slices+barrier"], [round(float(x), 3) for x in med])),
                          "whole_call_ms": round(whole, 3)}))
    db.close()
    tdist.destroy_process_group()


if __name__ == "__main__":
    main()
