"""N>1 path on CPU: world_size-2 gloo process group.  Each rank digests its slice of the unique
shell quartets of LiH/6-31G (numpy stand-in for the device accumulate, using the golden tensor
minted from the reference), the partial half-accumulators are summed with ONE all-reduce and
finalised -- the same plumbing pychem_b200.engine.DeviceBasis.jk_direct drives with NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as tdist
import torch.multiprocessing as mp

from pychem_b200 import dist
from pychem_b200.basis_table import BasisTable
from tests import helpers

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lih_631g.npz")


def partial_accumulators(G, table, Dt, Da, Db, rank, nranks, general):
    """Half-accumulators [J|Ka|Kb] of this rank's slice of the unique shell quartets, with the
    degeneracy factors and update rules of pc_digest_jk (csrc/pc_common.cuh)."""
    N = table.nbf
    acc = np.zeros((3, N, N))
    pairs = [(a, b) for a in range(table.nshell) for b in range(a, table.nshell)]
    quartets = [(p, q) for p in range(len(pairs)) for q in range(p, len(pairs))]
    lo, hi = dist.slice_bounds(len(quartets), rank, nranks)
    rng = lambda s: range(int(table.first_fn[s]), int(table.first_fn[s] + table.nfn[s]))  # noqa: E731
    for p, q in quartets[lo:hi]:
        (a, b), (c, d) = pairs[p], pairs[q]
        fac = (0.5 if a == b else 1.0) * (0.5 if c == d else 1.0) * (0.5 if p == q else 1.0)
        for i in rng(a):
            for j in rng(b):
                for k in rng(c):
                    for l in rng(d):
                        g = fac * G[i, j, k, l]
                        acc[0, i, j] += (Dt[k, l] + Dt[l, k]) * g
                        acc[0, k, l] += (Dt[i, j] + Dt[j, i]) * g
                        for s, D in ((1, Da), (2, Db)):
                            acc[s, i, l] += D[k, j] * g
                            acc[s, j, l] += D[k, i] * g
                            acc[s, i, k] += D[l, j] * g
                            acc[s, j, k] += D[l, i] * g
                            if general:
                                acc[s, k, j] += D[i, l] * g
                                acc[s, k, i] += D[j, l] * g
                                acc[s, l, j] += D[i, k] * g
                                acc[s, l, i] += D[j, k] * g
    return acc


def _worker(rank, world, port, general, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, w, _ = dist.init("gloo")
    assert (r, w) == (rank, world)
    g = np.load(GOLD)
    table = BasisTable(helpers.molecule("lih"))
    if general:
        Dt, Da, Db = g["Dt"], g["Da"], g["Db"]                 # non-symmetric (NOCI-shaped)
    else:
        Da = 0.5 * (g["Da"] + g["Da"].T)
        Db = 0.5 * (g["Db"] + g["Db"].T)
        Dt = Da + Db
    acc = torch.from_numpy(partial_accumulators(g["G"], table, Dt, Da, Db, rank, world, general).ravel())
    dist.allreduce_sum_(acc)
    J, Xa, Xb = dist.finalize_accumulators(acc.numpy(), table.nbf, 4 if general else 3)
    ref = (np.einsum("cd,abcd->ab", Dt, g["G"]), np.einsum("cb,abcd->ad", -Da, g["G"]),
           np.einsum("cb,abcd->ad", -Db, g["G"]))
    err = max(float(np.abs(x - y).max()) for x, y in zip((J, Xa, Xb), ref))
    if rank == 0:
        out.put(err)
    tdist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("general", [False, True])
def test_two_rank_partition_allreduce_finalize(general):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, general, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert out.get(timeout=5) < 1e-11


# ---------------------------------------------------------------------------------------------
# NodeShare: the host buffer the ranks of one node share for the results (pychem_b200/dist.py)
# ---------------------------------------------------------------------------------------------
def _share_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init("gloo")
    N = 37                                     # not a multiple of the world size: ragged last slice
    sh = dist.NodeShare((3, N, N), pin=False)
    lo, hi = sh.rows(N)
    ok = True
    for call in range(5):                      # more calls than buffers: the rotation is exercised
        full = np.arange(3 * N * N, dtype=float).reshape(3, N, N) + 1000.0 * call
        buf = sh.buffer()
        buf[:, lo:hi] = full[:, lo:hi]          # this rank's rows only
        if rank == 1:
            import time
            time.sleep(0.02 * (call % 2))      # skewed arrival at the barrier
        sh.barrier()
        ok = ok and bool((buf == full).all())  # everybody sees everybody's rows
        prev = buf
    # the previous call's view is still intact while the next buffer is being filled
    nxt = sh.buffer()
    nxt[:, lo:hi] = -1.0
    ok = ok and bool((prev == full).all()) and nxt is not prev
    sh.barrier()
    rows = [None] * world
    tdist.all_gather_object(rows, (lo, hi))
    if rank == 0:
        covered = sorted(rows)
        ok = ok and covered[0][0] == 0 and covered[-1][1] == N and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
        out.put(ok)
    sh.close()
    tdist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_node_share_publishes_row_slices_to_every_rank(world):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_share_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=10) is True
