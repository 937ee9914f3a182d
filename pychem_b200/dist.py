"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on GPUs, gloo in CPU tests).

The path shards naturally: shell quartets are independent and J/K are
linear in the ERIs.  Every rank holds the full (tiny) basis and density matrices, digests its
slice of every (bra bucket, ket bucket) task range (pc_plan) and the partial half-accumulators
[J | Ka | Kb] are summed with ONE all-reduce per Fock build.
"""
import os


def env_rank():
    """(rank, world_size, local_rank) from the torchrun environment (1 process if absent)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init(backend=None):
    """Initialise the default process group from the environment; returns (rank, world, local)."""
    import torch
    import torch.distributed as dist
    rank, world, local = env_rank()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def slice_bounds(total, rank, nranks):
    """Equal contiguous slices [begin, end) of `total` items -- the schedule the gloo tests of the
    N>1 path use for host-side work.  pc_plan (csrc/pc_api.cu) does NOT cut this way: it cuts
    every bucket pair at SEGMENT boundaries into nranks slices of equal modelled cost
    (primitive quartets x flop_prim + quartets x per-quartet cost); the slices of all bucket
    pairs of a class run in one fused launch."""
    begin = total * rank // nranks
    end = total * (rank + 1) // nranks
    return begin, end


def allreduce_sum_(tensor, group=None):
    """In-place SUM all-reduce of the packed accumulators (no-op on one rank)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
    return tensor


def finalize_accumulators(acc, nbf, variant):
    """Host mirror of jk_finalize_kernel (csrc/pc_api.cu) for accumulators that already live on
    the host (used by the gloo tests of the N>1 path): J = A + A^T; X = -(K + K^T) for symmetric
    densities, -K for general ones."""
    import numpy as np
    a = np.asarray(acc, dtype=float).reshape(3, nbf, nbf)
    general = variant == 4
    J = a[0] + a[0].T
    Xa = -(a[1] if general else a[1] + a[1].T)
    Xb = Xa.copy() if variant == 2 else -(a[2] if general else a[2] + a[2].T)
    return J, Xa, Xb
