"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on GPUs, gloo in CPU tests).

The path shards naturally: shell quartets are independent and J/K are
linear in the ERIs.  Every rank holds the full (tiny) basis and density matrices, digests its
slice of every (bra bucket, ket bucket) task range (pc_plan) and the partial half-accumulators
[J | Ka | Kb] are summed with ONE all-reduce per Fock build.

Host traffic of the N>1 path (NodeShare below): the ranks of one node are N copies of the same
driver, so every rank would push the same 3 N^2 doubles up its PCIe link and pull the same 2-3 N^2
down -- at 8 ranks that costs more than the all-reduce.  Instead every rank uploads 1/N of the rows
of each density (all-gather over NVLink), downloads 1/N of the rows of each result into a host
buffer the ranks share, and a flag barrier in that buffer publishes the whole result to all of them.
"""
import os
import time


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


class NodeShareUnavailable(RuntimeError):
    pass


class NodeShare:
    """Result buffers in host memory shared by the ranks of ONE node (a file in /dev/shm mapped by
    every rank, page-locked for the device->host copies) plus a sequence-number barrier in the same
    mapping.  `nbuf` buffers are used in rotation: a result stays valid until the owner's NEXT call
    (a faster rank may already be filling the other buffer; it cannot reach the one after that
    before every rank has passed the barrier of the call in between)."""

    FLAG_STRIDE = 8            # int64 per rank, one cache line apart

    def __init__(self, shape, group=None, nbuf=2, pin=True, directory="/dev/shm"):
        import socket

        import numpy as np
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            raise NodeShareUnavailable("no process group")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        hosts = [None] * self.world
        dist.all_gather_object(hosts, socket.gethostname(), group=group)
        if len(set(hosts)) != 1 or not os.path.isdir(directory):
            raise NodeShareUnavailable("ranks on different hosts: %s" % sorted(set(hosts)))
        self.shape = tuple(int(x) for x in shape)
        nelem = int(np.prod(self.shape))
        self.nbuf = nbuf
        data_bytes = ((nbuf * nelem * 8 + 4095) // 4096) * 4096
        size = data_bytes + 4096 * ((self.world * self.FLAG_STRIDE * 8 + 4095) // 4096)
        path = [None]
        src = dist.get_global_rank(group, 0) if group is not None else 0
        if self.rank == 0:
            try:
                path[0] = os.path.join(directory, "pychem_b200_%d_%x" % (os.getpid(), int(time.time() * 1e6) & 0xffffffff))
                with open(path[0], "wb") as fh:
                    fh.truncate(size)                  # zero pages: flags start at sequence 0
            except OSError:
                path[0] = None
        dist.broadcast_object_list(path, src=src, group=group)
        err = None
        try:
            if path[0] is None:
                raise OSError("rank 0 could not create the file")
            self._map = np.memmap(path[0], dtype=np.uint8, mode="r+", shape=(size,))
        except (OSError, ValueError) as e:
            err = str(e)
        errs = [None] * self.world                     # every rank takes the same decision
        dist.all_gather_object(errs, err, group=group)
        if self.rank == 0 and path[0] is not None:
            try:
                os.unlink(path[0])                     # the mappings keep it alive; nothing to leak
            except OSError:
                pass
        if any(e is not None for e in errs):
            self._map = None
            raise NodeShareUnavailable("shared host buffer: %s" % [e for e in errs if e is not None][0])
        self.buffers = [np.ndarray(self.shape, dtype=np.float64, buffer=self._map, offset=k * nelem * 8)
                        for k in range(nbuf)]
        self.flags = np.ndarray((self.world, self.FLAG_STRIDE), dtype=np.int64, buffer=self._map, offset=data_bytes)
        self.seq = 0
        self.pinned = False
        if pin:
            import torch
            if torch.cuda.is_available():
                try:
                    rc = torch.cuda.cudart().cudaHostRegister(self._map.ctypes.data, size, 0)
                    self.pinned = int(rc) == 0
                except RuntimeError:                   # torch raises on a CUDA error code
                    self.pinned = False
                self._registered = self._map.ctypes.data if self.pinned else None
            flags = [None] * self.world                # page-locked on every rank or used by none
            dist.all_gather_object(flags, bool(self.pinned), group=group)
            if not all(flags):
                self.close()
                raise NodeShareUnavailable("cudaHostRegister of the shared host buffer failed on rank(s) %s"
                                           % [r for r, f in enumerate(flags) if not f])

    def rows(self, n):
        """[lo, hi) of the n rows this rank moves over its PCIe link."""
        r = (n + self.world - 1) // self.world
        return min(n, self.rank * r), min(n, (self.rank + 1) * r)

    def buffer(self):
        """The buffer of the call in progress (advance with barrier())."""
        return self.buffers[self.seq % self.nbuf]

    def barrier(self, timeout=300.0):
        """Publish this rank's part (its stores and completed device->host copies precede the flag
        store) and wait for every other rank's."""
        self.seq += 1
        self.flags[self.rank, 0] = self.seq
        t0 = time.monotonic()
        spins = 0
        while int(self.flags[:, 0].min()) < self.seq:
            spins += 1
            if spins % 4096 == 0 and time.monotonic() - t0 > timeout:
                raise RuntimeError("NodeShare.barrier: rank %d waited %.0f s at sequence %d (flags %s)"
                                   % (self.rank, timeout, self.seq, self.flags[:, 0].tolist()))

    def close(self):
        if getattr(self, "_registered", None):
            import torch
            torch.cuda.cudart().cudaHostUnregister(self._registered)
            self._registered = None
        self.pinned = False
        self.buffers = []
        self.flags = None
        self._map = None
