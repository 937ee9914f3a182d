"""Thin Python owner of a `pc_basis` handle (include/pychem_b200.h).

PyTorch is used for two things only: owning device buffers (the dense tensor, the J/K
accumulators) and the NCCL all-reduce of the partial J/K accumulators over ranks.
All arithmetic happens in the CUDA library behind the C ABI.
"""
import ctypes

import numpy as np

from . import _lib
from .basis_table import BasisTable

AUTO, RHF, UHF, GEN = 0, 2, 3, 4  # PC_JK_* in include/pychem_b200.h
INTEGRAL_THRESHOLD = 1.0e-8        # Data/constants.py:31


def _ptr(x):
    """void* of a numpy array (host) or a torch tensor (device/host)."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        if x.dtype != np.float64 or not x.flags["C_CONTIGUOUS"]:
            raise TypeError("expected a C-contiguous float64 array")
        return ctypes.c_void_p(x.ctypes.data)
    return ctypes.c_void_p(x.data_ptr())        # torch.Tensor


def _as_f64(x):
    if isinstance(x, np.ndarray):
        return np.ascontiguousarray(x, dtype=np.float64)
    if hasattr(x, "data_ptr"):
        import torch
        if x.dtype != torch.float64 or not x.is_contiguous():
            raise TypeError("expected a contiguous float64 tensor")
        return x
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


def classify_densities(Dt, Da, Db, tol=0.0):
    """Pick the cheapest exact digestion variant for these densities."""
    Dt, Da, Db = (np.asarray(x) for x in (Dt, Da, Db))
    sym = all(np.array_equal(x, x.T) for x in (Dt, Da, Db))
    if not sym:
        return GEN
    return RHF if np.array_equal(Da, Db) else UHF


class DeviceBasis:
    """Device-resident basis + shell-pair tables + screened quartet plan for one molecule."""

    def __init__(self, molecule_or_table, device=None):
        self.table = (molecule_or_table if isinstance(molecule_or_table, BasisTable)
                      else BasisTable(molecule_or_table))
        self.lib = _lib.load()
        if device is None:
            import torch
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        self.device = int(device)
        t = self.table
        h = ctypes.c_void_p()
        ip = lambda a: a.ctypes.data_as(_lib.c_ip)      # noqa: E731
        dp = lambda a: a.ctypes.data_as(_lib.c_dp)      # noqa: E731
        _lib.check(self.lib.pc_basis_create(self.device, t.nshell, ip(t.l), ip(t.K), ip(t.is_cart),
                                            ip(t.first_fn), dp(t.centres), dp(t.exps), dp(t.scc),
                                            ctypes.byref(h)))
        self.h = h
        self.nbf = t.nbf
        self.nshell = t.nshell
        self.npair = t.nshell * (t.nshell + 1) // 2
        self.counts = None
        self._acc = None
        self.ints_type = 0

    def drop_share(self):
        """Release the host buffer shared with the other ranks (the next multi-rank jk_direct with
        host arrays sets it up again, reading PYCHEM_B200_SHARED_RESULTS)."""
        cur = getattr(self, "_share", None)
        if cur is not None and cur[1] is not None:
            cur[1].close()
        self._share = None

    def close(self):
        self.drop_share()
        if getattr(self, "h", None):
            self.lib.pc_basis_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ plumbing
    def stream_ptr(self):
        s = ctypes.c_void_p()
        _lib.check(self.lib.pc_basis_stream(self.h, ctypes.byref(s)))
        return s.value

    def torch_stream(self):
        import torch
        return torch.cuda.ExternalStream(self.stream_ptr(), device=self.device)

    def _order_after_torch(self, *args):
        """The library works on its own (non-blocking) stream.  When a caller hands over CUDA
        tensors, they were produced -- or their memory was recycled by torch's caching allocator
        -- on torch's current stream: make the library's stream wait for everything queued there
        so far before it reads or overwrites them."""
        if not any(hasattr(x, "is_cuda") and x.is_cuda for x in args):
            return
        import torch
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.torch_stream().wait_event(ev)

    def launch_count(self):
        n = ctypes.c_longlong()
        _lib.check(self.lib.pc_launch_count(self.h, ctypes.byref(n)))
        return n.value

    # ------------------------------------------------------------------ Schwarz + plan
    def set_ints_type(self, ints_type=0, grid_value=-1.0):
        """0 = electron repulsion, 1 = scattering kernel at ``grid_value`` (the ints_type /
        grid_value arguments of integrals.two_electron, Methods/integrals.py:427).  Switching
        drops the Schwarz factors and the plan."""
        _lib.check(self.lib.pc_basis_set_ints_type(self.h, int(ints_type), float(grid_value)))
        key = (int(ints_type), float(grid_value) if int(ints_type) == 1 else None)
        if key != getattr(self, "_ints_key", (0, None)):
            # the C side dropped bounds, pair order and plan: forget ours, keep the slicing so
            # that the next direct J/K re-plans for the same rank / nranks
            if self.counts is not None:
                self._replan = (self.counts["thresh"], self.counts["rank"], self.counts["nranks"])
            self.counts = None
        self._ints_key = key
        self.ints_type = int(ints_type)

    def _ensure_plan(self):
        if self.counts is None:
            self.plan(*getattr(self, "_replan", (INTEGRAL_THRESHOLD, 0, 1)))

    def schwarz(self):
        """(bounds[npair,49], pmax[npair]); hartree_fock.py:244-254."""
        bounds = np.zeros((self.npair, 49))
        pmax = np.zeros(self.npair)
        _lib.check(self.lib.pc_schwarz(self.h, bounds.ctypes.data_as(_lib.c_dp),
                                       pmax.ctypes.data_as(_lib.c_dp)))
        return bounds, pmax

    def plan(self, thresh=INTEGRAL_THRESHOLD, rank=0, nranks=1):
        v = [ctypes.c_longlong() for _ in range(4)]
        _lib.check(self.lib.pc_plan(self.h, float(thresh), int(rank), int(nranks),
                                    *[ctypes.byref(x) for x in v]))
        self.counts = dict(my_quartets=v[0].value, my_eris=v[1].value,
                           all_quartets=v[2].value, all_eris=v[3].value,
                           thresh=float(thresh), rank=rank, nranks=nranks)
        return self.counts

    def set_profiling(self, on):
        _lib.check(self.lib.pc_set_profiling(self.h, int(bool(on))))

    def plan_items(self):
        """Per (bra bucket, ket bucket) launch: class, contraction depths, task counts and the
        device time of the last profiled accumulate."""
        n = ctypes.c_int()
        _lib.check(self.lib.pc_plan_items(self.h, 0, ctypes.byref(n), None, None, None, None, None))
        m = n.value
        cls = np.zeros((m, 4), dtype=np.int32)
        kprim = np.zeros((m, 2), dtype=np.int32)
        tasks = np.zeros((m, 2), dtype=np.int64)
        ms = np.zeros(m, dtype=np.float32)
        prim = np.zeros(m, dtype=np.float64)
        _lib.check(self.lib.pc_plan_items(self.h, m, ctypes.byref(n), cls.ctypes.data_as(_lib.c_ip),
                                          kprim.ctypes.data_as(_lib.c_ip),
                                          tasks.ctypes.data_as(_lib.c_llp),
                                          ms.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                                          prim.ctypes.data_as(_lib.c_dp)))
        self.prim_exec = prim
        return cls, kprim, tasks, ms

    # ------------------------------------------------------------------ one-electron matrices
    def one_electron(self, charges, positions):
        """(Core, Overlap), N x N each: kinetic + nuclear attraction, and overlap
        (hartree_fock.py:207-222).  positions in bohr."""
        Z = np.ascontiguousarray(charges, dtype=np.float64)
        R = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1, 3)
        core = np.empty((self.nbf, self.nbf))
        overlap = np.empty((self.nbf, self.nbf))
        _lib.check(self.lib.pc_one_electron(self.h, len(Z), Z.ctypes.data_as(_lib.c_dp),
                                            R.ctypes.data_as(_lib.c_dp), _ptr(core), _ptr(overlap)))
        return core, overlap

    # ------------------------------------------------------------------ ERIs
    def eri_quartets(self, quartets):
        """Blocks (nfa,nfb,nfc,nfd) for shell quartets [(a,b,c,d)], a<=b, c<=d."""
        q = np.ascontiguousarray(np.asarray(quartets, dtype=np.int32).reshape(-1, 4))
        nfn = self.table.nfn
        sizes = [int(nfn[a] * nfn[b] * nfn[c] * nfn[d]) for a, b, c, d in q]
        offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        out = np.zeros(int(offs[-1]))
        _lib.check(self.lib.pc_eri_quartets(self.h, len(q), q.ctypes.data_as(_lib.c_ip),
                                            offs.ctypes.data_as(_lib.c_llp),
                                            out.ctypes.data_as(_lib.c_dp)))
        return [out[offs[k]:offs[k + 1]].reshape(nfn[a], nfn[b], nfn[c], nfn[d])
                for k, (a, b, c, d) in enumerate(q)]

    def eri_tensor(self, thresh=INTEGRAL_THRESHOLD, to_host=True):
        """Dense (N,N,N,N) tensor on the device (torch tensor) and optionally a host copy."""
        import torch
        self.plan(thresh, 0, 1)
        N = self.nbf
        G_dev = torch.empty((N, N, N, N), dtype=torch.float64, device="cuda:%d" % self.device)
        G_host = np.empty((N, N, N, N)) if to_host else None
        self._order_after_torch(G_dev)          # the block may be recycled from pending torch work
        _lib.check(self.lib.pc_eri_tensor(self.h, _ptr(G_dev), _ptr(G_host)))
        return G_dev, G_host

    # ------------------------------------------------------------------ J/K
    def _outputs(self, like):
        N = self.nbf
        if isinstance(like, np.ndarray):
            # page-locked result arrays: the device->host copies run at full PCIe rate
            # (numpy views of pinned torch storage; torch's caching host allocator recycles them)
            import torch
            return [torch.empty((N, N), dtype=torch.float64, pin_memory=True).numpy() for _ in range(3)]
        import torch
        return [torch.empty((N, N), dtype=torch.float64, device=like.device) for _ in range(3)]

    def jk_stored(self, G_dev, Dt, Da, Db):
        Dt, Da, Db = _as_f64(Dt), _as_f64(Da), _as_f64(Db)
        J, Xa, Xb = self._outputs(Dt)
        self._order_after_torch(G_dev, Dt, Da, Db, J, Xa, Xb)
        _lib.check(self.lib.pc_jk_stored(self.h, _ptr(G_dev), _ptr(Dt), _ptr(Da), _ptr(Db),
                                         _ptr(J), _ptr(Xa), _ptr(Xb)))
        return J, Xa, Xb

    def accumulator(self):
        import torch
        if self._acc is None:
            self._acc = torch.empty(3 * self.nbf * self.nbf, dtype=torch.float64,
                                    device="cuda:%d" % self.device)
        return self._acc

    def jk_direct(self, Dt, Da, Db, variant=None, group=None):
        """Integral-direct J/K.  With an initialised torch.distributed process group and a plan
        built with nranks>1, partial accumulators are summed with one NCCL all-reduce.  With host
        (numpy) densities and all ranks on one node, every rank moves only its share of the rows of
        the densities and of the results over PCIe; the result arrays are then views of a buffer the
        ranks share and stay valid until the next jk_direct call on this basis."""
        Dt, Da, Db = _as_f64(Dt), _as_f64(Da), _as_f64(Db)
        if self.ints_type != 0:
            raise _lib.PychemB200Error("jk_direct: J/K digestion is defined for the repulsion integrals; "
                                       "call set_ints_type(0) (evaluate_2e_ints(molecule)) first")
        self._ensure_plan()
        share = None
        if self.counts["nranks"] > 1 and isinstance(Dt, np.ndarray):
            share = self._node_share(group)
        if share is None:
            J, Xa, Xb = self._outputs(Dt)
            self._order_after_torch(Dt, Da, Db, J, Xa, Xb)
        if self.counts["nranks"] == 1:
            if variant is None:
                # classified on the device; closed-shell densities: X_beta is X_alpha (the same
                # array object is returned for both, one device->host copy less; the reference's
                # callers only read the exchange matrices, hartree_fock.py:354-355, noci.py:250-292)
                v = ctypes.c_int()
                _lib.check(self.lib.pc_jk_direct_auto(self.h, _ptr(Dt), _ptr(Da), _ptr(Db),
                                                      _ptr(J), _ptr(Xa), _ptr(Xb), ctypes.byref(v)))
                return (J, Xa, Xa) if v.value == RHF else (J, Xa, Xb)
            _lib.check(self.lib.pc_jk_direct(self.h, variant, _ptr(Dt), _ptr(Da), _ptr(Db),
                                             _ptr(J), _ptr(Xa), _ptr(Xb)))
            return J, Xa, Xb
        import torch
        import torch.distributed as dist
        acc = self.accumulator()
        if share is not None:
            # 1/nranks of the rows of every density over this rank's PCIe link, the rest over NVLink
            Dt, Da, Db = self._gather_densities(share, Dt, Da, Db, group)
            self._order_after_torch(Dt)
        if variant is None:
            v = ctypes.c_int()
            _lib.check(self.lib.pc_jk_direct_accumulate_auto(self.h, _ptr(Dt), _ptr(Da), _ptr(Db), _ptr(acc),
                                                             ctypes.byref(v)))
            variant = v.value
        else:
            _lib.check(self.lib.pc_jk_direct_accumulate(self.h, variant, _ptr(Dt), _ptr(Da), _ptr(Db), _ptr(acc)))
        # the library launched on its own stream: order the collective after it
        ev = torch.cuda.Event()
        ev.record(self.torch_stream())
        torch.cuda.current_stream().wait_event(ev)
        nn = self.nbf * self.nbf
        # closed-shell: only [J | K_alpha] carry data (two thirds of the all-reduce, no X_beta copy)
        dist.all_reduce(acc[:2 * nn] if variant == RHF else acc, op=dist.ReduceOp.SUM, group=group)
        ev2 = torch.cuda.Event()
        ev2.record(torch.cuda.current_stream())
        self.torch_stream().wait_event(ev2)
        if share is not None:
            return self._finalize_shared(share, variant, acc)
        if variant == RHF:
            _lib.check(self.lib.pc_jk_finalize(self.h, variant, _ptr(acc), _ptr(J), _ptr(Xa), None))
            return J, Xa, Xa
        _lib.check(self.lib.pc_jk_finalize(self.h, variant, _ptr(acc), _ptr(J), _ptr(Xa), _ptr(Xb)))
        return J, Xa, Xb

    # ---- host traffic of the N>1 path, sharded over the ranks' PCIe links (pychem_b200/dist.py NodeShare)
    def _node_share(self, group):
        """NodeShare of this basis and group, or None (one node only; PYCHEM_B200_SHARED_RESULTS=0
        switches it off: every rank then moves whole matrices over its own link)."""
        import os
        key = id(group)
        cur = getattr(self, "_share", None)
        if cur is not None and cur[0] == key:
            return cur[1]
        sh = None
        if os.environ.get("PYCHEM_B200_SHARED_RESULTS", "1") != "0":
            from . import dist as pdist
            try:
                sh = pdist.NodeShare((3, self.nbf, self.nbf), group=group)
            except pdist.NodeShareUnavailable:
                sh = None                              # (a collective decision: the same on every rank)
        self._share = (key, sh)
        return sh

    def _gather_densities(self, share, Dt, Da, Db, group):
        import torch
        import torch.distributed as dist
        N, W = self.nbf, share.world
        R = (N + W - 1) // W
        dev = "cuda:%d" % self.device
        if getattr(self, "_dens_dev", None) is None:
            self._dens_chunk = torch.zeros((3, R, N), dtype=torch.float64, device=dev)
            self._dens_all = torch.empty((W, 3, R, N), dtype=torch.float64, device=dev)
            self._dens_dev = torch.empty((3, W * R, N), dtype=torch.float64, device=dev)
        lo, hi = share.rows(N)
        same = Db is Da
        for m, D in enumerate((Dt, Da, Db)):
            if m == 2 and same:
                break
            if hi > lo:
                self._dens_chunk[m, :hi - lo].copy_(torch.from_numpy(D[lo:hi]), non_blocking=True)
        dist.all_gather_into_tensor(self._dens_all.view(-1), self._dens_chunk.view(-1), group=group)
        self._dens_dev.view(3, W, R, N).copy_(self._dens_all.permute(1, 0, 2, 3))
        full = self._dens_dev
        return full[0, :N], full[1, :N], (full[1, :N] if same else full[2, :N])

    def _finalize_shared(self, share, variant, acc):
        import torch
        N = self.nbf
        dev = "cuda:%d" % self.device
        if getattr(self, "_out_dev", None) is None:
            self._out_dev = torch.empty((3, N, N), dtype=torch.float64, device=dev)
        out = self._out_dev
        nout = 2 if variant == RHF else 3
        _lib.check(self.lib.pc_jk_finalize(self.h, variant, _ptr(acc), _ptr(out[0]), _ptr(out[1]),
                                           None if variant == RHF else _ptr(out[2])))
        buf = share.buffer()
        lo, hi = share.rows(N)
        if hi > lo:
            host = torch.from_numpy(buf)
            for m in range(nout):                      # contiguous row blocks: plain async copies
                host[m, lo:hi].copy_(out[m, lo:hi], non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        share.barrier()
        return buf[0], buf[1], (buf[1] if variant == RHF else buf[2])


    # ------------------------------------------------------------------ batched J/K (NOCI)
    def _batch_in(self, D):
        D = _as_f64(D)
        N = self.nbf
        if tuple(D.shape[1:]) != (3, N, N):
            raise ValueError("batched densities must have shape (nset, 3, N, N): [Dt, Da, Db] per set")
        return D, int(D.shape[0])

    def _batch_out(self, like, nset):
        import torch
        N = self.nbf
        if isinstance(like, np.ndarray):
            return torch.empty((nset, 3, N, N), dtype=torch.float64, pin_memory=True).numpy()
        return torch.empty((nset, 3, N, N), dtype=torch.float64, device=like.device)

    def jk_stored_batch(self, G_dev, D):
        """J/K for D[s] = (Dt, Da, Db), s = 0..nset-1, from the stored tensor: out[s] = (J, Xa, Xb)."""
        D, nset = self._batch_in(D)
        out = self._batch_out(D, nset)
        self._order_after_torch(G_dev, D, out)
        _lib.check(self.lib.pc_jk_stored_batch(self.h, _ptr(G_dev), nset, _ptr(D), _ptr(out)))
        return out

    def jk_direct_batch(self, D, group=None):
        """Integral-direct J/K for all sets in one pass over the ERIs (general variant).  With a
        multi-rank plan the partial accumulators of all sets are summed with one all-reduce."""
        D, nset = self._batch_in(D)
        if self.ints_type != 0:
            raise _lib.PychemB200Error("jk_direct_batch: J/K digestion is defined for the repulsion integrals")
        self._ensure_plan()
        out = self._batch_out(D, nset)
        self._order_after_torch(D, out)
        if self.counts["nranks"] == 1:
            _lib.check(self.lib.pc_jk_direct_batch(self.h, nset, _ptr(D), _ptr(out)))
            return out
        import torch
        import torch.distributed as dist
        acc = torch.empty(nset * 3 * self.nbf * self.nbf, dtype=torch.float64, device="cuda:%d" % self.device)
        _lib.check(self.lib.pc_jk_direct_batch_accumulate(self.h, nset, _ptr(D), _ptr(acc)))
        ev = torch.cuda.Event()
        ev.record(self.torch_stream())
        torch.cuda.current_stream().wait_event(ev)
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
        ev2 = torch.cuda.Event()
        ev2.record(torch.cuda.current_stream())
        self.torch_stream().wait_event(ev2)
        _lib.check(self.lib.pc_jk_finalize_batch(self.h, nset, _ptr(acc), _ptr(out)))
        return out


def fp64_peak_tflops(device=0):
    v = ctypes.c_double()
    _lib.check(_lib.load().pc_fp64_peak(int(device), ctypes.byref(v)))
    return v.value
