"""TEST INFRASTRUCTURE ONLY: drive tests/emu/libpychem_b200_emu.so (the host emulation of the
CUDA kernels, see tests/emu/include/cuda_runtime.h) through the same C ABI as the product.

`EmuBasis` mirrors the methods of pychem_b200.engine.DeviceBasis that the parity tests use, with
numpy arrays standing in for device buffers.  Nothing in pychem_b200/ imports this module.
"""
import ctypes
import os

import numpy as np

from pychem_b200 import _lib
from pychem_b200.basis_table import BasisTable

HERE = os.path.dirname(os.path.abspath(__file__))
AUTO, RHF, UHF, GEN = 0, 2, 3, 4

_EMU = None


def load(build=True):
    global _EMU
    if _EMU is None:
        path = os.environ.get("PYCHEM_B200_EMU_LIB") or os.path.join(HERE, "libpychem_b200_emu.so")
        if os.environ.get("PYCHEM_B200_EMU_LIB"):
            pass                      # a generator variant built by tools/build_variant.py --emu
        elif build:
            from tests.emu import build_emu
            path = build_emu.build()
        lib = ctypes.CDLL(path)
        for name, argtypes in _lib.SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = ctypes.c_char_p if name.endswith("last_error") else ctypes.c_int
        lib.pcemu_launches.restype = ctypes.c_ulonglong
        lib.pcemu_switches.restype = ctypes.c_ulonglong
        _EMU = lib
    return _EMU


class EmuError(RuntimeError):
    pass


def _p(x):
    if x is None:
        return None
    assert x.dtype == np.float64 and x.flags["C_CONTIGUOUS"]
    return ctypes.c_void_p(x.ctypes.data)


def _f64(x):
    return np.ascontiguousarray(x, dtype=np.float64)


class EmuBasis:
    def __init__(self, molecule_or_table):
        self.table = (molecule_or_table if isinstance(molecule_or_table, BasisTable)
                      else BasisTable(molecule_or_table))
        self.lib = load()
        t = self.table
        h = ctypes.c_void_p()
        ip = lambda a: a.ctypes.data_as(_lib.c_ip)      # noqa: E731
        dp = lambda a: a.ctypes.data_as(_lib.c_dp)      # noqa: E731
        self.check(self.lib.pc_basis_create(0, t.nshell, ip(t.l), ip(t.K), ip(t.is_cart), ip(t.first_fn),
                                            dp(t.centres), dp(t.exps), dp(t.scc), ctypes.byref(h)))
        self.h = h
        self.nbf = t.nbf
        self.nshell = t.nshell
        self.npair = t.nshell * (t.nshell + 1) // 2
        self.counts = None

    def check(self, status):
        if status != 0:
            raise EmuError((self.lib.pc_last_error() or b"").decode() or "emulation: unknown error")

    def close(self):
        if getattr(self, "h", None):
            self.lib.pc_basis_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    ints_type = 0

    def set_ints_type(self, ints_type=0, grid_value=-1.0):
        self.check(self.lib.pc_basis_set_ints_type(self.h, int(ints_type), float(grid_value)))
        key = (int(ints_type), float(grid_value) if int(ints_type) == 1 else None)
        if key != getattr(self, "_ints_key", (0, None)):
            self.counts = None            # the C side dropped its plan (as engine.DeviceBasis does)
        self._ints_key = key
        self.ints_type = int(ints_type)

    def schwarz(self):
        bounds = np.zeros((self.npair, 49))
        pmax = np.zeros(self.npair)
        self.check(self.lib.pc_schwarz(self.h, bounds.ctypes.data_as(_lib.c_dp), pmax.ctypes.data_as(_lib.c_dp)))
        return bounds, pmax

    def plan(self, thresh=1.0e-8, rank=0, nranks=1):
        v = [ctypes.c_longlong() for _ in range(4)]
        self.check(self.lib.pc_plan(self.h, float(thresh), int(rank), int(nranks), *[ctypes.byref(x) for x in v]))
        self.counts = dict(my_quartets=v[0].value, my_eris=v[1].value, all_quartets=v[2].value,
                           all_eris=v[3].value, thresh=float(thresh), rank=rank, nranks=nranks)
        return self.counts

    def one_electron(self, charges, positions):
        Z = _f64(charges)
        R = _f64(positions).reshape(-1, 3)
        core = np.empty((self.nbf, self.nbf))
        overlap = np.empty((self.nbf, self.nbf))
        self.check(self.lib.pc_one_electron(self.h, len(Z), Z.ctypes.data_as(_lib.c_dp),
                                            R.ctypes.data_as(_lib.c_dp), _p(core), _p(overlap)))
        return core, overlap

    def eri_quartets(self, quartets):
        q = np.ascontiguousarray(np.asarray(quartets, dtype=np.int32).reshape(-1, 4))
        nfn = self.table.nfn
        sizes = [int(nfn[a] * nfn[b] * nfn[c] * nfn[d]) for a, b, c, d in q]
        offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        out = np.zeros(int(offs[-1]))
        self.check(self.lib.pc_eri_quartets(self.h, len(q), q.ctypes.data_as(_lib.c_ip),
                                            offs.ctypes.data_as(_lib.c_llp), out.ctypes.data_as(_lib.c_dp)))
        return [out[offs[k]:offs[k + 1]].reshape(nfn[a], nfn[b], nfn[c], nfn[d]) for k, (a, b, c, d) in enumerate(q)]

    def eri_tensor(self, thresh=1.0e-8):
        self.plan(thresh, 0, 1)
        N = self.nbf
        G = np.empty((N, N, N, N))
        self.check(self.lib.pc_eri_tensor(self.h, _p(G), None))
        return G

    def jk_stored(self, G, Dt, Da, Db):
        Dt, Da, Db = _f64(Dt), _f64(Da), _f64(Db)
        J, Xa, Xb = (np.empty((self.nbf, self.nbf)) for _ in range(3))
        self.check(self.lib.pc_jk_stored(self.h, _p(G), _p(Dt), _p(Da), _p(Db), _p(J), _p(Xa), _p(Xb)))
        return J, Xa, Xb

    def jk_direct(self, Dt, Da, Db, variant=AUTO):
        Dt, Da, Db = _f64(Dt), _f64(Da), _f64(Db)
        if self.counts is None:
            self.plan()
        J, Xa, Xb = (np.empty((self.nbf, self.nbf)) for _ in range(3))
        self.check(self.lib.pc_jk_direct(self.h, variant, _p(Dt), _p(Da), _p(Db), _p(J), _p(Xa), _p(Xb)))
        return J, Xa, Xb

    def jk_direct_partial(self, Dt, Da, Db, variant):
        """The un-finalised accumulator of this rank's slice (multi-rank additivity checks)."""
        Dt, Da, Db = _f64(Dt), _f64(Da), _f64(Db)
        acc = np.empty(3 * self.nbf * self.nbf)
        self.check(self.lib.pc_jk_direct_accumulate(self.h, variant, _p(Dt), _p(Da), _p(Db), _p(acc)))
        return acc

    def jk_finalize(self, acc, variant):
        J, Xa, Xb = (np.empty((self.nbf, self.nbf)) for _ in range(3))
        self.check(self.lib.pc_jk_finalize(self.h, variant, _p(acc), _p(J), _p(Xa), _p(Xb)))
        return J, Xa, Xb

    def jk_stored_batch(self, G, D):
        D = _f64(D)
        out = np.empty_like(D)
        self.check(self.lib.pc_jk_stored_batch(self.h, _p(G), int(D.shape[0]), _p(D), _p(out)))
        return out

    def jk_direct_batch(self, D):
        D = _f64(D)
        if self.counts is None:
            self.plan()
        out = np.empty_like(D)
        self.check(self.lib.pc_jk_direct_batch(self.h, int(D.shape[0]), _p(D), _p(out)))
        return out

    def jk_direct_batch_partial(self, D):
        D = _f64(D)
        acc = np.empty(D.size)
        self.check(self.lib.pc_jk_direct_batch_accumulate(self.h, int(D.shape[0]), _p(D), _p(acc)))
        return acc

    def jk_finalize_batch(self, acc, nset):
        out = np.empty((nset, 3, self.nbf, self.nbf))
        self.check(self.lib.pc_jk_finalize_batch(self.h, nset, _p(acc), _p(out)))
        return out


class EmuDeviceBasis(EmuBasis):
    """EmuBasis with the method signatures of pychem_b200.engine.DeviceBasis, so that a CPU test
    can run the host mirrors (pychem_b200.hartree_fock / noci) end to end over the emulated
    kernels by monkeypatching ``pychem_b200.integrals.DeviceBasis``."""

    def __init__(self, molecule_or_table, device=None):
        super().__init__(molecule_or_table)

    def eri_tensor(self, thresh=1.0e-8, to_host=True):
        G = super().eri_tensor(thresh)
        return G, (G if to_host else None)

    def jk_direct(self, Dt, Da, Db, variant=None, group=None):
        return super().jk_direct(Dt, Da, Db, AUTO if variant is None else variant)

    def jk_direct_batch(self, D, group=None):
        return super().jk_direct_batch(D)
