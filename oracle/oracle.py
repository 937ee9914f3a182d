"""ctypes front-end of oracle/eri_oracle.c (the plain-C restatement of the reference path).

TEST INFRASTRUCTURE ONLY.  See eri_oracle.c for the reference file:line each routine follows.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_dp = ctypes.POINTER(ctypes.c_double)
c_ip = ctypes.POINTER(ctypes.c_int)


def build(force=False):
    so = os.path.join(HERE, "liberi_oracle.so")
    src = os.path.join(HERE, "eri_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "liberi_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        L.orc_basis_new.restype = ctypes.c_void_p
        L.orc_basis_new.argtypes = [ctypes.c_int, c_ip, c_ip, c_ip, c_ip, c_dp, c_dp, c_dp]
        L.orc_basis_free.argtypes = [ctypes.c_void_p]
        L.orc_basis_nbf.argtypes = [ctypes.c_void_p]
        L.orc_eri_quartet.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 4 + [c_dp]
        L.orc_schwarz.argtypes = [ctypes.c_void_p, c_dp, c_dp]
        L.orc_eri_tensor.restype = ctypes.c_long
        L.orc_eri_tensor.argtypes = [ctypes.c_void_p, ctypes.c_double, c_dp]
        L.orc_jk.argtypes = [ctypes.c_int] + [c_dp] * 7
        L.orc_set_ints_type.argtypes = [ctypes.c_int, ctypes.c_double]
        L.orc_jk_sample.argtypes = [ctypes.c_void_p, ctypes.c_double, c_dp, ctypes.c_int, c_ip, ctypes.c_int, c_ip,
                                    c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, ctypes.c_int, ctypes.POINTER(ctypes.c_long)]
        L.orc_boys_coeff.restype = ctypes.c_double
        L.orc_boys_coeff.argtypes = [ctypes.c_int] * 3
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(c_dp)


def _ip(a):
    return a.ctypes.data_as(c_ip)


class OracleBasis:
    """Wraps a flattened shell table (any object with l, K, is_cart, first_fn, nfn, centres,
    exps, scc numpy attributes -- e.g. pychem_b200.basis_table.BasisTable)."""

    def __init__(self, table):
        self.t = table
        self.L = lib()
        self.h = self.L.orc_basis_new(
            table.nshell, _ip(table.l), _ip(table.K), _ip(table.is_cart), _ip(table.first_fn),
            _dp(table.centres), _dp(table.exps), _dp(table.scc))
        if not self.h:
            raise ValueError("oracle: angular momentum above its limit")
        self.nbf = table.nbf
        self.nshell = table.nshell

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_basis_free(self.h)
            self.h = None

    def quartet(self, a, b, c, d):
        """(ab|cd) block, shape (nfa, nfb, nfc, nfd) -- integrals.two_electron."""
        n = self.t.nfn
        out = np.zeros((n[a], n[b], n[c], n[d]))
        self.L.orc_eri_quartet(self.h, a, b, c, d, _dp(out))
        return out

    def schwarz(self):
        """(bounds[npair,49], pmax[npair]) -- hartree_fock.py:244-254."""
        npair = self.nshell * (self.nshell + 1) // 2
        bounds = np.zeros((npair, 49))
        pmax = np.zeros(npair)
        self.L.orc_schwarz(self.h, _dp(bounds), _dp(pmax))
        return bounds, pmax

    def tensor(self, thresh=1.0e-8):
        """Dense (N,N,N,N) tensor with the reference's screening -- evaluate_2e_ints."""
        N = self.nbf
        G = np.zeros((N, N, N, N))
        nsurv = self.L.orc_eri_tensor(self.h, thresh, _dp(G))
        return G, nsurv


def jk_sample(ob, Dt, Da, Db, j_pairs, k_shells, thresh=1.0e-8, pmax=None, nthreads=None):
    """Sampled J blocks (shell pairs `j_pairs`) and X rows (shells `k_shells`) of
    make_coulomb_exchange_matrices for molecules whose N^4 tensor cannot be stored.  Returns
    (J, Xa, Xb, mask_J, mask_X, nquartets): N x N arrays holding the sampled entries, boolean
    masks of the entries that are filled."""
    N = ob.nbf
    if pmax is None:
        _, pmax = ob.schwarz()
    pmax = np.ascontiguousarray(pmax, dtype=float)
    Dt, Da, Db = (np.ascontiguousarray(x, dtype=float) for x in (Dt, Da, Db))
    jp = np.ascontiguousarray(np.asarray(j_pairs, dtype=np.int32).reshape(-1, 2))
    ks = np.ascontiguousarray(np.asarray(k_shells, dtype=np.int32).ravel())
    J, Xa, Xb = np.zeros((N, N)), np.zeros((N, N)), np.zeros((N, N))
    nq = ctypes.c_long()
    if nthreads is None:
        nthreads = min(32, os.cpu_count() or 1)
    ob.L.orc_jk_sample(ob.h, float(thresh), _dp(pmax), len(jp), _ip(jp), len(ks), _ip(ks), _dp(Dt), _dp(Da), _dp(Db),
                       _dp(J), _dp(Xa), _dp(Xb), int(nthreads), ctypes.byref(nq))
    mJ, mX = np.zeros((N, N), dtype=bool), np.zeros((N, N), dtype=bool)
    f, nf = ob.t.first_fn, ob.t.nfn
    for a, b in jp:
        mJ[f[a]:f[a] + nf[a], f[b]:f[b] + nf[b]] = True
        mJ[f[b]:f[b] + nf[b], f[a]:f[a] + nf[a]] = True
    for a in ks:
        mX[f[a]:f[a] + nf[a], :] = True
    return J, Xa, Xb, mJ, mX, nq.value


def jk(G, Dt, Da, Db):
    """(J, Xa, Xb) with Exchange = -K -- make_coulomb_exchange_matrices."""
    N = G.shape[0]
    G = np.ascontiguousarray(G)
    Dt, Da, Db = (np.ascontiguousarray(x, dtype=float) for x in (Dt, Da, Db))
    J, Xa, Xb = np.zeros((N, N)), np.zeros((N, N)), np.zeros((N, N))
    lib().orc_jk(N, _dp(G), _dp(Dt), _dp(Da), _dp(Db), _dp(J), _dp(Xa), _dp(Xb))
    return J, Xa, Xb


def set_ints_type(ints_type=0, grid_value=-1.0):
    """0 = electron repulsion (default), 1 = scattering fundamentals at `grid_value`
    (two_electron_scattering.c).  Global switch of the C oracle; reset it to 0 after use."""
    lib().orc_set_ints_type(int(ints_type), float(grid_value))
