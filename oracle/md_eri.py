"""Independent check values for two-electron integrals: McMurchie-Davidson scheme in plain numpy.

TEST INFRASTRUCTURE ONLY.  A second, algorithmically unrelated evaluation (Hermite expansion
coefficients E_t^{ij}, Hermite Coulomb integrals R_{tuv}, Boys function from scipy's confluent
hypergeometric function) used where the reference itself cannot serve as the yardstick: the
reference's HRR for "goofy" pairs (lower angular momentum first) indexes its base array with
stride `angmom_index(0,0,lb+1)` instead of `angmom_index(0,0,lb+1)+1`
(Methods/c_ints/two_electron_hrr.c:18), which is harmless while the lower shell is s or p -- every
basis with l <= 2, and (s f), (p f) pairs -- and gives wrong integrals for (d f) pairs.

Functions follow the reference's conventions so that blocks are directly comparable with
integrals.two_electron: primitive = cc (2a)^((l+1.5)/2) (Gamma(lx+1/2)Gamma(ly+1/2)Gamma(lz+1/2))^-1/2
x^lx y^ly z^lz exp(-a r^2) (Util/structures.py:843,850-856), Cartesian order lx descending then ly
descending, real spherical combinations of Data/transform_basis.py evaluated with true division.
"""
import math

import numpy as np
from scipy.special import hyp1f1


def boys(m, T):
    return hyp1f1(m + 0.5, m + 1.5, -T) / (2 * m + 1)


def comps(l):
    return [(lx, ly, l - lx - ly) for lx in range(l, -1, -1) for ly in range(l - lx, -1, -1)]


def hermite_E(la, lb, PA, PB, p):
    E = np.zeros((la + 1, lb + 1, la + lb + 2))
    E[0, 0, 0] = 1.0
    for i in range(la):
        for t in range(i + 2):
            E[i + 1, 0, t] = PA * E[i, 0, t] + (t + 1) * E[i, 0, t + 1] + (E[i, 0, t - 1] / (2 * p) if t > 0 else 0.0)
    for i in range(la + 1):
        for j in range(lb):
            for t in range(i + j + 2):
                E[i, j + 1, t] = PB * E[i, j, t] + (t + 1) * E[i, j, t + 1] + (E[i, j, t - 1] / (2 * p) if t > 0 else 0.0)
    return E


def hermite_R(L, alpha, X, Y, Z):
    """R^0_{tuv} as an array [t, u, v], t + u + v <= L."""
    R = np.zeros((L + 1, L + 1, L + 1, L + 1))       # [n, t, u, v]
    T = alpha * (X * X + Y * Y + Z * Z)
    for n in range(L + 1):
        R[n, 0, 0, 0] = (-2 * alpha) ** n * boys(n, T)
    for tot in range(1, L + 1):
        for t in range(tot + 1):
            for u in range(tot - t + 1):
                v = tot - t - u
                for n in range(L - tot + 1):
                    if t > 0:
                        val = (t - 1) * (R[n + 1, t - 2, u, v] if t > 1 else 0.0) + X * R[n + 1, t - 1, u, v]
                    elif u > 0:
                        val = (u - 1) * (R[n + 1, t, u - 2, v] if u > 1 else 0.0) + Y * R[n + 1, t, u - 1, v]
                    else:
                        val = (v - 1) * (R[n + 1, t, u, v - 2] if v > 1 else 0.0) + Z * R[n + 1, t, u, v - 1]
                    R[n, t, u, v] = val
    return R[0]


def _pair_hermite(a, A, la, b, B, lb):
    """Hermite coefficients of every Cartesian component pair: H[ia, ib, t, u, v]."""
    p = a + b
    P = (a * A + b * B) / p
    E = [hermite_E(la, lb, P[k] - A[k], P[k] - B[k], p) for k in range(3)]
    n = la + lb + 1
    H = np.zeros((len(comps(la)), len(comps(lb)), n, n, n))
    for ia, ca in enumerate(comps(la)):
        for ib, cb in enumerate(comps(lb)):
            H[ia, ib] = np.einsum("t,u,v->tuv", E[0][ca[0], cb[0], :n], E[1][ca[1], cb[1], :n], E[2][ca[2], cb[2], :n])
    K = math.exp(-a * b / p * float(np.dot(A - B, A - B)))
    return p, P, K, H


def primitive_cartesian(a, A, la, b, B, lb, c, C, lc, d, D, ld):
    """[ab|cd] over unnormalised Cartesian primitives, shape (ncart a, ncart b, ncart c, ncart d)."""
    p, P, Kab, Hab = _pair_hermite(a, A, la, b, B, lb)
    q, Q, Kcd, Hcd = _pair_hermite(c, C, lc, d, D, ld)
    alpha = p * q / (p + q)
    nb, nk = la + lb + 1, lc + ld + 1
    R = hermite_R(la + lb + lc + ld, alpha, *(P - Q))
    # W[t,u,v,tt,uu,vv] = (-1)^(tt+uu+vv) R[t+tt, u+uu, v+vv]
    W = np.zeros((nb, nb, nb, nk, nk, nk))
    for t in range(nb):
        for u in range(nb):
            for v in range(nb):
                for tt in range(nk):
                    for uu in range(nk):
                        for vv in range(nk):
                            if t + u + v + tt + uu + vv <= la + lb + lc + ld:
                                W[t, u, v, tt, uu, vv] = (-1) ** (tt + uu + vv) * R[t + tt, u + uu, v + vv]
    pref = 2 * math.pi ** 2.5 / (p * q * math.sqrt(p + q)) * Kab * Kcd
    return pref * np.einsum("abtuv,tuvxyz,cdxyz->abcd", Hab, W, Hcd, optimize=True)


def shell_quartet(shells, c2s):
    """Contracted, normalised, spherical block.  shells: four (centre[3], l, exps[K], scaled_ccs[K],
    contraction_scaling[ncart]); c2s: dict l -> (nfn, ncart) matrix."""
    (A, la, ea, ca, na), (B, lb, eb, cb, nb), (C, lc, ec, cc, nc), (D, ld, ed, cd, nd) = shells
    g = 0.0
    for a, wa in zip(ea, ca):
        for b, wb in zip(eb, cb):
            for c, wc in zip(ec, cc):
                for d, wd in zip(ed, cd):
                    g = g + wa * wb * wc * wd * primitive_cartesian(a, np.asarray(A), la, b, np.asarray(B), lb,
                                                                   c, np.asarray(C), lc, d, np.asarray(D), ld)
    mats = [np.asarray(c2s[l], dtype=float) * np.asarray(n)[None, :] for l, n in ((la, na), (lb, nb), (lc, nc), (ld, nd))]
    return np.einsum("abcd,ia,jb,kc,ld->ijkl", g, *mats)
