"""Flatten a molecule (ours or the reference's, duck-typed) into the shell table the C-ABI takes.

Function indexing follows the reference exactly: atoms in input order, shells in basis-file
order, a shell's functions contiguous (Util/structures.py:511-520).
"""
import numpy as np


class BasisTable:
    """Plain arrays describing every contracted shell:

    l[s], K[s], is_cart[s], first_fn[s], nfn[s], centres[s,3], prim_off[s], exps[], scc[]
    (scc = cc*(2a)^((l+1.5)/2), Util/structures.py:843).
    """

    def __init__(self, molecule):
        l, K, is_cart, first_fn, nfn, centres, exps, scc = [], [], [], [], [], [], [], []
        count = 0
        for atom in molecule.Atoms:
            for cgtf in atom.Basis:
                ll = int(cgtf.AngularMomentum)
                l.append(ll)
                K.append(int(cgtf.NPrimitives))
                cart = int(ll >= 2 and int(cgtf.NAngMom) == int(cgtf.NAngMomCart))
                is_cart.append(cart)
                first_fn.append(count)
                nfn.append(int(cgtf.NAngMom))
                count += int(cgtf.NAngMom)
                centres.append([float(x) for x in atom.Coordinates])
                exps.extend(float(x) for x in cgtf.Exponents)
                scc.extend(float(x) for x in cgtf.ScaledCCs)
        self.nshell = len(l)
        self.nbf = count
        self.l = np.array(l, dtype=np.int32)
        self.K = np.array(K, dtype=np.int32)
        self.is_cart = np.array(is_cart, dtype=np.int32)
        self.first_fn = np.array(first_fn, dtype=np.int32)
        self.nfn = np.array(nfn, dtype=np.int32)
        self.centres = np.ascontiguousarray(np.array(centres, dtype=np.float64).reshape(-1, 3))
        self.prim_off = np.concatenate([[0], np.cumsum(self.K)]).astype(np.int32)
        self.exps = np.array(exps, dtype=np.float64)
        self.scc = np.array(scc, dtype=np.float64)

    def pair_index(self, a, b):
        """Row of shell pair a<=b in upper-triangular pair order."""
        n = self.nshell
        return a * n - a * (a - 1) // 2 + (b - a)
