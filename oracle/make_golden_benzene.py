#!/usr/bin/env python
"""Mint tests/golden/benzene_631gs_rhf_mp2.npz: BASELINE config 3 (benzene, 6-31G*, RHF + MP2)
from the REAL reference (oracle/_ref).

TEST INFRASTRUCTURE ONLY.
  * RHF through the reference's own driver (pychem.main -> hartree_fock.do with its own C
    integrals), once with its stock criterion |dE| < 1e-7 (energy) and once with the criterion at
    1e-11 (energy_tight: the SCF limit itself, see oracle/make_golden_hf_parts.py for why).
  * MP2: the reference's mp2.do is four nested Python loops over N = 96 with three dot products
    each (Methods/mp2.py:43-69, ~1e9 numpy calls): it cannot finish.  The SAME formulas
    (mp2.py:43-94: transform (ib|jq) with the converged MOs, same-spin sums over i >= j, p >= q of
    (A_ipjq - A_iqjp)^2 / (e_i + e_j - e_p - e_q), opposite-spin sum over all i, j, p, q) are
    evaluated here with numpy.einsum on the tensor the reference built
    (molecule.CoulombIntegrals).  Checked against the reference's own mp2.do on H2O/6-31G**
    (tests/golden/h2o_631gss_mp2.npz) by tests/test_oracle.py.
Stored: energies, MP2 components, converged MOs and orbital energies, sampled tensor entries.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_driver  # noqa: E402
from pychem_b200 import structures as S  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
TIGHT = 1.0e-11


def mp2_numpy(G, Ca, Cb, Ea, Eb, na, nb):
    """(Eaa, Eab, Ebb) -- Methods/mp2.py:43-94 with einsum instead of Python loops."""
    N = G.shape[0]

    def transform(Ci, Cp, Cj, Cq, ni, nj):
        T = np.einsum("mi,mnls->inls", Ci[:, :ni], G, optimize=True)
        T = np.einsum("np,inls->ipls", Cp[:, ni:], T, optimize=True)
        T = np.einsum("lj,ipls->ipjs", Cj[:, :nj], T, optimize=True)
        return np.einsum("sq,ipjs->ipjq", Cq[:, nj:], T, optimize=True)       # (i p | j q)

    def same(C, E, no):
        T = transform(C, C, C, C, no, no)
        D = E[:no, None, None, None] - E[None, no:, None, None] + E[None, None, :no, None] - E[None, None, None, no:]
        A = T - T.transpose(0, 3, 2, 1)
        i, p, j, q = np.ogrid[:no, :N - no, :no, :N - no]
        return float(np.sum(np.where((j <= i) & (q <= p), A ** 2 / D, 0.0)))

    Eaa, Ebb = same(Ca, Ea, na), same(Cb, Eb, nb)
    T = transform(Ca, Ca, Cb, Cb, na, nb)
    D = Ea[:na, None, None, None] - Ea[None, na:, None, None] + Eb[None, None, :nb, None] - Eb[None, None, None, nb:]
    return Eaa, float(np.sum(T ** 2 / D)), Ebb


def main():
    ns = ref_driver.modules()
    inp = os.path.join(GOLD, "_benzene.inp")
    ref_driver.write_input(inp, "benzene", S.benzene(), "6-31G*", maxiter=200)
    conv = ns.constants.energy_convergence
    t = time.time()
    try:
        ns.constants.energy_convergence = TIGHT
        tight = ref_driver.run(inp)
        print("tight RHF %.1fs" % (time.time() - t), repr(tight.States[0].TotalEnergy), flush=True)
        ns.constants.energy_convergence = conv
        t = time.time()
        stock = ref_driver.run(inp)
        print("stock RHF %.1fs" % (time.time() - t), repr(stock.States[0].TotalEnergy), flush=True)
    finally:
        ns.constants.energy_convergence = conv
        os.remove(inp)
    st = tight.States[0]
    G = np.asarray(tight.CoulombIntegrals)
    Ca, Cb = np.array(st.Alpha.MOs), np.array(st.Beta.MOs)
    Ea, Eb = np.array(st.Alpha.Energies), np.array(st.Beta.Energies)
    na, nb = int(tight.NAlphaElectrons), int(tight.NBetaElectrons)
    t = time.time()
    Eaa, Eab, Ebb = mp2_numpy(G, Ca, Cb, Ea, Eb, na, nb)
    print("MP2 %.1fs" % (time.time() - t), Eaa, Eab, Ebb, "total", st.TotalEnergy + Eaa + Eab + Ebb)
    np.savez_compressed(os.path.join(GOLD, "benzene_631gs_rhf_mp2.npz"),
                        energy=stock.States[0].TotalEnergy, energy_tight=st.TotalEnergy, tight_convergence=TIGHT,
                        Eaa=Eaa, Eab=Eab, Ebb=Ebb, mp2_total=st.TotalEnergy + Eaa + Eab + Ebb,
                        Ca=Ca, Cb=Cb, Ea=Ea, Eb=Eb, na=na, nb=nb, Da=np.array(st.Alpha.Density),
                        G_sum=G.sum(), G_sq=(G ** 2).sum(), G_sample=G.ravel()[::9973].copy(),
                        core=np.array(tight.Core), overlap=np.array(tight.Overlap))
    print("wrote benzene_631gs_rhf_mp2.npz")


if __name__ == "__main__":
    main()
