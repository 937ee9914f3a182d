"""Shared test helpers: the molecules the golden fixtures were minted on."""
import os

from pychem_b200 import structures as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

H2 = [["H", 1.0, 0.0, 0.0, 0.0], ["H", 1.0, 0.0, 0.0, 1.0]]              # Tests/H2_HF.test.inp
LIH = [["Li", 3.0, 0.0, 0.0, 0.0], ["H", 1.0, 2.2, 0.0, 0.0]]            # Tests/LiH_SFS_NOCI.test.inp
H3 = [["H", 1.0, 0.0, 0.0, 0.0], ["H", 1.0, 1.0, 0.0, 0.0], ["H", 1.0, 0.4, 1.0, 0.0]]  # Tests/example1.inp


def molecule(name):
    if name == "h2":
        return S.Molecule(H2, "6-311G")
    if name == "lih":
        return S.Molecule(LIH, "6-31G")
    if name == "h3":
        return S.Molecule(H3, "STO-3G", multiplicity=2)
    if name == "h2o":
        return S.Molecule(S.H2O_MONOMER, "6-31G**")
    if name == "h2o2":
        return S.Molecule(S.water_cluster(2), "6-31G**")
    if name == "benzene":
        return S.Molecule(S.benzene(), "6-31G*")
    raise KeyError(name)


def bounds_from_flat(table, flat):
    """[npair,49] fixture layout -> dict[(a,b)] -> (nfa,nfb) array."""
    out = {}
    p = 0
    for a in range(table.nshell):
        for b in range(a, table.nshell):
            na, nb = int(table.nfn[a]), int(table.nfn[b])
            out[(a, b)] = flat[p, :na * nb].reshape(na, nb)
            p += 1
    return out
