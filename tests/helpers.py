"""Shared test helpers: the molecules the golden fixtures were minted on."""
import os

from pychem_b200 import structures as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

H2 = [["H", 1.0, 0.0, 0.0, 0.0], ["H", 1.0, 0.0, 0.0, 1.0]]              # Tests/H2_HF.test.inp
LIH = [["Li", 3.0, 0.0, 0.0, 0.0], ["H", 1.0, 2.2, 0.0, 0.0]]            # Tests/LiH_SFS_NOCI.test.inp
H3 = [["H", 1.0, 0.0, 0.0, 0.0], ["H", 1.0, 1.0, 0.0, 0.0], ["H", 1.0, 0.4, 1.0, 0.0]]  # Tests/example1.inp
# f-shell cases (oracle/make_golden_f.py): four heavy atoms at general positions; hydrogen fluoride
CNON = [["C", 6.0, 0.00, 0.00, 0.00], ["N", 7.0, 1.05, 0.35, -0.20], ["O", 8.0, -0.40, 1.15, 0.55],
        ["N", 7.0, 0.60, -0.85, 1.10]]
HYDROGEN_FLUORIDE = [["F", 9.0, 0.0, 0.0, 0.0], ["H", 1.0, 0.31, 0.42, 0.74]]


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
    if name == "cnon_tz":
        return S.Molecule(CNON, "cc-pVTZ")
    if name == "hf_tz":
        return S.Molecule(HYDROGEN_FLUORIDE, "cc-pVTZ")
    raise KeyError(name)


def f_shell_targets(g):
    """(quartet, target block, source) for every sampled quartet of tests/golden/f_shell_ccpvtz.npz:
    the reference's own block where the reference is right; for (d f) pairs in goofy order, where its
    HRR stride is off by one (Methods/c_ints/two_electron_hrr.c:18), the block of the reference
    with that one expression corrected -- which agrees with independent McMurchie-Davidson values
    to 5e-15 (oracle/make_golden_f.py; tests/test_oracle.py checks the stored evidence)."""
    out = []
    for q, lo, hi, bad in zip(g["quartets"], g["offsets"][:-1], g["offsets"][1:], g["affected"]):
        src = g["fixed_blocks"] if bad else g["blocks"]
        out.append((tuple(int(x) for x in q), src[lo:hi], "reference, HRR stride corrected" if bad else "reference"))
    return out


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


def water_cluster_samples(n):
    """Deterministic sample of outputs for the benchmark-size J/K parity checks on (H2O)_n
    6-31G** (12 shells per water: O s s p s p d, H s s p, H s s p): eight shell pairs (a, b) whose
    Coulomb blocks are compared and four shells whose exchange rows are compared, covering every
    shell type, near and far pairs."""
    w = lambda k, s: 12 * (k % n) + s          # noqa: E731  shell s of water k
    j_pairs = [(w(0, 0), w(0, 0)), (w(0, 5), w(0, 5)), (w(0, 2), w(1, 4)), (w(0, 5), w(1, 7)),
               (w(0, 3), w(5, 5)), (w(3, 8), w(3, 11)), (w(0, 4), w(n - 1, 4)), (w(7, 1), w(9, 5))]
    j_pairs = [(min(a, b), max(a, b)) for a, b in j_pairs]
    k_shells = [w(0, 5), w(0, 4), w(n // 2, 0), w(n - 1, 8)]
    return j_pairs, k_shells
