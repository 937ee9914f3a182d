"""CPU tests: the oracle restatement (oracle/eri_oracle.c) against the golden vectors minted
from the reference's own C extension + Python driver (oracle/make_golden.py), and -- when
oracle/_ref is present (authoring container / shipped snapshot) -- against the reference live.
Bar: 1e-12 absolute per ERI (north_star), in practice ~1e-14.
"""
import numpy as np
import pytest

from oracle import oracle, ref_driver
from pychem_b200.basis_table import BasisTable
from tests import helpers

ERI_TOL = 1.0e-12


@pytest.mark.parametrize("name,fixture", [("h2", "h2_6311g.npz"), ("lih", "lih_631g.npz"),
                                          ("h2o", "h2o_631gss.npz")])
def test_full_tensor_matches_reference(gold, name, fixture):
    g = gold(fixture)
    ob = oracle.OracleBasis(BasisTable(helpers.molecule(name)))
    G, _ = ob.tensor(1.0e-8)
    assert G.shape == g["G"].shape
    assert np.abs(G - g["G"]).max() < ERI_TOL
    # zeros of the reference (screened / symmetry) stay tiny here
    zero = g["G"] == 0.0
    assert not zero.any() or np.abs(G[zero]).max() < ERI_TOL


@pytest.mark.parametrize("name,fixture", [("h2", "h2_6311g.npz"), ("lih", "lih_631g.npz"),
                                          ("h2o", "h2o_631gss.npz")])
def test_schwarz_bounds(gold, name, fixture):
    g = gold(fixture)
    tb = BasisTable(helpers.molecule(name))
    bounds, pmax = oracle.OracleBasis(tb).schwarz()
    assert np.abs(bounds - g["bounds"]).max() < 1e-12
    assert np.allclose(pmax, g["bounds"].max(axis=1), rtol=0, atol=1e-12)


@pytest.mark.parametrize("name,fixture", [("h2o2", "h2o2_631gss.npz"), ("benzene", "benzene_631gs.npz")])
def test_sampled_quartets_all_classes(gold, name, fixture):
    g = gold(fixture)
    tb = BasisTable(helpers.molecule(name))
    ob = oracle.OracleBasis(tb)
    classes = set()
    for (a, b, c, d), lo, hi in zip(g["quartets"], g["offsets"][:-1], g["offsets"][1:]):
        blk = ob.quartet(int(a), int(b), int(c), int(d)).ravel()
        assert blk.size == hi - lo
        assert np.abs(blk - g["blocks"][lo:hi]).max() < ERI_TOL
        classes.add(tuple(sorted([tuple(sorted((tb.l[a], tb.l[b]))), tuple(sorted((tb.l[c], tb.l[d])))])))
    assert len(classes) == 21          # every (pair class, pair class) combination with l <= 2


@pytest.mark.parametrize("fixture,keys", [("h2_6311g.npz", ("",)), ("lih_631g.npz", ("",)),
                                          ("h2o_631gss.npz", ("", "2"))])
def test_jk_matches_reference_einsum(gold, fixture, keys):
    g = gold(fixture)
    for k in keys:
        J, Xa, Xb = oracle.jk(g["G"], g["Dt" + k], g["Da" + k], g["Db" + k])
        for mine, ref in ((J, g["J" + k]), (Xa, g["Xa" + k]), (Xb, g["Xb" + k])):
            assert np.abs(mine - ref).max() < 1e-11 * max(1.0, np.abs(ref).max())


def test_boys_table_is_cubic_taylor():
    """Table coefficients reproduce F_m at interval centres (erf closed form for m=0)."""
    from math import erf, pi, sqrt
    L = oracle.lib()
    d = 0.002
    for j in (0, 1, 17, 1234, 7749):
        T = (2 * j + 1) * d
        sT = T / (2 * d)
        f = sum(L.orc_boys_coeff(k, 0, j) * sT ** k for k in range(4))
        assert abs(f - 0.5 * sqrt(pi / T) * erf(sqrt(T))) < 2e-15


@pytest.mark.skipif(not ref_driver.available(), reason="oracle/_ref not built")
def test_live_reference_quartets():
    """Same shell quartets through the reference's integrals.two_electron, live."""
    ns = ref_driver.modules()
    from pychem_b200 import structures as S
    mol, _ = ref_driver.build_molecule(S.water_cluster(2), "6-31G**")
    tb = BasisTable(mol)
    ob = oracle.OracleBasis(tb)
    rng = np.random.default_rng(5)
    for _ in range(40):
        a, b, c, d = (int(x) for x in rng.integers(0, tb.nshell, 4))
        a, b = min(a, b), max(a, b)
        c, d = min(c, d), max(c, d)
        ref = ns.integrals.two_electron(mol.ShellPairs[(a, b)], mol.ShellPairs[(c, d)], 0, -1.0)
        assert np.abs(ob.quartet(a, b, c, d) - ref).max() < ERI_TOL


def _unpack(packed, n):
    """Golden scattering tensors hold one value per canonical (ab|cd), a>=b, c>=d, ab>=cd."""
    a, b = np.tril_indices(n)
    p, q = np.tril_indices(len(a))
    G = np.zeros((n,) * 4)
    ia, ib, ic, id_ = a[p], b[p], a[q], b[q]
    for x in ((ia, ib, ic, id_), (ib, ia, ic, id_), (ia, ib, id_, ic), (ib, ia, id_, ic),
              (ic, id_, ia, ib), (id_, ic, ia, ib), (ic, id_, ib, ia), (id_, ic, ib, ia)):
        G[x] = packed
    return G


def test_scattering_integrals_match_reference(gold):
    """ints_type = 1 (two_electron_scattering.c, spherical_bessel_j.c) at S = 0, 0.5, 2, 7.5:
    tensor, Schwarz factors and the scattering intensity of the RHF state (properties.py:19-23)."""
    g = gold("h2o_631gss_scattering.npz")
    tb = BasisTable(helpers.molecule("h2o"))
    ob = oracle.OracleBasis(tb)
    try:
        for k, S in enumerate(g["grid"]):
            oracle.set_ints_type(1, float(S))
            G, _ = ob.tensor(1.0e-8)
            ref = _unpack(g["G%d" % k], tb.nbf)
            assert np.abs(G - ref).max() < ERI_TOL
            bounds, _ = ob.schwarz()
            assert np.abs(bounds - g["bounds%d" % k]).max() < 1e-12
            Dt, Da, Db = g["scf_Dt"], g["scf_Da"], g["scf_Db"]
            J, Xa, Xb = oracle.jk(G, Dt, Da, Db)
            val = 10 + (Dt * J).sum() + (Da * Xa).sum() + (Db * Xb).sum()
            assert abs(val - g["intensity"][k]) < 1e-9
        assert abs(g["intensity"][0] - 100.0) < 1e-9      # S = 0: N_el^2
    finally:
        oracle.set_ints_type(0, -1.0)
