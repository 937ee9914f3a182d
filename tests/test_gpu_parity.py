"""GPU parity tests (run on the B200 box): the CUDA path, called through the C ABI
(include/pychem_b200.h via pychem_b200.engine), against

  * the golden vectors minted from the reference itself (tests/golden/, oracle/make_golden.py),
  * the CPU oracle (oracle/eri_oracle.c) on seeded inputs it finishes in seconds,
  * size-independent properties at the full BASELINE sizes (stored == direct, linearity,
    symmetry, rank-partition additivity).

Bars (north_star): individual ERIs within 1e-12 absolute (FP64, bit-exact is not defined for
reordered floating-point sums); J/K within 1e-10 absolute; energies within 1e-8 Eh.
"""
import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu

ERI_TOL = 1.0e-12
JK_TOL = 1.0e-10


@pytest.fixture(scope="module")
def eng():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    from pychem_b200 import engine
    return engine


def _basis(eng, name):
    return eng.DeviceBasis(helpers.molecule(name))


# --------------------------------------------------------------------------------------------
# golden vectors from the reference
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,fixture", [("h2o2", "h2o2_631gss.npz"), ("benzene", "benzene_631gs.npz")])
def test_sampled_quartets_all_21_classes(eng, gold, name, fixture):
    g = gold(fixture)
    db = _basis(eng, name)
    blocks = db.eri_quartets(g["quartets"])
    worst = 0.0
    for blk, lo, hi in zip(blocks, g["offsets"][:-1], g["offsets"][1:]):
        assert blk.size == hi - lo
        worst = max(worst, float(np.abs(blk.ravel() - g["blocks"][lo:hi]).max()))
    assert worst < ERI_TOL, worst
    db.close()


@pytest.mark.parametrize("name,fixture", [("h2", "h2_6311g.npz"), ("lih", "lih_631g.npz"),
                                          ("h2o", "h2o_631gss.npz")])
def test_dense_tensor_and_bounds(eng, gold, name, fixture):
    g = gold(fixture)
    db = _basis(eng, name)
    bounds, pmax = db.schwarz()
    assert np.abs(bounds - g["bounds"]).max() < 1e-12
    G_dev, G = db.eri_tensor(1.0e-8, to_host=True)
    assert G.shape == g["G"].shape
    assert np.abs(G - g["G"]).max() < ERI_TOL
    assert np.abs(G_dev.cpu().numpy() - G).max() == 0.0
    db.close()


@pytest.mark.parametrize("fixture,name,keys", [("h2_6311g.npz", "h2", ("",)), ("lih_631g.npz", "lih", ("",)),
                                               ("h2o_631gss.npz", "h2o", ("", "2"))])
def test_jk_stored_and_direct_vs_reference_einsum(eng, gold, fixture, name, keys):
    g = gold(fixture)
    db = _basis(eng, name)
    db.schwarz()
    G_dev, _ = db.eri_tensor(1.0e-8, to_host=False)
    db.plan(1.0e-8, 0, 1)
    for k in keys:
        Dt, Da, Db = g["Dt" + k], g["Da" + k], g["Db" + k]
        ref = (g["J" + k], g["Xa" + k], g["Xb" + k])
        scale = max(1.0, max(np.abs(r).max() for r in ref))
        for got in (db.jk_stored(G_dev, Dt, Da, Db), db.jk_direct(Dt, Da, Db),
                    db.jk_direct(Dt, Da, Db, variant=eng.GEN)):
            for mine, r in zip(got, ref):
                assert np.abs(mine - r).max() < JK_TOL * scale
    db.close()


def test_jk_variants_rhf_uhf(eng, gold):
    g = gold("h2o_631gss.npz")
    db = _basis(eng, "h2o")
    db.plan(1.0e-8, 0, 1)
    G = g["G"]
    rng = np.random.default_rng(3)
    X = rng.uniform(-1, 1, (24, 24)); Da = 0.5 * (X + X.T)
    X = rng.uniform(-1, 1, (24, 24)); Db = 0.5 * (X + X.T)
    # RHF-shaped (Da == Db) and UHF-shaped, against the reference's einsum patterns
    for a, b, variant in ((Da, Da, eng.RHF), (Da, Db, eng.UHF)):
        J = np.einsum("cd,abcd->ab", a + b, G)
        Xa = np.einsum("cb,abcd->ad", -a, G)
        Xb = np.einsum("cb,abcd->ad", -b, G)
        assert eng.classify_densities(a + b, a, b) == variant
        got = db.jk_direct(a + b, a, b)
        for mine, r in zip(got, (J, Xa, Xb)):
            assert np.abs(mine - r).max() < JK_TOL * max(1.0, np.abs(r).max())
    db.close()


# --------------------------------------------------------------------------------------------
# against the CPU oracle on seeded inputs
# --------------------------------------------------------------------------------------------
def test_water_trimer_full_tensor_vs_oracle(eng):
    from oracle import oracle
    from pychem_b200 import structures as S
    rng = np.random.default_rng(42)
    coords = S.water_cluster(3)
    for row in coords:                      # break the lattice symmetry
        for k in (2, 3, 4):
            row[k] += float(rng.uniform(-0.15, 0.15))
    mol = S.Molecule(coords, "6-31G**")
    db = eng.DeviceBasis(mol)
    bounds, pmax = db.schwarz()
    ob = oracle.OracleBasis(db.table)
    ob_bounds, ob_pmax = ob.schwarz()
    assert np.abs(bounds - ob_bounds).max() < 1e-12
    G_dev, G = db.eri_tensor(1.0e-8, to_host=True)
    G_ref, nsurv = ob.tensor(1.0e-8)
    assert np.abs(G - G_ref).max() < ERI_TOL
    # identical screening decisions: the zero patterns agree block by block
    assert np.array_equal(G == 0.0, G_ref == 0.0) or np.abs(G[G_ref == 0.0]).max() < 1e-13
    db.close()


def test_edge_cases(eng):
    """Single shell, single atom, far-apart atoms (asymptotic Boys branch), coincident centres
    (T = 0 branch)."""
    from oracle import oracle
    from pychem_b200 import structures as S
    cases = [
        [["H", 1.0, 0.0, 0.0, 0.0]],
        [["O", 8.0, 0.0, 0.0, 0.0]],
        [["O", 8.0, 0.0, 0.0, 0.0], ["O", 8.0, 0.0, 0.0, 9.0]],
        [["H", 1.0, 0.0, 0.0, 0.0], ["H", 1.0, 0.0, 0.0, 1.0e-9]],
    ]
    for coords in cases:
        mol = S.Molecule(coords, "6-31G**")
        db = eng.DeviceBasis(mol)
        db.schwarz()
        _, G = db.eri_tensor(1.0e-8, to_host=True)
        G_ref, _ = oracle.OracleBasis(db.table).tensor(1.0e-8)
        assert np.abs(G - G_ref).max() < ERI_TOL
        db.close()


def test_errors_are_loud(eng):
    from pychem_b200 import _lib, structures as S
    db = _basis(eng, "h2")
    with pytest.raises(_lib.PychemB200Error):
        db.eri_quartets([(1, 0, 0, 0)])    # a > b
    db.close()


# --------------------------------------------------------------------------------------------
# full-size, size-independent properties: (H2O)8 6-31G** (N = 192)
# --------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def water8(eng):
    from pychem_b200 import structures as S
    db = eng.DeviceBasis(S.Molecule(S.water_cluster(8), "6-31G**"))
    db.schwarz()
    yield db
    db.close()


def _sym(rng, n):
    X = rng.uniform(-1, 1, (n, n))
    return 0.5 * (X + X.T)


def test_full_size_direct_equals_stored(eng, water8):
    db = water8
    N = db.nbf
    assert N == 192
    G_dev, _ = db.eri_tensor(1.0e-8, to_host=False)
    db.plan(1.0e-8, 0, 1)
    rng = np.random.default_rng(7)
    Da, Db = _sym(rng, N), _sym(rng, N)
    stored = db.jk_stored(G_dev, Da + Db, Da, Db)
    direct = db.jk_direct(Da + Db, Da, Db)
    for s, d in zip(stored, direct):
        assert np.abs(s - d).max() < 1e-9 * max(1.0, np.abs(s).max())
    # non-symmetric (NOCI-shaped) densities through the general variant
    A, B = rng.uniform(-1, 1, (N, N)), rng.uniform(-1, 1, (N, N))
    stored = db.jk_stored(G_dev, A + B, A, B)
    direct = db.jk_direct(A + B, A, B)
    for s, d in zip(stored, direct):
        assert np.abs(s - d).max() < 1e-9 * max(1.0, np.abs(s).max())
    del G_dev


def test_full_size_linearity_symmetry_partition(eng, water8):
    db = water8
    N = db.nbf
    rng = np.random.default_rng(8)
    D1, D2 = _sym(rng, N), _sym(rng, N)
    db.plan(1.0e-8, 0, 1)
    J1, X1, _ = db.jk_direct(2 * D1, D1, D1)
    J2, X2, _ = db.jk_direct(2 * D2, D2, D2)
    J12, X12, _ = db.jk_direct(2 * (D1 + D2), D1 + D2, D1 + D2)
    scale = np.abs(J12).max()
    assert np.abs(J1 + J2 - J12).max() < 1e-10 * scale
    assert np.abs(X1 + X2 - X12).max() < 1e-10 * scale
    assert np.abs(J12 - J12.T).max() < 1e-10 * scale
    assert np.abs(X12 - X12.T).max() < 1e-10 * scale
    # energy-like invariant: sum(D1*J(D2)) == sum(D2*J(D1))
    assert abs(np.sum(D1 * J2) - np.sum(D2 * J1)) < 1e-9 * abs(np.sum(D1 * J2))
    # static partition: the accumulators of 3 ranks add up to the 1-rank result
    import torch
    from pychem_b200 import _lib, dist
    total = torch.zeros(3 * N * N, dtype=torch.float64, device="cuda")
    quartets = 0
    for r in range(3):
        c = db.plan(1.0e-8, r, 3)
        quartets += c["my_quartets"]
        acc = db.accumulator()
        Dt1 = np.ascontiguousarray(2 * D1)
        _lib.check(db.lib.pc_jk_direct_accumulate(db.h, eng.RHF, eng._ptr(Dt1), eng._ptr(D1), eng._ptr(D1), eng._ptr(acc)))
        torch.cuda.synchronize()
        total += acc
    assert quartets == c["all_quartets"]
    J, Xa, _ = dist.finalize_accumulators(total.cpu().numpy(), N, eng.RHF)
    assert np.abs(J - J1).max() < 1e-10 * scale
    assert np.abs(Xa - X1).max() < 1e-10 * scale
    db.plan(1.0e-8, 0, 1)


# --------------------------------------------------------------------------------------------
# other basis sets / elements: contraction depths and shell mixes the water tests do not have
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("basis,coords", [
    ("STO-3G", [["C", 6.0, 0.0, 0.0, 0.0], ["O", 8.0, 0.0, 0.0, 1.13]]),
    ("3-21G", [["N", 7.0, 0.0, 0.0, 0.12], ["H", 1.0, 0.0, 0.94, -0.27], ["H", 1.0, 0.81, -0.47, -0.27],
               ["H", 1.0, -0.81, -0.47, -0.27]]),
    ("6-311G**", [["F", 9.0, 0.0, 0.0, 0.0], ["H", 1.0, 0.0, 0.0, 0.92]]),
    ("cc-pVDZ", [["O", 8.0, 0.0, 0.0, 0.117790], ["H", 1.0, 0.0, 0.755453, -0.471161],
                 ["H", 1.0, 0.0, -0.755453, -0.471161]]),
    ("6-31G*", [["Li", 3.0, 0.0, 0.0, 0.0], ["F", 9.0, 0.0, 0.0, 1.56]]),
])
def test_other_basis_sets_vs_oracle(eng, basis, coords):
    from oracle import oracle
    from pychem_b200 import structures as S
    mol = S.Molecule(coords, basis)
    db = eng.DeviceBasis(mol)
    bounds, _ = db.schwarz()
    ob = oracle.OracleBasis(db.table)
    ob_bounds, _ = ob.schwarz()
    assert np.abs(bounds - ob_bounds).max() < 1e-12
    G_dev, G = db.eri_tensor(1.0e-8, to_host=True)
    G_ref, _ = ob.tensor(1.0e-8)
    assert np.abs(G - G_ref).max() < ERI_TOL
    rng = np.random.default_rng(9)
    N = db.nbf
    A, B = rng.uniform(-1, 1, (N, N)), rng.uniform(-1, 1, (N, N))
    ref = oracle.jk(G_ref, A + B, A, B)
    db.plan(1.0e-8, 0, 1)
    for got in (db.jk_stored(G_dev, A + B, A, B), db.jk_direct(A + B, A, B)):
        for mine, r in zip(got, ref):
            assert np.abs(mine - r).max() < JK_TOL * max(1.0, np.abs(r).max())
    db.close()


def test_cartesian_d_shells(eng, gold):
    """Cartesian_L = [2] (Util/structures.py:844-849): six Cartesian d functions per shell; golden
    tensor from the reference run with that keyword, plus a water dimer against the oracle."""
    from oracle import oracle
    from pychem_b200 import structures as S
    g = gold("h2o_631gss_cartd.npz")
    mol = S.Molecule(S.H2O_MONOMER, "6-31G**", cartesian_l=[2])
    db = eng.DeviceBasis(mol)
    assert db.nbf == 25
    db.schwarz()
    G_dev, G = db.eri_tensor(1.0e-8, to_host=True)
    assert np.abs(G - g["G"]).max() < ERI_TOL
    db.close()
    mol = S.Molecule(S.water_cluster(2), "6-31G**", cartesian_l=[2])
    db = eng.DeviceBasis(mol)
    bounds, _ = db.schwarz()
    ob = oracle.OracleBasis(db.table)
    assert np.abs(bounds - ob.schwarz()[0]).max() < 1e-12
    G_dev, G = db.eri_tensor(1.0e-8, to_host=True)
    G_ref, _ = ob.tensor(1.0e-8)
    assert np.abs(G - G_ref).max() < ERI_TOL
    rng = np.random.default_rng(2)
    N = db.nbf
    A, B = rng.uniform(-1, 1, (N, N)), rng.uniform(-1, 1, (N, N))
    ref = oracle.jk(G_ref, A + B, A, B)
    db.plan(1.0e-8, 0, 1)
    for got in (db.jk_stored(G_dev, A + B, A, B), db.jk_direct(A + B, A, B)):
        for mine, r in zip(got, ref):
            assert np.abs(mine - r).max() < JK_TOL * max(1.0, np.abs(r).max())
    Ds = 0.5 * (A + A.T)
    ref = oracle.jk(G_ref, 2 * Ds, Ds, Ds)
    for mine, r in zip(db.jk_direct(2 * Ds, Ds, Ds), ref):
        assert np.abs(mine - r).max() < JK_TOL * max(1.0, np.abs(r).max())
    db.close()


# --------------------------------------------------------------------------------------------
# scattering fundamentals (ints_type = 1): same kernels, other fundamentals
# --------------------------------------------------------------------------------------------
def _unpack(packed, n):
    a, b = np.tril_indices(n)
    p, q = np.tril_indices(len(a))
    G = np.zeros((n,) * 4)
    ia, ib, ic, id_ = a[p], b[p], a[q], b[q]
    for x in ((ia, ib, ic, id_), (ib, ia, ic, id_), (ia, ib, id_, ic), (ib, ia, id_, ic),
              (ic, id_, ia, ib), (id_, ic, ia, ib), (ic, id_, ib, ia), (id_, ic, ib, ia)):
        G[x] = packed
    return G


def test_scattering_tensor_bounds_intensity_vs_reference(eng, gold):
    """H2O 6-31G** at S = 0, 0.5, 2, 7.5 against the reference's evaluate_2e_ints(molecule, 1, S)
    (golden: oracle/make_golden_scattering.py) and the printed intensities of properties.py."""
    from pychem_b200 import hartree_fock as hf_gpu, integrals as ints_gpu, properties as prop_gpu
    g = gold("h2o_631gss_scattering.npz")
    mol = helpers.molecule("h2o")

    class Spin:
        pass

    class State:
        Total, Alpha, Beta = Spin(), Spin(), Spin()
    State.Total.Density, State.Alpha.Density, State.Beta.Density = g["scf_Dt"], g["scf_Da"], g["scf_Db"]
    try:
        for k, S in enumerate(g["grid"]):
            hf_gpu.evaluate_2e_ints(mol, 1, float(S))
            ref = _unpack(g["G%d" % k], mol.NOrbitals)
            assert np.abs(mol.CoulombIntegrals - ref).max() < ERI_TOL
            refb = helpers.bounds_from_flat(ints_gpu.device_basis(mol).table, g["bounds%d" % k])
            for (a, b), blk in refb.items():
                assert np.abs(np.asarray(mol.Bounds[a][b]) - blk).max() < 1e-12
            val = prop_gpu.scattering_intensity(mol, State, float(S), evaluate=False)
            assert abs(val - g["intensity"][k]) < 1e-8
        # back to the repulsion integrals on the same handle: Schwarz/plan are rebuilt
        hf_gpu.evaluate_2e_ints(mol)
        assert np.abs(mol.CoulombIntegrals - gold("h2o_631gss.npz")["G"]).max() < ERI_TOL
    finally:
        hf_gpu.release()
        ints_gpu.release()


def test_scattering_quartets_all_classes_vs_oracle(eng):
    """Every class at a few grid values (all three regimes of the z^-m j_m(z) evaluation:
    series z < 10, recursion, asymptotic z > 100) against the CPU oracle."""
    from oracle import oracle
    mol = helpers.molecule("h2o2")
    db = eng.DeviceBasis(mol)
    ob = oracle.OracleBasis(db.table)
    rng = np.random.default_rng(5)
    ns = db.table.nshell
    quartets = []
    for _ in range(400):
        a, b, c, d = rng.integers(0, ns, 4)
        quartets.append((min(a, b), max(a, b), min(c, d), max(c, d)))
    try:
        for S in (0.0, 1.0e-3, 0.3, 3.0, 12.0, 40.0):
            db.set_ints_type(1, S)
            oracle.set_ints_type(1, S)
            blocks = db.eri_quartets(quartets)
            worst = 0.0
            for q, blk in zip(quartets, blocks):
                worst = max(worst, float(np.abs(blk - ob.quartet(*[int(x) for x in q])).max()))
            assert worst < ERI_TOL, (S, worst)
    finally:
        oracle.set_ints_type(0, -1.0)
        db.close()


def test_scattering_refuses_direct_digestion(eng):
    db = _basis(eng, "h2")
    db.set_ints_type(1, 0.5)
    db.plan()
    D = np.eye(db.nbf)
    with pytest.raises(RuntimeError):
        db.jk_direct(D, D, D)
    db.close()
