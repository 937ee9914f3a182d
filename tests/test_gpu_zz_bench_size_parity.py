"""Oracle-pinned parity at the BENCHMARK sizes: integral-direct J/K of (H2O)16 and (H2O)32 in
6-31G** (N = 384 / 768, bench.py's workload) against the CPU oracle.

N^4 cannot be stored at these sizes, so the oracle (oracle/eri_oracle.c: orc_jk_sample -- the
reference's screening rule, Methods/hartree_fock.py:276-295, and its contraction patterns,
:345-347, evaluated block by block with the restated integrals.two_electron) produces a sample of
the outputs: eight Coulomb blocks J[ab] and four full exchange rows X[a, :] (1.2e6 / 3.0e6 shell
quartets).  Bar: 1e-10 absolute.  RHF-shaped symmetric densities (the run kernels / resident
images) and general non-symmetric ones (the NOCI variant).
"""
import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu

TOL = 1.0e-10


@pytest.mark.parametrize("n", [16, 32])
def test_direct_jk_vs_oracle_sampled_rows(n):
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    from oracle import oracle
    from pychem_b200 import engine, structures as S
    mol = S.Molecule(S.water_cluster(n), "6-31G**")
    db = engine.DeviceBasis(mol)
    ob = oracle.OracleBasis(db.table)
    _, pm = db.schwarz()
    _, pm0 = ob.schwarz()
    assert np.abs(pm - pm0).max() < 1e-12
    db.plan(1.0e-8, 0, 1)
    j_pairs, k_shells = helpers.water_cluster_samples(n)
    rng = np.random.default_rng(100 + n)
    N = db.nbf
    X = rng.uniform(-1, 1, (N, N))
    Da = 0.5 * (X + X.T)
    A, B = rng.uniform(-1, 1, (N, N)), rng.uniform(-1, 1, (N, N))
    for (Dt, D1, D2, variant) in ((2 * Da, Da, Da, engine.RHF), (A + B, A, B, engine.GEN)):
        J0, Xa0, Xb0, mJ, mX, nq = oracle.jk_sample(ob, Dt, D1, D2, j_pairs, k_shells, pmax=pm0)
        assert nq > 1000 * n
        for v in (variant, None):                     # explicit variant and the device-side classification
            J, Xa, Xb = db.jk_direct(Dt, D1, D2, variant=v)
            assert np.abs(np.asarray(J) - J0)[mJ].max() < TOL
            assert np.abs(np.asarray(Xa) - Xa0)[mX].max() < TOL
            assert np.abs(np.asarray(Xb) - Xb0)[mX].max() < TOL
    db.close()
