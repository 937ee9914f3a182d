"""GPU tests of the MP2 AO->MO transform on the FP64 tensor cores (csrc/pc_mp2.cu) behind the
reference's mp2.do surface.  Goldens: reference's own mp2.do (oracle/make_golden.py).
Bar: total MP2 energies within 1e-8 Eh (north_star)."""
import ctypes
import os

import numpy as np
import pytest

from oracle import ref_driver
from tests import helpers

pytestmark = pytest.mark.gpu
E_TOL = 1.0e-8


def test_dgemm_dmma_matches_numpy():
    import torch
    from pychem_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(0)
    for M, N, K in ((64, 64, 16), (5, 7, 3), (130, 257, 96), (21, 884736 // 64, 96)):
        A = torch.from_numpy(rng.uniform(-1, 1, (M, K))).cuda()
        B = torch.from_numpy(rng.uniform(-1, 1, (K, N))).cuda()
        C = torch.empty((M, N), dtype=torch.float64, device="cuda")
        _lib.check(lib.pc_dgemm_dmma(0, M, N, K, ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(B.data_ptr()),
                                     ctypes.c_void_p(C.data_ptr())), mp2=True)
        ref = A.cpu().numpy() @ B.cpu().numpy()
        assert np.abs(C.cpu().numpy() - ref).max() < 1e-12 * K


class _M:
    pass


def _state(g):
    st = _M()
    st.Alpha, st.Beta = _M(), _M()
    st.Alpha.MOs, st.Beta.MOs = g["Ca"], g["Cb"]
    st.Alpha.Energies, st.Beta.Energies = g["Ea"], g["Eb"]
    st.TotalEnergy = float(g["hf"])
    return st


def test_h2o_mp2_sums_vs_reference(gold):
    """Restricted H2O/6-31G**: same MOs and orbital energies as the reference run, integrals from
    the CUDA path; HF + MP2 total against the reference's mp2.do output."""
    from pychem_b200 import hartree_fock as hf_gpu, integrals as ints_gpu, mp2 as mp2_gpu
    g = gold("h2o_631gss_mp2.npz")
    mol = helpers.molecule("h2o")
    mol.NAlphaElectrons, mol.NBetaElectrons = int(g["na"]), int(g["nb"])
    hf_gpu.evaluate_2e_ints(mol)
    st = _state(g)
    Eaa, Eab, Ebb = mp2_gpu.mp2_sums(mol, st)
    assert abs(st.TotalEnergy + Eaa + Eab + Ebb - float(g["mp2_total"])) < E_TOL
    assert abs(Eaa - Ebb) < 1e-12                   # restricted orbitals
    Eaa0, Eab0, Ebb0 = mp2_gpu.mp2_sums(mol, st, same_spin=False)
    assert Eaa0 == 0.0 and Ebb0 == 0.0 and abs(Eab0 - Eab) < 1e-13
    hf_gpu.release()
    ints_gpu.release()


def test_h3_open_shell_mp2_sums(gold):
    """Tests/example1.inp (H3, STO-3G, CUHF, doublet): unrestricted sums with na != nb."""
    from pychem_b200 import hartree_fock as hf_gpu, integrals as ints_gpu, mp2 as mp2_gpu
    g = gold("h3_sto3g_mp2.npz")
    mol = helpers.molecule("h3")
    assert (mol.NAlphaElectrons, mol.NBetaElectrons) == (int(g["na"]), int(g["nb"]))
    hf_gpu.evaluate_2e_ints(mol)
    st = _state(g)
    Eaa, Eab, Ebb = mp2_gpu.mp2_sums(mol, st)
    assert abs(st.TotalEnergy + Eaa + Eab + Ebb - float(g["mp2_total"])) < E_TOL
    hf_gpu.release()
    ints_gpu.release()


@pytest.mark.skipif(not ref_driver.available(), reason="oracle/_ref (reference copy) not shipped")
def test_dropin_mp2_through_reference_driver(gold):
    """pychem.main on Tests/example1.inp with evaluate_2e_ints, make_coulomb_exchange_matrices AND
    mp2.do rebound to the CUDA path: the line the reference writes to <section>.out must match."""
    from pychem_b200 import hartree_fock as hf_gpu, integrals as ints_gpu, mp2 as mp2_gpu
    ns = ref_driver.modules()
    undo1 = hf_gpu.install(ns.hartree_fock)
    undo2 = mp2_gpu.install(ns.mp2)
    try:
        mol = ref_driver.run(os.path.join(ref_driver.REF_ROOT, "Tests", "example1.inp"))
    finally:
        undo1()
        undo2()
    g = gold("h3_sto3g_mp2.npz")
    emp2 = [float(l.split()[-1]) for l in mol.OutText.splitlines() if "Total MP2 energy" in l][0]
    assert abs(emp2 - float(g["mp2_total"])) < E_TOL
    hf_gpu.release()
    ints_gpu.release()


def test_benzene_mp2_transform_vs_numpy():
    """Benzene 6-31G* (N = 96, the BASELINE MP2 config): the DMMA transform + energy sums against
    an O(N^5) numpy evaluation of the same formulas on the same (GPU-built) tensor, with
    synthetic orthonormal orbitals."""
    from pychem_b200 import hartree_fock as hf_gpu, integrals as ints_gpu, mp2 as mp2_gpu
    mol = helpers.molecule("benzene")
    hf_gpu.evaluate_2e_ints(mol)
    G = mol.CoulombIntegrals
    N = mol.NOrbitals
    rng = np.random.default_rng(5)
    C, _ = np.linalg.qr(rng.uniform(-1, 1, (N, N)))
    E = np.sort(rng.uniform(-2, 2, N))
    E[mol.NAlphaElectrons:] += 3.0                 # keep the denominators away from zero
    st = _M()
    st.Alpha, st.Beta = _M(), _M()
    st.Alpha.MOs = st.Beta.MOs = C
    st.Alpha.Energies = st.Beta.Energies = E
    Eaa, Eab, Ebb = mp2_gpu.mp2_sums(mol, st)
    no = mol.NAlphaElectrons
    Co, Cv = C[:, :no], C[:, no:]
    T = np.einsum("mi,mnls->inls", Co, G, optimize=True)
    T = np.einsum("np,inls->ipls", Cv, T, optimize=True)
    T = np.einsum("lj,ipls->ipjs", Co, T, optimize=True)
    T = np.einsum("sq,ipjs->ipjq", Cv, T, optimize=True)
    D = E[:no, None, None, None] - E[None, no:, None, None] + E[None, None, :no, None] - E[None, None, None, no:]
    Eab_ref = float(np.sum(T ** 2 / D))
    A = T - T.transpose(0, 3, 2, 1)
    i, p, j, q = np.ogrid[:no, :N - no, :no, :N - no]
    Eaa_ref = float(np.sum(np.where((j <= i) & (q <= p), A ** 2 / D, 0.0)))
    assert abs(Eab - Eab_ref) < 1e-10 * max(1.0, abs(Eab_ref))
    assert abs(Eaa - Eaa_ref) < 1e-10 * max(1.0, abs(Eaa_ref))
    assert abs(Ebb - Eaa) < 1e-12 * max(1.0, abs(Eaa))
    hf_gpu.release()
    ints_gpu.release()


def test_benzene_mp2_sums_at_reference_orbitals(gold):
    """BASELINE config 3, MP2 part: the reference's converged benzene/6-31G* orbitals and orbital
    energies (tests/golden/benzene_631gs_rhf_mp2.npz, oracle/make_golden_benzene.py), integrals and
    DMMA transform from the device, against the reference's formulas evaluated with numpy on the
    reference's own tensor."""
    from pychem_b200 import hartree_fock as hf_gpu, integrals as ints_gpu, mp2 as mp2_gpu
    g = gold("benzene_631gs_rhf_mp2.npz")
    mol = helpers.molecule("benzene")
    assert (mol.NAlphaElectrons, mol.NBetaElectrons) == (int(g["na"]), int(g["nb"]))
    hf_gpu.evaluate_2e_ints(mol)
    G = np.asarray(mol.CoulombIntegrals)
    assert np.abs(G.ravel()[::9973] - g["G_sample"]).max() < 1e-12
    assert abs(G.sum() - float(g["G_sum"])) < 1e-7
    st = _state({"Ca": g["Ca"], "Cb": g["Cb"], "Ea": g["Ea"], "Eb": g["Eb"], "hf": g["energy_tight"]})
    Eaa, Eab, Ebb = mp2_gpu.mp2_sums(mol, st)                 # restricted shortcut (Ca == Cb)
    for mine, key in ((Eaa, "Eaa"), (Eab, "Eab"), (Ebb, "Ebb")):
        assert abs(mine - float(g[key])) < 1e-10
    assert abs(st.TotalEnergy + Eaa + Eab + Ebb - float(g["mp2_total"])) < E_TOL
    # the general (three-transform) path on the same orbitals: beta perturbed in the last bits only
    st.Beta.MOs = g["Cb"] * (1.0 + 1e-16)
    st.Beta.MOs[0, 0] = np.nextafter(st.Beta.MOs[0, 0], 1.0)
    Eaa2, Eab2, Ebb2 = mp2_gpu.mp2_sums(mol, st)
    assert abs(Eaa2 - Eaa) < 1e-10 and abs(Eab2 - Eab) < 1e-10 and abs(Ebb2 - Ebb) < 1e-10
    hf_gpu.release()
    ints_gpu.release()


@pytest.mark.skipif(not ref_driver.available(), reason="oracle/_ref (reference copy) not shipped")
def test_benzene_rhf_mp2_dropin_through_reference_driver(gold, tmp_path):
    """BASELINE config 3 end to end: pychem.main on benzene / 6-31G* / MP2 with the hot functions,
    the one-electron matrices and mp2.do rebound to the CUDA path.  RHF total energy and the MP2
    total the driver prints, within 1e-8 Eh of the reference.

    The SCF starts from the reference's converged orbitals (SCF_Guess = "READ", the driver's own
    option, hartree_fock.py:35-41): benzene's core guess puts a degenerate pair at the Fermi level,
    which of the two gets occupied is decided by rounding noise, and the reference's erratic DIIS
    then lands on different stationary points (-253.0426 with its own integrals, -253.1089 here
    from the core guess, integrals equal to 1e-14).  From the reference's orbitals the device
    integrals must keep the SCF where it is: same stationary point, same energy, then MP2."""
    from pychem_b200 import hartree_fock as hf_gpu, integrals as ints_gpu, mp2 as mp2_gpu
    from pychem_b200 import structures as S
    g = gold("benzene_631gs_rhf_mp2.npz")
    ns = ref_driver.modules()
    undo1 = hf_gpu.install(ns.hartree_fock, one_electron=True)
    undo2 = mp2_gpu.install(ns.mp2)
    conv = ns.constants.energy_convergence
    try:
        inp = str(tmp_path / "benzene.inp")
        ref_driver.write_input(inp, "benzene", S.benzene(), "6-31G*", method="MP2", maxiter=200,
                               extra='SCF_Guess = "READ"\nMO_Read_Basis = "6-31G*"\nMO_Read_State = [0]')
        ns.constants.energy_convergence = float(g["tight_convergence"])
        mol = ref_driver.run(inp, files={"631GS_0.alpha_MOs": g["Ca"], "631GS_0.beta_MOs": g["Cb"]})
    finally:
        ns.constants.energy_convergence = conv
        undo1()
        undo2()
    assert abs(mol.States[0].TotalEnergy - float(g["energy_tight"])) < E_TOL
    emp2 = [float(l.split()[-1]) for l in mol.OutText.splitlines() if "Total MP2 energy" in l][0]
    assert abs(emp2 - float(g["mp2_total"])) < E_TOL
    hf_gpu.release()
    ints_gpu.release()
