"""GPU tests of the one-electron matrices (csrc/pc_one_electron.cuh) against
the reference's own make_core_matrices (goldens from oracle/make_golden.py)."""
import numpy as np
import pytest

from oracle import ref_driver
from tests import helpers

pytestmark = pytest.mark.gpu


def _molecule(name):
    from pychem_b200 import structures as S
    if name == "h2o_cartd":
        return S.Molecule(S.H2O_MONOMER, "6-31G**", cartesian_l=[2])
    return helpers.molecule(name)


@pytest.mark.parametrize("name", ["h2", "lih", "h2o", "h2o_cartd", "h2o2", "benzene"])
def test_core_and_overlap_vs_reference(gold, name):
    from pychem_b200 import integrals as ints_gpu
    g = gold("one_electron.npz")
    mol = _molecule(name)
    core, overlap = ints_gpu.one_electron_matrices(mol)
    ref_core, ref_ov = g[name + "_core"], g[name + "_overlap"]
    assert core.shape == ref_core.shape
    assert np.abs(overlap - ref_ov).max() < 1e-12
    # Core elements reach ~35 Eh (O 1s): 1e-12 relative to the largest element
    assert np.abs(core - ref_core).max() < 1e-12 * max(1.0, np.abs(ref_core).max())
    assert np.abs(core - core.T).max() == 0.0
    # per-shell-pair blocks through the reference's call signature
    sp = mol.ShellPairs[(0, mol.NCgtf - 1)]
    c_blk, s_blk = ints_gpu.one_electron(mol, sp)
    assert np.array_equal(c_blk, core[np.ix_(sp.Centre1.Ivec, sp.Centre2.Ivec)])
    assert np.array_equal(s_blk, overlap[np.ix_(sp.Centre1.Ivec, sp.Centre2.Ivec)])
    ints_gpu.release()


@pytest.mark.skipif(not ref_driver.available(), reason="oracle/_ref (reference copy) not shipped")
def test_full_scf_with_device_one_electron(gold, tmp_path):
    """The reference's driver with make_core_matrices, evaluate_2e_ints and
    make_coulomb_exchange_matrices all rebound: H2O/6-31G** RHF energy within 1e-8 Eh."""
    from pychem_b200 import hartree_fock as hf_gpu, integrals as ints_gpu, structures as S
    ns = ref_driver.modules()
    undo = hf_gpu.install(ns.hartree_fock, one_electron=True)
    try:
        inp = str(tmp_path / "h2o.inp")
        ref_driver.write_input(inp, "h2o", S.H2O_MONOMER, "6-31G**")
        mol = ref_driver.run(inp)
    finally:
        undo()
    g = gold("h2o_631gss.npz")
    assert abs(mol.States[0].TotalEnergy - float(g["energy"])) < 1.0e-8
    hf_gpu.release()
    ints_gpu.release()
