"""GPU tests of the batched co-density J/K (include/pychem_b200.h
pc_jk_stored_batch / pc_jk_direct_batch) and of the batched NOCI driver mirror
(pychem_b200.noci, Methods/noci.py:46-122).  The same checks run on the CPU over the emulated
kernels in tests/test_emu_cpu.py; this file is named to run after the other GPU files."""
import os

import numpy as np
import pytest

from oracle import ref_driver
from tests import helpers

pytestmark = pytest.mark.gpu

JK_TOL = 1.0e-10
E_TOL = 1.0e-8


def _general_sets(rng, nset, N):
    D = np.empty((nset, 3, N, N))
    for s in range(nset):
        D[s, 1] = rng.uniform(-1, 1, (N, N))
        D[s, 2] = rng.uniform(-1, 1, (N, N))
        D[s, 0] = D[s, 1] + D[s, 2]
    return D


@pytest.mark.parametrize("name,fixture,nset", [("h2o", "h2o_631gss.npz", 5), ("lih", "lih_631g.npz", 6),
                                               ("h2", "h2_6311g.npz", 1)])
def test_batched_jk_vs_reference_einsum(gold, name, fixture, nset):
    import torch
    assert torch.cuda.is_available()
    from pychem_b200 import engine
    g = gold(fixture)
    G = g["G"]
    N = G.shape[0]
    db = engine.DeviceBasis(helpers.molecule(name))
    D = _general_sets(np.random.default_rng(11), nset, N)
    ref = np.empty_like(D)
    for s in range(nset):
        ref[s, 0] = np.einsum("cd,abcd->ab", D[s, 0], G)
        ref[s, 1] = np.einsum("cb,abcd->ad", -D[s, 1], G)
        ref[s, 2] = np.einsum("cb,abcd->ad", -D[s, 2], G)
    scale = max(1.0, np.abs(ref).max())
    db.schwarz()
    G_dev, _ = db.eri_tensor(1.0e-8, to_host=False)
    assert np.abs(db.jk_stored_batch(G_dev, D) - ref).max() < JK_TOL * scale
    db.plan(1.0e-8, 0, 1)
    direct = db.jk_direct_batch(D)
    assert np.abs(direct - ref).max() < JK_TOL * scale
    single = db.jk_direct(D[0, 0], D[0, 1], D[0, 2], variant=engine.GEN)
    for k in range(3):
        assert np.abs(direct[0, k] - single[k]).max() < 1e-12 * scale
    # device-resident inputs and outputs
    D_dev = torch.from_numpy(D).cuda()
    out_dev = db.jk_direct_batch(D_dev)
    assert out_dev.is_cuda
    torch.cuda.synchronize()
    assert np.abs(out_dev.cpu().numpy() - ref).max() < JK_TOL * scale
    db.close()


def test_batched_jk_full_size_stored_equals_direct():
    """(H2O)8 6-31G** (N = 192): 5 NOCI-shaped sets, stored (two passes over the 10.9 GB tensor)
    against direct (one ERI generation), plus a 2-rank partition of the direct accumulators."""
    import torch
    from pychem_b200 import _lib, engine, structures as S
    db = engine.DeviceBasis(S.Molecule(S.water_cluster(8), "6-31G**"))
    N = db.nbf
    db.schwarz()
    G_dev, _ = db.eri_tensor(1.0e-8, to_host=False)
    D = _general_sets(np.random.default_rng(5), 5, N)
    stored = np.array(db.jk_stored_batch(G_dev, D))
    del G_dev
    db.plan(1.0e-8, 0, 1)
    direct = np.array(db.jk_direct_batch(D))
    scale = np.abs(stored).max()
    assert np.abs(stored - direct).max() < 1e-9 * scale
    total = torch.zeros(D.size, dtype=torch.float64, device="cuda")
    acc = torch.empty_like(total)
    for r in range(2):
        db.plan(1.0e-8, r, 2)
        _lib.check(db.lib.pc_jk_direct_batch_accumulate(db.h, 5, engine._ptr(D), engine._ptr(acc)))
        torch.cuda.synchronize()
        total += acc
    torch.cuda.synchronize()          # `total` is summed on torch's stream, finalised on the library's
    out = np.empty_like(D)
    _lib.check(db.lib.pc_jk_finalize_batch(db.h, 5, engine._ptr(total), engine._ptr(out)))
    assert np.abs(out - direct).max() < 1e-10 * scale
    db.close()


@pytest.mark.skipif(not ref_driver.available(), reason="oracle/_ref (reference copy) not shipped")
@pytest.mark.parametrize("mode", ["stored", "direct"])
def test_lih_sfs_noci_batched_driver(gold, monkeypatch, mode):
    from pychem_b200 import hartree_fock as hf_gpu, integrals as ints_gpu, noci as noci_gpu
    monkeypatch.setenv("PYCHEM_B200_MODE", mode)
    ns = ref_driver.modules()
    undo_hf = hf_gpu.install(ns.hartree_fock)
    undo_noci = noci_gpu.install(ns.noci)
    try:
        mol = ref_driver.run(os.path.join(ref_driver.REF_ROOT, "Tests", "LiH_SFS_NOCI.test.inp"))
        g = gold("lih_631g.npz")
        assert np.abs(np.array([s.TotalEnergy for s in mol.States]) - g["hf"]).max() < E_TOL
        assert np.abs(np.asarray(mol.NOCIEnergies) - g["noci"]).max() < E_TOL
        assert "NOCI output" in mol.OutText
    finally:
        undo_noci()
        undo_hf()
        hf_gpu.release()
        ints_gpu.release()


@pytest.mark.skipif(not ref_driver.available(), reason="oracle/_ref (reference copy) not shipped")
def test_lih_chain2_noci_batched_driver(gold, tmp_path):
    """BASELINE config 4 (LiH chain, SFS-NOCI) through the batched driver."""
    from pychem_b200 import hartree_fock as hf_gpu, integrals as ints_gpu, noci as noci_gpu, structures as S
    ns = ref_driver.modules()
    undo_hf = hf_gpu.install(ns.hartree_fock)
    undo_noci = noci_gpu.install(ns.noci)
    try:
        inp = str(tmp_path / "lih.inp")
        ref_driver.write_input(inp, "lih2", S.lih_chain(2), "6-31G", method="NOCI", reference="UHF",
                               extra='Constrain_Excited = True\nExcitations = "SFS"')
        mol = ref_driver.run(inp)
        g = gold("lih_chain2_noci.npz")
        assert np.abs(np.array([s.TotalEnergy for s in mol.States]) - g["hf"]).max() < E_TOL
        assert np.abs(np.asarray(mol.NOCIEnergies) - g["noci"]).max() < E_TOL
    finally:
        undo_noci()
        undo_hf()
        hf_gpu.release()
        ints_gpu.release()
