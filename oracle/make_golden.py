#!/usr/bin/env python
"""Mint tests/golden/*.npz by running the REAL reference (oracle/_ref, see build_ref.py).

TEST INFRASTRUCTURE ONLY.  Run in the authoring container (needs /root/reference to have been
built into oracle/_ref).  Every array stored here comes out of the reference's own C extension
driven by the reference's own Python (integrals.two_electron, hartree_fock.evaluate_2e_ints,
hartree_fock.make_coulomb_exchange_matrices, pychem.main) -- nothing from our oracle or CUDA
path goes in.

Fixtures:
  h2_6311g.npz     Tests/H2_HF.test.inp     : full tensor, Schwarz bounds, SCF energy
  lih_631g.npz     Tests/LiH_SFS_NOCI.test.inp: full tensor, HF/NOCI energies, J/K for the
                                              SCF densities and for a non-symmetric co-density
  h2o_631gss.npz   H2O 6-31G**              : full tensor (unique blocks), J/K, SCF energy
  h2o2_631gss.npz  (H2O)2 6-31G**           : sampled shell quartets covering all 21 l<=2
                                              classes on four distinct centres
  benzene_631gs.npz benzene 6-31G*          : sampled shell quartets
  h2o_631gss_cartd.npz H2O 6-31G**, Cartesian_L = [2]: full tensor, SCF energy
  h3_sto3g_mp2.npz Tests/example1.inp       : HF and HF+MP2 total energies
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_driver  # noqa: E402
from pychem_b200 import structures as S  # noqa: E402  (geometry helpers only)

GOLD = os.path.join(ROOT, "tests", "golden")
REF_TESTS = os.path.join(ref_driver.REF_ROOT, "Tests")


class State:
    """Minimal stand-in for ElectronicState / CoDensityState (noci.py:18-26)."""
    class M:
        pass

    def __init__(self, Dt, Da, Db):
        self.Total, self.Alpha, self.Beta = State.M(), State.M(), State.M()
        self.Total.Density, self.Alpha.Density, self.Beta.Density = Dt, Da, Db


def ref_jk(ns, mol, Dt, Da, Db):
    st = State(Dt, Da, Db)
    ns.hartree_fock.make_coulomb_exchange_matrices(mol, st)
    return st.Total.Coulomb, st.Alpha.Exchange, st.Beta.Exchange


def random_densities(N, seed, symmetric):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(2):
        X = rng.uniform(-1, 1, (N, N))
        out.append(0.5 * (X + X.T) if symmetric else X)
    Da, Db = out
    return Da + Db, Da, Db


def bounds_array(mol):
    n = mol.NCgtf
    npair = n * (n + 1) // 2
    out = np.zeros((npair, 49))
    p = 0
    for a in range(n):
        for b in range(a, n):
            blk = np.asarray(mol.Bounds[a][b])
            out[p, :blk.size] = blk.ravel()
            p += 1
    return out


def shell_class(mol):
    ls = []
    for atom in mol.Atoms:
        for cg in atom.Basis:
            ls.append(cg.AngularMomentum)
    return ls


def sample_quartets(ns, mol, per_class, seed, max_tries=200000):
    """Random shell quartets, `per_class` for every (sorted pair l, sorted pair l) class."""
    rng = np.random.default_rng(seed)
    ls = shell_class(mol)
    n = len(ls)
    got = {}
    quartets = []
    for _ in range(max_tries):
        a, b, c, d = (int(x) for x in rng.integers(0, n, 4))
        if a > b:
            a, b = b, a
        if c > d:
            c, d = d, c
        key = tuple(sorted([tuple(sorted((ls[a], ls[b]))), tuple(sorted((ls[c], ls[d])))]))
        if got.get(key, 0) >= per_class:
            continue
        got[key] = got.get(key, 0) + 1
        quartets.append((a, b, c, d))
    quartets.sort()
    blocks = []
    for (a, b, c, d) in quartets:
        blk = ns.integrals.two_electron(mol.ShellPairs[(a, b)], mol.ShellPairs[(c, d)], 0, -1.0)
        blocks.append(np.asarray(blk).ravel().copy())
    offs = np.concatenate([[0], np.cumsum([len(x) for x in blocks])])
    return np.array(quartets, dtype=np.int32), np.concatenate(blocks), offs.astype(np.int64), got


def mint_cartesian_d(ns, path=None):
    """h2o_631gss_cartd.npz: the reference's Cartesian_L keyword (Util/structures.py:844-849), RHF
    on H2O 6-31G** with six Cartesian d functions: dense tensor and SCF energy."""
    inp = os.path.join(GOLD, "_h2o_cartd.inp")
    ref_driver.write_input(inp, "h2o", S.H2O_MONOMER, "6-31G**", extra="Cartesian_L = [2]")
    mol = ref_driver.run(inp)
    os.remove(inp)
    np.savez_compressed(path or os.path.join(GOLD, "h2o_631gss_cartd.npz"),
                        G=np.asarray(mol.CoulombIntegrals, dtype=np.float64), energy=mol.States[0].TotalEnergy)
    print("H2O Cartesian d", mol.NOrbitals, repr(mol.States[0].TotalEnergy))


def main():
    os.makedirs(GOLD, exist_ok=True)
    ns = ref_driver.modules()

    # ---------------- H2 6-311G (Tests/H2_HF.test.inp, run end to end) ----------------
    mol = ref_driver.run(os.path.join(REF_TESTS, "H2_HF.test.inp"))
    G = mol.CoulombIntegrals
    Dt, Da, Db = random_densities(mol.NOrbitals, 11, True)
    J, Xa, Xb = ref_jk(ns, mol, Dt, Da, Db)
    np.savez_compressed(os.path.join(GOLD, "h2_6311g.npz"), G=G, bounds=bounds_array(mol),
                        energy=mol.States[0].TotalEnergy, Dt=Dt, Da=Da, Db=Db, J=J, Xa=Xa, Xb=Xb)
    print("H2", repr(mol.States[0].TotalEnergy), G.sum())

    # ---------------- LiH 6-31G SFS NOCI (Tests/LiH_SFS_NOCI.test.inp) ----------------
    mol = ref_driver.run(os.path.join(REF_TESTS, "LiH_SFS_NOCI.test.inp"))
    G = mol.CoulombIntegrals
    hf = np.array([s.TotalEnergy for s in mol.States])
    noci = np.array(getattr(mol, "NOCIEnergies", [])) if hasattr(mol, "NOCIEnergies") else np.zeros(0)
    Dt, Da, Db = random_densities(mol.NOrbitals, 12, False)       # non-symmetric (NOCI co-density shape)
    J, Xa, Xb = ref_jk(ns, mol, Dt, Da, Db)
    np.savez_compressed(os.path.join(GOLD, "lih_631g.npz"), G=G, bounds=bounds_array(mol), hf=hf, noci=noci,
                        Dt=Dt, Da=Da, Db=Db, J=J, Xa=Xa, Xb=Xb)
    print("LiH", hf, noci, G.sum(), np.count_nonzero(G))

    # ---------------- H2O 6-31G** -------------------------------------------------------
    inp = os.path.join(GOLD, "_h2o.inp")
    ref_driver.write_input(inp, "h2o", S.H2O_MONOMER, "6-31G**")
    t = time.time()
    mol = ref_driver.run(inp)
    os.remove(inp)
    G = mol.CoulombIntegrals
    Dt, Da, Db = random_densities(mol.NOrbitals, 13, True)
    J, Xa, Xb = ref_jk(ns, mol, Dt, Da, Db)
    Dt2, Da2, Db2 = random_densities(mol.NOrbitals, 14, False)
    J2, Xa2, Xb2 = ref_jk(ns, mol, Dt2, Da2, Db2)
    st = mol.States[0]
    np.savez_compressed(os.path.join(GOLD, "h2o_631gss.npz"), G=G.astype(np.float64),
                        bounds=bounds_array(mol), energy=st.TotalEnergy,
                        Dt=Dt, Da=Da, Db=Db, J=J, Xa=Xa, Xb=Xb,
                        Dt2=Dt2, Da2=Da2, Db2=Db2, J2=J2, Xa2=Xa2, Xb2=Xb2,
                        scf_Dt=st.Total.Density, scf_Da=st.Alpha.Density, scf_Db=st.Beta.Density)
    print("H2O", repr(st.TotalEnergy), G.sum(), (G ** 2).sum(), "%.1fs" % (time.time() - t))

    # ---------------- (H2O)2 6-31G**: sampled quartets, all classes ---------------------
    mol, _ = ref_driver.build_molecule(S.water_cluster(2), "6-31G**")
    q, blocks, offs, got = sample_quartets(ns, mol, 12, 21)
    np.savez_compressed(os.path.join(GOLD, "h2o2_631gss.npz"), quartets=q, blocks=blocks, offsets=offs)
    print("(H2O)2 sampled", len(q), "quartets in", len(got), "classes")

    # ---------------- benzene 6-31G*: sampled quartets ----------------------------------
    mol, _ = ref_driver.build_molecule(S.benzene(), "6-31G*")
    q, blocks, offs, got = sample_quartets(ns, mol, 8, 22)
    np.savez_compressed(os.path.join(GOLD, "benzene_631gs.npz"), quartets=q, blocks=blocks, offsets=offs)
    print("benzene sampled", len(q), "quartets in", len(got), "classes")

    # ---------------- H3 STO-3G CUHF + MP2 (Tests/example1.inp) --------------------------
    mol = ref_driver.run(os.path.join(REF_TESTS, "example1.inp"), quiet=True)
    ehf = mol.States[0].TotalEnergy
    # mp2.do writes HF+MP2 total energy to the .out file (mp2.py:113)
    emp2 = [float(l.split()[-1]) for l in mol.OutText.splitlines() if "Total MP2 energy" in l][0]
    np.savez_compressed(os.path.join(GOLD, "h3_sto3g_mp2.npz"), hf=ehf, mp2_total=emp2, G=mol.CoulombIntegrals,
                        Ca=mol.States[0].Alpha.MOs, Cb=mol.States[0].Beta.MOs,
                        Ea=mol.States[0].Alpha.Energies, Eb=mol.States[0].Beta.Energies,
                        na=mol.NAlphaElectrons, nb=mol.NBetaElectrons)
    print("H3", repr(ehf), repr(emp2))

    # ---------------- LiH chains, SFS-NOCI (Tests/LiH_SFS_NOCI.test.inp scaled up) --------
    for k in (2, 3):
        inp = os.path.join(GOLD, "_lih%d.inp" % k)
        ref_driver.write_input(inp, "lih%d" % k, S.lih_chain(k), "6-31G", method="NOCI", reference="UHF",
                               extra='Constrain_Excited = True\nExcitations = "SFS"')
        t = time.time()
        mol = ref_driver.run(inp)
        os.remove(inp)
        np.savez_compressed(os.path.join(GOLD, "lih_chain%d_noci.npz" % k),
                            hf=np.array([s.TotalEnergy for s in mol.States]),
                            noci=np.asarray(mol.NOCIEnergies), nbf=mol.NOrbitals)
        print("LiH chain", k, mol.NOrbitals, [s.TotalEnergy for s in mol.States], mol.NOCIEnergies,
              "%.1fs" % (time.time() - t))

    # ---------------- one-electron matrices (hartree_fock.make_core_matrices) ------------
    one_e = {}
    for name, coords, basis, extra in (("h2", [["H", 1.0, 0.0, 0.0, 0.0], ["H", 1.0, 0.0, 0.0, 1.0]], "6-311G", ""),
                                       ("lih", [["Li", 3.0, 0.0, 0.0, 0.0], ["H", 1.0, 2.2, 0.0, 0.0]], "6-31G", ""),
                                       ("h2o", S.H2O_MONOMER, "6-31G**", ""),
                                       ("h2o_cartd", S.H2O_MONOMER, "6-31G**", "Cartesian_L = [2]"),
                                       ("h2o2", S.water_cluster(2), "6-31G**", ""),
                                       ("benzene", S.benzene(), "6-31G*", "")):
        mol, _ = ref_driver.build_molecule(coords, basis, extra=extra)
        with np.errstate(all="ignore"):
            ns.hartree_fock.make_core_matrices(mol)
        one_e[name + "_core"] = np.array(mol.Core)
        one_e[name + "_overlap"] = np.array(mol.Overlap)
        print("one-electron", name, mol.NOrbitals, np.trace(mol.Core), np.trace(mol.Overlap))
    np.savez_compressed(os.path.join(GOLD, "one_electron.npz"), **one_e)

    # ---------------- H2O 6-31G** with Cartesian d functions (Cartesian_L = [2]) ---------
    mint_cartesian_d(ns)

    # ---------------- H2O 6-31G** RHF + MP2 (reference mp2.do, O(N^6) Python) -----------
    inp = os.path.join(GOLD, "_h2o_mp2.inp")
    ref_driver.write_input(inp, "h2omp2", S.H2O_MONOMER, "6-31G**", method="MP2")
    t = time.time()
    mol = ref_driver.run(inp)
    os.remove(inp)
    st = mol.States[0]
    emp2 = [float(l.split()[-1]) for l in mol.OutText.splitlines() if "Total MP2 energy" in l][0]
    np.savez_compressed(os.path.join(GOLD, "h2o_631gss_mp2.npz"), hf=st.TotalEnergy, mp2_total=emp2,
                        Ca=st.Alpha.MOs, Cb=st.Beta.MOs, Ea=st.Alpha.Energies, Eb=st.Beta.Energies,
                        na=mol.NAlphaElectrons, nb=mol.NBetaElectrons)
    print("H2O MP2", repr(st.TotalEnergy), repr(emp2), "%.1fs" % (time.time() - t))


if __name__ == "__main__":
    main()
