#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY (developer tool).  Bisect the hydrogen-fluoride / cc-pVTZ drop-in job quantity by quantity against the reference's
own pieces (tests/golden/hf_ccpvtz_parts.npz, oracle/make_golden_hf_parts.py):
Core, Overlap (one_electron_kernel<3>), J and X_alpha for the reference's converged density
(stored and direct), and the energy expression evaluated with them.

  python oracle/diag_f_shell.py            # GPU (product library)
  python oracle/diag_f_shell.py --emu      # host emulation of the same kernels
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import helpers  # noqa: E402


def report(name, got, ref):
    d = np.abs(np.asarray(got) - ref)
    k = np.unravel_index(int(d.argmax()), d.shape)
    rel = d / np.maximum(np.abs(ref), 1e-300)
    print("%-22s max|d| = %.3e at %s (ref %.6e)   max rel (|ref|>1e-6) = %.3e" %
          (name, d.max(), k, ref[k], rel[np.abs(ref) > 1e-6].max()), flush=True)
    return float(d.max())


def main():
    emu = "--emu" in sys.argv
    g = np.load(os.path.join(ROOT, "tests", "golden", "hf_ccpvtz_parts.npz"))
    mol = helpers.molecule("hf_tz")
    if emu:
        from tests.emu import emu_engine
        db = emu_engine.EmuBasis(mol)
    else:
        from pychem_b200 import engine
        db = engine.DeviceBasis(mol)
    Z = [float(a.NuclearCharge) for a in mol.Atoms]
    R = [[float(x) for x in a.Coordinates] for a in mol.Atoms]
    out = {}
    for rep in range(3):                      # repeated: a race / uninitialised read would not repeat
        core, overlap = db.one_electron(Z, R)
        out["core_%d" % rep] = report("Core (run %d)" % rep, core, g["core"])
        out["overlap_%d" % rep] = report("Overlap (run %d)" % rep, overlap, g["overlap"])
    db.schwarz()
    Dt, Da = np.ascontiguousarray(g["Dt"]), np.ascontiguousarray(g["Da"])
    if emu:
        G = db.eri_tensor(1.0e-8)
        J, Xa, _ = db.jk_stored(G, Dt, Da, Da)
    else:
        G_dev, G = db.eri_tensor(1.0e-8, to_host=True)
        J, Xa, _ = db.jk_stored(G_dev, Dt, Da, Da)
    out["J_stored"] = report("J stored", J, g["J"])
    out["Xa_stored"] = report("Xa stored", Xa, g["Xa"])
    db.plan(1.0e-8, 0, 1)
    J2, Xa2, _ = db.jk_direct(Dt, Da, Da)
    out["J_direct"] = report("J direct", J2, g["J"])
    out["Xa_direct"] = report("Xa direct", Xa2, g["Xa"])
    # E = 1/2 (Dt.Core + Da.Fa + Db.Fb) + Vnn with Fa = Core + J + Xa   (hartree_fock.py:188-201)
    def energy(core_, J_, Xa_):
        F = core_ + J_ + Xa_
        return 0.5 * (np.sum(Dt * core_) + 2.0 * np.sum(Da * F))
    e_ref = energy(g["core"], g["J"], g["Xa"])
    e_mine = energy(np.asarray(core), np.asarray(J), np.asarray(Xa))
    print("electronic energy with the reference's density: mine - reference = %.3e" % (e_mine - e_ref))
    out["dE_fixed_density"] = float(e_mine - e_ref)
    db.close()
    # the drop-in SCF itself, Core/Overlap from the reference's own code or from the device
    from oracle import ref_driver
    if ref_driver.available() and not emu:
        import tempfile
        from pychem_b200 import hartree_fock as hf_gpu, integrals as ints_gpu
        ns = ref_driver.modules()
        for one_e in (False, True):
            undo = hf_gpu.install(ns.hartree_fock, one_electron=one_e)
            try:
                inp = os.path.join(tempfile.mkdtemp(), "hf.inp")
                ref_driver.write_input(inp, "hf", helpers.HYDROGEN_FLUORIDE, "cc-pVTZ")
                m = ref_driver.run(inp)
                e = float(m.States[0].TotalEnergy)
                print("drop-in SCF one_electron=%s: E = %.14f  (E - reference = %.3e)" % (one_e, e, e - float(g["energy"])), flush=True)
                out["scf_dE_one_electron_%s" % one_e] = e - float(g["energy"])
                report("  SCF Core", np.asarray(m.Core), g["core"])
                report("  SCF density", np.asarray(m.States[0].Alpha.Density), g["Da"])
            finally:
                undo()
                hf_gpu.release()
                ints_gpu.release()
    print("DIAG " + json.dumps(out))


if __name__ == "__main__":
    main()
