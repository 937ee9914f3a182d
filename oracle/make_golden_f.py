#!/usr/bin/env python
"""Mint tests/golden/f_shell_ccpvtz.npz: f-shell goldens from the REAL reference (oracle/_ref).

TEST INFRASTRUCTURE ONLY.  Everything stored comes out of the reference's own C extension driven
by its own Python (integrals.two_electron, hartree_fock.make_core_matrices, pychem.main); nothing
from our oracle or CUDA path goes in.

  quartets/blocks/offsets  sampled shell quartets of a four-heavy-atom cluster (C N O N at
                           general positions, cc-pVTZ: f shells on four distinct centres), a few
                           for every (pair l, pair l) class that contains an f shell plus some
                           s/p/d ones -- integrals.two_electron, shell quartet by shell quartet
  affected                 per sampled quartet: 1 if one of its pairs is a (d f) pair in the
                           reference's "goofy" order (d shell first).  For those the reference's
                           HRR is wrong (Methods/c_ints/two_electron_hrr.c:18 uses the stride
                           angmom_index(0,0,lb+1) where ncart(lb+1) = angmom_index(0,0,lb+1)+1 is
                           meant; harmless while the lower shell is s or p).  Their `blocks`
                           are what the reference returns and are NOT used as parity targets.
  md_quartets/md_blocks/md_offsets   independent McMurchie-Davidson values (oracle/md_eri.py)
                           for affected quartets and, as a cross-check of the yardstick itself,
                           for some unaffected ones (md_vs_reference_unaffected = their largest
                           deviation from the reference)
  core, overlap            one-electron matrices of the same cluster (make_core_matrices)
  hf_*                     RHF on hydrogen fluoride / cc-pVTZ through pychem.main: the reference's
                           energy and tensor checksums, kept for the record only -- F carries d
                           and f shells, so the reference's own numbers contain the defect
  fixed_blocks, fixed_hf_* the same sampled quartets and the same RHF job from the reference with
                           that ONE expression corrected: its C sources are compiled once more
                           from a scratch copy in which two_electron_hrr.c:18 reads
                           `nlb0 = angmom_index(0,0,lb+1)+1;` (nothing else differs; the copy
                           lives in a temporary directory and is deleted).  With the stride
                           corrected the reference agrees with the McMurchie-Davidson values on
                           the affected quartets (fixed_vs_md) and is unchanged elsewhere
                           (fixed_vs_reference_unaffected = 0), which is the evidence for the
                           diagnosis; fixed_hf_energy is the parity target of the f-shell
                           drop-in test.

Cartesian->spherical convention: Data/transform_basis.py is evaluated with true division (the
Python-3 copy of the driver, oracle/build_ref.py), i.e. sqrt(5/2) is sqrt(2.5).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import md_eri, ref_driver  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# C N O N, no symmetry, no axis alignment (Angstrom)
CLUSTER = [["C", 6.0, 0.00, 0.00, 0.00],
           ["N", 7.0, 1.05, 0.35, -0.20],
           ["O", 8.0, -0.40, 1.15, 0.55],
           ["N", 7.0, 0.60, -0.85, 1.10]]
HYDROGEN_FLUORIDE = [["F", 9.0, 0.0, 0.0, 0.0], ["H", 1.0, 0.31, 0.42, 0.74]]     # r = 0.906 A


def shell_l(mol):
    return [cg.AngularMomentum for atom in mol.Atoms for cg in atom.Basis]


def sample(ns, mol, per_f_class, per_other_class, seed, max_tries=400000):
    rng = np.random.default_rng(seed)
    ls = shell_l(mol)
    n = len(ls)
    got, quartets = {}, []
    for _ in range(max_tries):
        a, b, c, d = (int(x) for x in rng.integers(0, n, 4))
        if a > b:
            a, b = b, a
        if c > d:
            c, d = d, c
        key = tuple(sorted([tuple(sorted((ls[a], ls[b]))), tuple(sorted((ls[c], ls[d])))]))
        want = per_f_class if max(ls[a], ls[b], ls[c], ls[d]) == 3 else per_other_class
        if got.get(key, 0) >= want:
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


def compile_fixed_c_ints():
    """The reference's extension with the goofy-HRR stride corrected, built in a scratch directory
    from a copy of its sources (authoring container only; /root/reference is read-only)."""
    import importlib.machinery
    import importlib.util
    import shutil
    import subprocess
    import sysconfig
    import tempfile
    from oracle import build_ref
    ref = "/root/reference"
    td = tempfile.mkdtemp(prefix="pychem_fixed_")
    meth = os.path.join(td, "Methods")
    shutil.copytree(os.path.join(ref, "Methods"), meth)
    hrr = os.path.join(meth, "c_ints", "two_electron_hrr.c")
    src = open(hrr).read()
    old = "nlb0 = angmom_index(0,0,lb+1); nlb1 = nlb;"
    assert src.count(old) == 1
    with open(hrr, "w") as fh:
        fh.write(src.replace(old, "nlb0 = angmom_index(0,0,lb+1)+1; nlb1 = nlb;"))
    table_c = os.path.join(td, "interpolation_table.c")
    build_ref.write_table_c(table_c)
    target = os.path.join(td, "_c_ints" + sysconfig.get_config_var("EXT_SUFFIX"))
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-w", "-I" + sysconfig.get_paths()["include"], "-I" + np.get_include(),
           "-I" + os.path.join(meth, "c_ints"), "-I" + meth]
    cmd += [os.path.join(meth, x) for x in build_ref.C_SOURCES] + [table_c, "-lm", "-o", target]
    subprocess.check_call(cmd)
    loader = importlib.machinery.ExtensionFileLoader("_c_ints", target)
    spec = importlib.util.spec_from_loader("_c_ints", loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    shutil.rmtree(td, ignore_errors=True)
    return mod


def main():
    ns = ref_driver.modules()
    out = {}
    t = time.time()
    mol, _ = ref_driver.build_molecule(CLUSTER, "cc-pVTZ")
    q, blocks, offs, got = sample(ns, mol, 3, 1, 31)
    nf = sum(1 for k in got if max(max(k[0]), max(k[1])) == 3)
    print("cluster: %d quartets, %d classes (%d with an f shell), %.1fs" % (len(q), len(got), nf, time.time() - t))
    ls = shell_l(mol)
    goofy_df = lambda a, b: ls[a] == 2 and ls[b] == 3          # noqa: E731  (a <= b: shell order)
    affected = np.array([int(goofy_df(a, b) or goofy_df(c, d)) for a, b, c, d in q], dtype=np.int32)
    out.update(quartets=q, blocks=blocks, offsets=offs, affected=affected)
    print("affected by the reference's goofy-HRR stride:", int(affected.sum()))
    # independent McMurchie-Davidson values: affected quartets whose four shells are cheap enough
    # for the plain-numpy evaluation, plus unaffected ones as a check of the yardstick
    from Data import transform_basis
    shells = []
    for atom in mol.Atoms:
        for cg in atom.Basis:
            shells.append((np.array(atom.Coordinates, dtype=float), int(cg.AngularMomentum),
                           [float(x) for x in cg.Exponents], [float(x) for x in cg.ScaledCCs],
                           [float(x) for x in cg.ContractionScaling]))
    cost = lambda qq: np.prod([len(shells[s][2]) for s in qq]) * 3.0 ** sum(shells[s][1] for s in qq)   # noqa: E731
    md_q, md_b, worst_unaff = [], [], 0.0
    t = time.time()
    n_aff = n_un = 0
    for k, qq in enumerate(q):
        if cost(qq) > 3.0 ** 10 * 4:
            continue
        if affected[k]:
            n_aff += 1
        elif n_un < 25:
            n_un += 1
        else:
            continue
        with np.errstate(all="ignore"):        # the reference switches numpy to raise on underflow
            blk = md_eri.shell_quartet([shells[s] for s in qq], transform_basis.cart_to_spher)
        md_q.append(qq)
        md_b.append(blk.ravel())
        if not affected[k]:
            worst_unaff = max(worst_unaff, float(np.abs(blk.ravel() - blocks[offs[k]:offs[k + 1]]).max()))
    print("McMurchie-Davidson: %d affected + %d unaffected quartets, %.1fs; unaffected vs reference %.2e"
          % (n_aff, n_un, time.time() - t, worst_unaff))
    out.update(md_quartets=np.array(md_q, dtype=np.int32), md_blocks=np.concatenate(md_b),
               md_offsets=np.concatenate([[0], np.cumsum([len(x) for x in md_b])]).astype(np.int64),
               md_vs_reference_unaffected=worst_unaff)
    with np.errstate(all="ignore"):
        ns.hartree_fock.make_core_matrices(mol)
    out.update(core=np.array(mol.Core), overlap=np.array(mol.Overlap))

    inp = os.path.join(GOLD, "_hf.inp")
    ref_driver.write_input(inp, "hf", HYDROGEN_FLUORIDE, "cc-pVTZ")
    t = time.time()
    mol = ref_driver.run(inp)
    os.remove(inp)
    G = np.asarray(mol.CoulombIntegrals)
    st = mol.States[0]
    print("HF cc-pVTZ", repr(st.TotalEnergy), G.sum(), (G ** 2).sum(), "%.1fs" % (time.time() - t))
    out.update(hf_energy=st.TotalEnergy, hf_G_sum=G.sum(), hf_G_sq=(G ** 2).sum(),
               hf_G_sample=G.ravel()[::997].copy(), hf_nbf=mol.NOrbitals,
               hf_Da=np.array(st.Alpha.Density))

    # ---- the reference with the HRR stride corrected ----
    t = time.time()
    fixed = compile_fixed_c_ints()
    stock = ns.integrals._c_ints
    ns.integrals._c_ints = fixed
    try:
        mol, _ = ref_driver.build_molecule(CLUSTER, "cc-pVTZ")
        fb = []
        for (a, b, c, d) in q:
            blk = ns.integrals.two_electron(mol.ShellPairs[(int(a), int(b))], mol.ShellPairs[(int(c), int(d))], 0, -1.0)
            fb.append(np.asarray(blk).ravel().copy())
        fb = np.concatenate(fb)
        un = np.concatenate([np.arange(offs[k], offs[k + 1]) for k in range(len(q)) if not affected[k]])
        d_un = float(np.abs(fb[un] - blocks[un]).max())
        d_md = 0.0
        pos = {tuple(int(x) for x in qq): k for k, qq in enumerate(q)}
        for qq, lo, hi in zip(out["md_quartets"], out["md_offsets"][:-1], out["md_offsets"][1:]):
            k = pos[tuple(int(x) for x in qq)]
            d_md = max(d_md, float(np.abs(fb[offs[k]:offs[k + 1]] - out["md_blocks"][lo:hi]).max()))
        print("fixed reference: unaffected vs stock %.2e, vs McMurchie-Davidson (all md quartets) %.2e" % (d_un, d_md))
        inp = os.path.join(GOLD, "_hf.inp")
        ref_driver.write_input(inp, "hf", HYDROGEN_FLUORIDE, "cc-pVTZ")
        mol = ref_driver.run(inp)
        os.remove(inp)
        G = np.asarray(mol.CoulombIntegrals)
        st = mol.States[0]
        print("HF cc-pVTZ, fixed reference", repr(st.TotalEnergy), G.sum(), (G ** 2).sum(), "%.1fs" % (time.time() - t))
        out.update(fixed_blocks=fb, fixed_vs_reference_unaffected=d_un, fixed_vs_md=d_md,
                   fixed_hf_energy=st.TotalEnergy, fixed_hf_G_sum=G.sum(), fixed_hf_G_sq=(G ** 2).sum(),
                   fixed_hf_G_sample=G.ravel()[::997].copy(), fixed_hf_Da=np.array(st.Alpha.Density))
    finally:
        ns.integrals._c_ints = stock
    np.savez_compressed(os.path.join(GOLD, "f_shell_ccpvtz.npz"), **out)
    print("wrote f_shell_ccpvtz.npz")


if __name__ == "__main__":
    main()
