"""Mint tests/golden/h2o_631gss_scattering.npz from the REAL reference (oracle/_ref).

TEST INFRASTRUCTURE ONLY.  Runs the reference's property job (Job_Type = Property,
Property_Type = Scattering; pychem.py:132-135 -> Methods/properties.py:6-32) for H2O 6-31G** and
stores, per grid value, the scattering-integral tensor (packed: one value per 8-fold-unique
(ab|cd)), the Schwarz factors and the printed scattering intensity of the RHF state.

    python oracle/build_ref.py && python oracle/make_golden_scattering.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_driver                         # noqa: E402
from oracle.make_golden import bounds_array           # noqa: E402
from pychem_b200 import structures as S               # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
GRID = [0.0, 0.5, 2.0, 7.5]
# At S = 0 the integrals are products of overlaps; diagonals that are zero by symmetry come out
# as -1e-36 .. -6e-17 and the reference, which runs under numpy.seterr(all="raise")
# (Methods/diis.py), dies in numpy.sqrt (hartree_fock.py:254).  The end-to-end property job is
# therefore run on GRID[1:], and S = 0 is evaluated by hand with the error state relaxed (the
# NaN bounds then only switch off blocks that vanish anyway); NaN bounds are stored as 0.
#
# The reference also never clears molecule.CoulombIntegrals between grid points
# (hartree_fock.py:298 "Already initialized to zero" holds for the first call only), so blocks
# screened out at one grid value keep the numbers of the previous evaluation -- the Coulomb
# integrals of the SCF for the first grid point.  Printed intensities are history dependent
# (H2O 6-31G**, S = 7.5: 10.5639 printed, 10.6947 from a cleared tensor; S = 0 must give
# N_el^2 = 100 and does only when cleared).  The golden tensors are minted from a CLEARED tensor,
# which is what one call of evaluate_2e_ints(molecule, 1, S) on a fresh molecule returns; the
# printed (stateful) values are kept as `printed` for the record.


def packed_index(n):
    """(a,b,c,d) index arrays of the canonical quadruples a>=b, c>=d, ab>=cd."""
    a, b = np.tril_indices(n)
    p, q = np.tril_indices(len(a))
    return a[p], b[p], a[q], b[q]


def main():
    ns = ref_driver.modules()
    inp = os.path.join(GOLD, "_h2o_scat.inp")
    ref_driver.write_input(inp, "h2oscat", S.H2O_MONOMER, "6-31G**", job_type="Property",
                           extra='Property_Type = "Scattering"\nProperty_Grid = %r' % GRID[1:])
    mol = ref_driver.run(inp)
    os.remove(inp)
    lines = mol.OutText.splitlines()
    start = [i for i, l in enumerate(lines) if "Grid value -> Scattering" in l][0]
    printed = np.array([[float(x) for x in lines[start + 2 + k].split()] for k in range(len(GRID) - 1)])
    assert np.allclose(printed[:, 0], GRID[1:])
    st = mol.States[0]
    ia, ib, ic, id_ = packed_index(mol.NOrbitals)
    intensity = np.zeros(len(GRID))
    out = dict(grid=np.array(GRID), intensity=intensity, printed=printed, energy=st.TotalEnergy,
               scf_Dt=st.Total.Density, scf_Da=st.Alpha.Density, scf_Db=st.Beta.Density)
    with np.errstate(all="ignore"):
        for k, g in enumerate(GRID):
            mol.CoulombIntegrals[...] = 0.0
            ns.hartree_fock.evaluate_2e_ints(mol, 1, g)
            G = np.asarray(mol.CoulombIntegrals)
            assert np.array_equal(G, G.transpose(1, 0, 2, 3))
            assert np.abs(G - G.transpose(2, 3, 0, 1)).max() < 1e-14    # diagonal blocks: roundoff
            out["G%d" % k] = G[ia, ib, ic, id_].copy()
            out["bounds%d" % k] = np.nan_to_num(bounds_array(mol), nan=0.0)
            two = ns.properties.make_two_particle_density_matrices(mol.NOrbitals, st)
            val = mol.NElectrons + ns.properties.contract_two(mol.NOrbitals, two, mol.CoulombIntegrals)
            if k in (1, 2):        # nothing screened out at these grid values: history free
                assert abs(val - printed[k - 1, 1]) < 5e-11, (val, printed[k - 1])
            intensity[k] = val
            print("grid", g, "intensity", repr(val), "nnz", np.count_nonzero(G), "sum", G.sum())
    np.savez_compressed(os.path.join(GOLD, "h2o_631gss_scattering.npz"), **out)


if __name__ == "__main__":
    main()
