"""GPU-backed mirror of the reference's property job, Methods/properties.py.

    calculate(settings, molecule)                                   (properties.py:6-34)

Electron-scattering intensities: for every grid value S the scattering-kernel integrals
(ints_type = 1, Methods/c_ints/two_electron_scattering.c) are generated on the device and
contracted with the mean-field two-particle density of each state (properties.py:38-70):

    I(S) = N_el + sum_abcd [ Dt_ab Dt_cd - Da_ad Da_cb - Db_ad Db_cb ] (ab|cd)_S
         = N_el + <Dt, J[Dt]> + <Da, X[Da]> + <Db, X[Db]>

with J / X = -K the matrices of hartree_fock.make_coulomb_exchange_matrices -- so the O(N^4)
Python loops of the reference become one streaming pass of the J/K kernel over the tensor in HBM.

One deliberate difference, stated in DESIGN.md: the reference never clears
molecule.CoulombIntegrals between grid points (hartree_fock.py:298), so blocks that are screened
out at one S keep the numbers of the previous evaluation and its printed intensities depend on
the history.  Here every grid value starts from a cleared tensor (S = 0 then gives N_el^2).
"""
import numpy as np

from . import hartree_fock


def scattering_intensity(molecule, state, grid_value, evaluate=True):
    """I(S) of one electronic state; ``evaluate=False`` reuses the tensor already on the device."""
    if evaluate:
        hartree_fock.evaluate_2e_ints(molecule, 1, grid_value)
    st = hartree_fock._STATE[id(molecule)]
    Dt = np.ascontiguousarray(state.Total.Density, dtype=np.float64)
    Da = np.ascontiguousarray(state.Alpha.Density, dtype=np.float64)
    Db = np.ascontiguousarray(state.Beta.Density, dtype=np.float64)
    J, Xa, Xb = st["db"].jk_stored(st["G_dev"], Dt, Da, Db)
    return float(molecule.NElectrons + (Dt * J).sum() + (Da * Xa).sum() + (Db * Xb).sum())


def calculate(settings, molecule):
    if getattr(settings, "PropertyType", None) != "SCATTERING":
        return None
    patterns = [[] for _ in molecule.States]
    for grid_point in settings.PropertyGrid:
        hartree_fock.evaluate_2e_ints(molecule, 1, grid_point)
        for index, state in enumerate(molecule.States):
            patterns[index].append(scattering_intensity(molecule, state, grid_point, evaluate=False))
    out = getattr(settings, "OutFile", None)
    if out is not None:          # same layout as properties.py:26-32
        text = "Grid value -> Scattering patterns for each electronic state \n\n"
        for i, grid_point in enumerate(settings.PropertyGrid):
            text += "%10.6f" % grid_point
            for index in range(len(molecule.States)):
                text += "%16.12f" % patterns[index][i]
            text += "\n"
        out.write(text + "\n")
    return patterns


def install(reference_properties):
    """Rebind ``calculate`` and the by-name import of evaluate_2e_ints (properties.py:3) inside
    the reference's Methods.properties module.  Returns the undo callable."""
    saved = (reference_properties.calculate, reference_properties.evaluate_2e_ints)
    printf = reference_properties.printf

    def calculate_with_banners(settings, molecule):       # banners of properties.py:10,34
        if settings.PropertyType == "SCATTERING":
            printf.delimited_text(settings.OutFile, " Property calculation - electron scattering intensities ")
        patterns = calculate(settings, molecule)
        printf.delimited_text(settings.OutFile, " End of property calculation ")
        return patterns

    reference_properties.calculate = calculate_with_banners
    reference_properties.evaluate_2e_ints = hartree_fock.evaluate_2e_ints

    def uninstall():
        reference_properties.calculate, reference_properties.evaluate_2e_ints = saved
    return uninstall
