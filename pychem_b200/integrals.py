"""GPU-backed mirror of the reference's two-electron entry point in Methods/integrals.py.

    two_electron(shell_pair1, shell_pair2, ints_type, grid_value) -> ndarray(nlA, nlB, nlC, nlD)

Same name, arguments and return convention as Methods/integrals.py:427-555 (including the
block being returned in (pair1 | pair2) order whichever pair carries more angular momentum).
The whole fundamentals -> VRR -> contraction -> HRR -> normalise -> spherical chain runs in one
CUDA kernel launch behind the C ABI (pc_eri_quartets); no CPU fallback exists.

``ints_type == 1`` selects the scattering fundamentals (Methods/c_ints/two_electron_scattering.c,
spherical_bessel_j.c) at ``grid_value``; the recursion and the kernels are the same.
"""
import numpy as np

from .engine import DeviceBasis

_BASIS_CACHE = {}


def _stamp(molecule):
    """What the device tables of a molecule depend on: basis label, shell / function counts, the
    Cartesian-shell choice and every atom's label and coordinates -- so a molecule that is mutated
    in place (Atom.update_coords, Util/structures.py:827; a basis swap, :49-55) gets fresh
    shell-pair tables, Schwarz bounds, one-electron matrices and ERI tensor.  (Edited exponents
    under an unchanged basis label are not detected; call release(molecule) after such edits.)"""
    atoms = tuple((getattr(a, "Label", None), float(getattr(a, "NuclearCharge", 0.0)),
                   tuple(float(x) for x in a.Coordinates), len(a.Basis)) for a in molecule.Atoms)
    cart = getattr(molecule, "CartesianL", None)
    return (getattr(molecule, "Basis", None), int(molecule.NCgtf), int(molecule.NOrbitals),
            tuple(cart) if cart is not None else None, atoms)


def device_basis(molecule):
    """One DeviceBasis per molecule object, rebuilt when its stamp changes.  The cache holds the
    molecule weakly: dropping the molecule frees its device state at the next lookup."""
    import weakref
    key = id(molecule)
    stamp = _stamp(molecule)
    ent = _BASIS_CACHE.get(key)
    if ent is None or ent[0] != stamp or ent[2]() is not molecule:
        if ent is not None:
            ent[1].close()
        for k in [k for k, e in _BASIS_CACHE.items() if e[2]() is None]:     # owners that are gone
            _BASIS_CACHE.pop(k)[1].close()
        try:
            ref = weakref.ref(molecule)
        except TypeError:                       # an object without weak-reference support
            ref = (lambda m: (lambda: m))(molecule)
        ent = (stamp, DeviceBasis(molecule), ref)
        _BASIS_CACHE[key] = ent
    return ent[1]


def release(molecule=None):
    """Free cached device state (all molecules when called without argument)."""
    keys = list(_BASIS_CACHE) if molecule is None else [id(molecule)]
    for k in keys:
        ent = _BASIS_CACHE.pop(k, None)
        if ent is not None:
            ent[1].close()


def release_key(key):
    ent = _BASIS_CACHE.pop(key, None)
    if ent is not None:
        ent[1].close()


def shell_index_of(molecule, shell):
    """Global shell index of a ``Shell`` (the reference keeps only the first basis-function
    index list ``Ivec`` on it, Util/structures.py:962-968)."""
    first = shell.Ivec[0]
    table = device_basis(molecule).table
    idx = int(np.searchsorted(table.first_fn, first))
    if idx >= table.nshell or table.first_fn[idx] != first:
        raise ValueError("shell does not belong to this molecule")
    return idx


def two_electron(shell_pair1, shell_pair2, ints_type=0, grid_value=-1.0, molecule=None):
    """(ab|cd) block for two ShellPair objects.

    ``molecule`` must be given (or have been bound with ``bind(molecule)``) so the shells can
    be located in the device tables; the reference's signature has no such argument because
    its ShellPair objects carry the primitive tables themselves.
    """
    if ints_type not in (0, 1):
        raise ValueError("two_electron: ints_type must be 0 (repulsion) or 1 (scattering)")
    molecule = molecule or _BOUND.get("molecule")
    if molecule is None:
        raise ValueError("two_electron: bind a molecule first (pychem_b200.integrals.bind)")
    db = device_basis(molecule)
    db.set_ints_type(ints_type, grid_value)
    a = shell_index_of(molecule, shell_pair1.Centre1)
    b = shell_index_of(molecule, shell_pair1.Centre2)
    c = shell_index_of(molecule, shell_pair2.Centre1)
    d = shell_index_of(molecule, shell_pair2.Centre2)
    return db.eri_quartets([(a, b, c, d)])[0]


_BOUND = {}


def bind(molecule):
    """Make ``molecule`` the default for two_electron()."""
    _BOUND["molecule"] = molecule
    return device_basis(molecule)


_ONE_ELECTRON = {}


def one_electron_matrices(molecule):
    """(Core, Overlap) of the whole molecule from one kernel launch (cached per molecule)."""
    db = device_basis(molecule)
    ent = _ONE_ELECTRON.get(id(molecule))
    if ent is None or ent[0] is not db:
        Z = [float(a.NuclearCharge) for a in molecule.Atoms]
        R = [[float(x) for x in a.Coordinates] for a in molecule.Atoms]
        ent = (db,) + db.one_electron(Z, R)
        _ONE_ELECTRON.clear()
        _ONE_ELECTRON[id(molecule)] = ent
    return ent[1], ent[2]


def one_electron(molecule, shell_pair):
    """(core, overlap) blocks of one shell pair -- same signature and return value as
    integrals.one_electron (Methods/integrals.py:220-370)."""
    core, overlap = one_electron_matrices(molecule)
    ia, ib = shell_pair.Centre1.Ivec, shell_pair.Centre2.Ivec
    return core[np.ix_(ia, ib)].copy(), overlap[np.ix_(ia, ib)].copy()
