"""Host-side data model for the ERI / J-K hot path.

Mirrors the part of the reference's data model the hot path consumes, with the same attribute
names, so that code written against the reference's objects (and the reference's own
``Molecule`` instances, which are accepted everywhere by duck typing) keeps working:

  reference                                   here
  ---------                                   ----
  Util/structures.py:834-856 ContractedGaussian -> ContractedGaussian
  Util/structures.py:806-829 Atom               -> Atom
  Util/structures.py:962-968 Shell              -> Shell
  Util/structures.py:918-956 ShellPair          -> ShellPair (lazy: no host primitive tables,
                                                   the device builds them)
  Util/structures.py:319-523 Molecule           -> Molecule (geometry/basis part only)
  Util/structures.py:974-977 remove_punctuation -> remove_punctuation
  Data/constants.py:2,31                        -> TO_BOHR, INTEGRAL_THRESHOLD

Only what the two-electron path needs is kept; the SCF/NOCI/MP2 drivers are out of scope and
run unchanged from the reference.
"""
import json
import math
import os

import numpy as np

TO_BOHR = 1.8897161646320724        # Data/constants.py:2 (kept verbatim: parity depends on it)
INTEGRAL_THRESHOLD = 1.0e-8         # Data/constants.py:31
N_ELECTRONS = {"H": 1, "HE": 2, "LI": 3, "BE": 4, "B": 5, "C": 6, "N": 7, "O": 8, "F": 9, "NE": 10}

_BASIS_CACHE = None


def basis_library():
    """The shipped subset of basis-set data (written once in the authoring container from the reference's Data/basis.py)."""
    global _BASIS_CACHE
    if _BASIS_CACHE is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "basis_subset.json")
        with open(path) as fh:
            _BASIS_CACHE = json.load(fh)
    return _BASIS_CACHE


def remove_punctuation(basis_set):
    """'6-31G**' -> '631GSS' (Util/structures.py:974-977)."""
    out = basis_set.replace("*", "s")
    for ch in "-(),":
        out = out.replace(ch, "")
    return out.upper()


def n_cart(l):
    return (l + 1) * (l + 2) // 2


def cart_components(l):
    """Cartesian powers in the reference's order: lx descending, then ly descending
    (Methods/c_ints/two_electron_vrr.c:27-29)."""
    return [(lx, ly, l - lx - ly) for lx in range(l, -1, -1) for ly in range(l - lx, -1, -1)]


class ContractedGaussian:
    def __init__(self, function, cartesian_l=()):
        self.AngularMomentum = int(function[0])
        l = self.AngularMomentum
        self.NAngMomCart = n_cart(l)
        self.NAngMomSpher = 2 * l + 1
        self.Primitives = [list(p) for p in function[1:]]
        self.NPrimitives = len(self.Primitives)
        self.Exponents = np.array([p[0] for p in self.Primitives], dtype=float)
        self.DoubleExponents = 2.0 * self.Exponents
        # exponent-dependent part of the normalisation folded into the coefficients
        self.ScaledCCs = [cc * (2 * ex) ** ((l + 1.5) / 2.0) for ex, cc in self.Primitives]
        self.IsCartesian = l in cartesian_l
        self.NAngMom = self.NAngMomCart if self.IsCartesian else self.NAngMomSpher
        self.ContractionScaling = [
            (math.gamma(lx + 0.5) * math.gamma(ly + 0.5) * math.gamma(lz + 0.5)) ** -0.5
            for lx, ly, lz in cart_components(l)]


class Atom:
    def __init__(self, index, row, basis_set, cartesian_l=(), max_l=None, to_bohr=TO_BOHR):
        label, charge, x, y, z = row
        self.Index = index
        self.Label = label.upper()
        self.NuclearCharge = charge
        self.Coordinates = [x * to_bohr, y * to_bohr, z * to_bohr]
        lib = basis_library()
        if basis_set not in lib or self.Label not in lib[basis_set]:
            raise KeyError("basis %s / element %s is not in pychem_b200/data/basis_subset.json "
                           "(extend the element / basis lists of the extraction script and rerun it)" % (basis_set, self.Label))
        self.Basis = [ContractedGaussian(f, cartesian_l) for f in lib[basis_set][self.Label]
                      if max_l is None or f[0] <= max_l]
        self.NFunctions = sum(c.NAngMom for c in self.Basis)
        self.MaxAng = max(c.AngularMomentum for c in self.Basis)


class Shell:
    def __init__(self, coords, cgtf, index, index_vec):
        self.Coords = coords
        self.Cgtf = cgtf
        self.Index = index
        self.Ivec = index_vec


class ShellPair:
    """Geometry-only view of a shell pair.  ``Index1/Index2`` are the global shell indices the
    device tables are addressed with (the reference object carries per-atom indices only, see
    ``shell_index_of`` in integrals.py for how its instances are resolved)."""

    def __init__(self, coords_a, cgtf_a, ia, ia_vec, coords_b, cgtf_b, ib, ib_vec):
        self.Centre1 = Shell(coords_a, cgtf_a, ia, ia_vec)
        self.Centre2 = Shell(coords_b, cgtf_b, ib, ib_vec)
        self.Index1 = ia
        self.Index2 = ib
        self.Ltot = cgtf_a.AngularMomentum + cgtf_b.AngularMomentum


class _LazyShellPairs:
    """molecule.ShellPairs[(a, b)] for a <= b, built on demand (73 920 pairs at (H2O)32)."""

    def __init__(self, shells):
        self._shells = shells
        self._cache = {}

    def __getitem__(self, key):
        if key not in self._cache:
            a, b = key
            if not (0 <= a <= b < len(self._shells)):
                raise KeyError(key)
            (ca, ga, va), (cb, gb, vb) = self._shells[a], self._shells[b]
            self._cache[key] = ShellPair(ca, ga, a, va, cb, gb, b, vb)
        return self._cache[key]

    def __len__(self):
        n = len(self._shells)
        return n * (n + 1) // 2


class Molecule:
    """Geometry + basis.  ``coords`` rows are ``[symbol, Z, x, y, z]`` as in the reference's
    input format (Documentation/README.input); units Angstrom unless ``coords_units`` says
    BOHR/ATOMIC."""

    def __init__(self, coords, basis, charge=0, multiplicity=1, cartesian_l=(), max_l=None,
                 coords_units="ANGSTROM", allocate_tensor=False):
        self.Coords = coords
        self.NAtom = len(coords)
        self.Charge = charge
        self.Multiplicity = multiplicity
        self.CartesianL = list(cartesian_l)
        self.CoordsScaleFactor = TO_BOHR if coords_units.upper() == "ANGSTROM" else 1.0
        self.Basis = remove_punctuation(basis)
        self.Atoms = [Atom(i, row, self.Basis, self.CartesianL, max_l, self.CoordsScaleFactor)
                      for i, row in enumerate(coords)]
        self.NElectrons = sum(N_ELECTRONS[a.Label] for a in self.Atoms) - charge
        self.NAlphaElectrons = (self.NElectrons + (multiplicity - 1)) // 2
        self.NBetaElectrons = (self.NElectrons - (multiplicity - 1)) // 2
        self.NOrbitals = sum(a.NFunctions for a in self.Atoms)
        self.NCgtf = sum(len(a.Basis) for a in self.Atoms)
        shells = []
        count = 0
        for atom in self.Atoms:
            for cgtf in atom.Basis:
                shells.append((atom.Coordinates, cgtf, list(range(count, count + cgtf.NAngMom))))
                count += cgtf.NAngMom
        self.ShellPairs = _LazyShellPairs(shells)
        self.Bounds = [[0.0] * self.NCgtf for _ in range(self.NCgtf)]
        self.CoulombIntegrals = None
        if allocate_tensor:
            self.CoulombIntegrals = np.zeros((self.NOrbitals,) * 4)


# -------------------------------------------------------------------------------------------
# synthetic benchmark geometries
# -------------------------------------------------------------------------------------------
H2O_MONOMER = [["O", 8.0, 0.0, 0.0, 0.117790],
               ["H", 1.0, 0.0, 0.755453, -0.471161],
               ["H", 1.0, 0.0, -0.755453, -0.471161]]


def water_cluster(n, spacing=3.1):
    """(H2O)_n on a simple-cubic lattice, ceil(n^(1/3)) per side, filled x-fastest, no rotation."""
    side = 1
    while side ** 3 < n:
        side += 1
    coords = []
    for k in range(n):
        ix, iy, iz = k % side, (k // side) % side, k // (side * side)
        for sym, z, x, y, zz in H2O_MONOMER:
            coords.append([sym, z, x + ix * spacing, y + iy * spacing, zz + iz * spacing])
    return coords


def benzene():
    """D6h benzene, r(CC)=1.39 A, r(CH)=1.09 A, planar in xy."""
    coords = []
    for k in range(6):
        ang = math.pi / 3 * k
        coords.append(["C", 6.0, 1.39 * math.cos(ang), 1.39 * math.sin(ang), 0.0])
    for k in range(6):
        ang = math.pi / 3 * k
        coords.append(["H", 1.0, 2.48 * math.cos(ang), 2.48 * math.sin(ang), 0.0])
    return coords


def lih_chain(k, bond=2.2, spacing=4.4):
    coords = []
    for i in range(k):
        coords.append(["Li", 3.0, i * spacing, 0.0, 0.0])
        coords.append(["H", 1.0, i * spacing + bond, 0.0, 0.0])
    return coords
