"""GPU-backed mirrors of the two hot functions of the reference's Methods/hartree_fock.py.

    evaluate_2e_ints(molecule, ints_type=0, grid_value=-1.0)      (hartree_fock.py:241-325)
    make_coulomb_exchange_matrices(molecule, this)                 (hartree_fock.py:329-347)
    make_coulomb_exchange_matrices_batch(molecule, states)         (the same for many states at once,
                                                                    used by pychem_b200.noci)

Same names, arguments and attribute side effects, so the reference's SCF driver
(hartree_fock.do), NOCI (noci.py:247,275,291) and MP2 (mp2.py:46) run unchanged once
``install()`` has rebound the two names in the reference's module:

  * ``molecule.Bounds[a][b]``   per-function Schwarz factors sqrt((mn|mn))      (:249-254)
  * ``molecule.CoulombIntegrals`` dense (N,N,N,N) ndarray incl. screening zeros   (:266-325)
  * ``this.Total.Coulomb``, ``this.Alpha.Exchange``, ``this.Beta.Exchange`` fresh N x N arrays,
    Exchange carrying the minus sign, for possibly non-symmetric densities        (:345-347)

Two modes, chosen per molecule (override with the environment variable PYCHEM_B200_MODE=
stored|direct, no new input keyword is needed):
  stored  N^4 fits the budget: the tensor lives in HBM, J/K is one streaming pass over it.
  direct  otherwise: ERIs are regenerated and digested on the fly every Fock build;
          ``molecule.CoulombIntegrals`` is then left as None (mp2/properties need `stored`).
There is no CPU fallback: without a CUDA device / the built library these functions raise.
"""
import os

import numpy as np

from . import engine
from .integrals import device_basis

STORED_LIMIT_BYTES = int(float(os.environ.get("PYCHEM_B200_STORED_LIMIT_GB", "24")) * 2 ** 30)

_STATE = {}     # id(molecule) -> dict(mode=..., G_dev=..., db=...)


def _mode_for(molecule):
    forced = os.environ.get("PYCHEM_B200_MODE", "").lower()
    if forced in ("stored", "direct"):
        return forced
    return "stored" if 8 * int(molecule.NOrbitals) ** 4 <= STORED_LIMIT_BYTES else "direct"


def evaluate_2e_ints(molecule, ints_type=0, grid_value=-1.0):
    if ints_type not in (0, 1):
        raise ValueError("evaluate_2e_ints: ints_type must be 0 (repulsion) or 1 (scattering)")
    db = device_basis(molecule)
    # ints_type 1: scattering kernel at grid_value (Methods/properties.py:19-23); Schwarz factors,
    # screening and the dense tensor are then those of the scattering integrals, as in the reference
    db.set_ints_type(ints_type, grid_value)
    table = db.table
    bounds, _ = db.schwarz()
    if not isinstance(molecule.Bounds, list) or len(molecule.Bounds) != table.nshell:
        molecule.Bounds = [[0.0] * table.nshell for _ in range(table.nshell)]
    p = 0
    for a in range(table.nshell):
        na = int(table.nfn[a])
        for b in range(a, table.nshell):
            nb = int(table.nfn[b])
            molecule.Bounds[a][b] = bounds[p, :na * nb].reshape(na, nb).copy()
            p += 1
    mode = _mode_for(molecule)
    if ints_type == 1:
        if 8 * int(molecule.NOrbitals) ** 4 > STORED_LIMIT_BYTES:
            raise MemoryError("scattering integrals are consumed as a dense tensor (properties.py:22); "
                              "N^4 exceeds PYCHEM_B200_STORED_LIMIT_GB")
        mode = "stored"
    import weakref
    try:
        mref = weakref.ref(molecule)
    except TypeError:
        mref = (lambda m: (lambda: m))(molecule)
    st = {"mode": mode, "db": db, "G_dev": None, "molecule": mref,
          "key": _quick_key(molecule)}
    if mode == "stored":
        G_dev, G_host = db.eri_tensor(engine.INTEGRAL_THRESHOLD, to_host=True)
        st["G_dev"] = G_dev
        molecule.CoulombIntegrals = G_host
    else:
        # one process per GPU: every rank keeps its slice of the quartet schedule and the partial
        # J/K are all-reduced inside make_coulomb_exchange_matrices (engine.DeviceBasis.jk_direct)
        rank, world = 0, 1
        try:
            import torch.distributed as tdist
            if tdist.is_available() and tdist.is_initialized():
                rank, world = tdist.get_rank(), tdist.get_world_size()
        except ImportError:
            pass
        db.plan(engine.INTEGRAL_THRESHOLD, rank, world)
        molecule.CoulombIntegrals = None
    _STATE[id(molecule)] = st
    _evict_other_molecules(molecule)


def _evict_other_molecules(molecule, keep=2):
    """The dense tensor of one molecule can be tens of GB of HBM: keep device state only for the
    `keep` most recently evaluated molecules (pychem.main walks the input sections one by one)."""
    order = _STATE.setdefault("__order__", [])
    key = id(molecule)
    if key in order:
        order.remove(key)
    order.append(key)
    while len(order) > keep:
        old = order.pop(0)
        st = _STATE.pop(old, None)
        if st is not None:
            st["G_dev"] = None
            from . import integrals
            integrals.release_key(old)


def _state_for(molecule):
    """The device state of `molecule`, (re)built by evaluate_2e_ints when it is missing, belongs
    to an earlier geometry / basis of the same object (integrals.device_basis hands out a new
    DeviceBasis then) or was left on the scattering integrals by a property job."""
    st = _STATE.get(id(molecule))
    # Per Fock build only the cheap part of the check: same object, live handle, repulsion integrals,
    # same basis label and size, same coordinate checksum.  The full stamp (integrals._stamp: 0.17 ms
    # of host time at 96 atoms, paid before anything is queued on the device) is compared where the
    # reference recomputes its integrals too: in evaluate_2e_ints, which hartree_fock.do calls once
    # per SCF (Methods/hartree_fock.py:32).
    if (st is None or st["molecule"]() is not molecule or st["db"].h is None or st["db"].ints_type != 0
            or st["key"] != _quick_key(molecule)):
        evaluate_2e_ints(molecule)
        st = _STATE[id(molecule)]
    return st


def _quick_key(molecule):
    """Basis label, sizes and a weighted checksum of the coordinates: ~30 us at 96 atoms, catches a
    geometry edited in place (Atom.update_coords, Util/structures.py:827) between Fock builds."""
    chk = 0.0
    for k, a in enumerate(molecule.Atoms):
        c = a.Coordinates
        chk += (k + 1) * (c[0] + 2.0 * c[1] + 3.0 * c[2])
    return (getattr(molecule, "Basis", None), int(molecule.NOrbitals), len(molecule.Atoms), chk)


def make_coulomb_exchange_matrices(molecule, this):
    st = _state_for(molecule)
    db = st["db"]
    Dt = np.ascontiguousarray(this.Total.Density, dtype=np.float64)
    Da = np.ascontiguousarray(this.Alpha.Density, dtype=np.float64)
    Db = np.ascontiguousarray(this.Beta.Density, dtype=np.float64)
    if st["mode"] == "stored":
        J, Xa, Xb = db.jk_stored(st["G_dev"], Dt, Da, Db)
    else:
        J, Xa, Xb = db.jk_direct(Dt, Da, Db)
    this.Total.Coulomb = J
    this.Alpha.Exchange = Xa
    this.Beta.Exchange = Xb


def make_coulomb_exchange_matrices_batch(molecule, states):
    """make_coulomb_exchange_matrices for a LIST of state-like objects in one pass over the
    integrals.  NOCI calls make_coulomb_exchange_matrices once per determinant
    pair with non-symmetric co-densities (Methods/noci.py:247,275,291); in direct mode every call
    regenerates all ERIs, in stored mode every call streams the N^4 tensor.  Here all pairs share
    one ERI generation (direct) or one tensor pass per four pairs (stored).  Same side effects
    per state as the single call: fresh ``Total.Coulomb``, ``Alpha.Exchange``, ``Beta.Exchange``."""
    states = list(states)
    if not states:
        return
    st = _state_for(molecule)
    db = st["db"]
    N = int(molecule.NOrbitals)
    D = np.empty((len(states), 3, N, N))
    for k, this in enumerate(states):
        D[k, 0] = this.Total.Density
        D[k, 1] = this.Alpha.Density
        D[k, 2] = this.Beta.Density
    out = db.jk_stored_batch(st["G_dev"], D) if st["mode"] == "stored" else db.jk_direct_batch(D)
    for k, this in enumerate(states):
        this.Total.Coulomb = np.array(out[k, 0])
        this.Alpha.Exchange = np.array(out[k, 1])
        this.Beta.Exchange = np.array(out[k, 2])


def make_core_matrices(molecule):
    """molecule.Core / Overlap from the device (one launch for all shell pairs), then the same
    canonical-orthogonalisation data the reference derives from the overlap matrix
    (hartree_fock.py:207-237): X, Xt (eigenvalues above Data/constants.linear_dependence = 1e-6)
    and the half-overlap matrix S."""
    from .integrals import one_electron_matrices
    core, overlap = one_electron_matrices(molecule)
    molecule.Core = np.array(core)
    molecule.Overlap = np.array(overlap)
    evals, evecs = np.linalg.eigh(molecule.Overlap)
    order = evals.argsort()[::-1]
    evals, evecs = evals[order], evecs[:, order]
    keep = evals > 1.0e-6
    U = evecs[:, keep]
    inv_root = evals[keep] ** -0.5
    molecule.X = U * inv_root
    molecule.Xt = molecule.X.T
    molecule.S = (U * (1.0 / inv_root)).dot(U.T)


def install(reference_hartree_fock, reference_noci=None, one_electron=False):
    """Rebind the two hot functions inside the reference's own modules (the drop-in).

    ``reference_hartree_fock`` is the reference's ``Methods.hartree_fock`` module;
    noci.py calls ``hf.make_coulomb_exchange_matrices`` through that module object, so
    rebinding there covers NOCI as well.  With ``one_electron=True`` make_core_matrices
    (hartree_fock.py:207-237) is rebound too.  Returns a callable that undoes the patch."""
    saved = (reference_hartree_fock.evaluate_2e_ints,
             reference_hartree_fock.make_coulomb_exchange_matrices)
    reference_hartree_fock.evaluate_2e_ints = evaluate_2e_ints
    reference_hartree_fock.make_coulomb_exchange_matrices = make_coulomb_exchange_matrices
    saved_core = reference_hartree_fock.make_core_matrices
    if one_electron:
        reference_hartree_fock.make_core_matrices = make_core_matrices

    def uninstall():
        reference_hartree_fock.evaluate_2e_ints, reference_hartree_fock.make_coulomb_exchange_matrices = saved
        reference_hartree_fock.make_core_matrices = saved_core
    return uninstall


def release(molecule=None):
    keys = [k for k in _STATE if k != "__order__"] if molecule is None else [id(molecule)]
    for k in keys:
        _STATE.pop(k, None)
        if k in _STATE.get("__order__", []):
            _STATE["__order__"].remove(k)
