"""Batched mirror of the reference's NOCI driver, Methods/noci.py:46-122 (`noci.do`).

The reference builds the CI matrix pair by pair and calls
``hf.make_coulomb_exchange_matrices(molecule, codensity_state)`` once per determinant pair
(noci.py:247, 275, 291): n(n+1)/2 Fock-like builds with non-symmetric co-densities, each of which
regenerates every ERI in direct mode.  This mirror keeps the reference's arithmetic and its
helper functions (biorthogonalize, process_overlaps, make_weighted_density, resize_array,
inner_product, CoDensityState, reorder_orbitals are CALLED from the reference's own module) and
only reorders the driver loop into three passes:

  1. per pair: Lowdin pairing, overlaps, zero list, co-density state     (noci.py:58-87, 242-246,
                                                                           265-273, 287-290)
  2. ONE batched J/K over all pairs' co-density states                    (noci.py:247, 275, 291 ->
     pychem_b200.hartree_fock.make_coulomb_exchange_matrices_batch)
  3. per pair: matrix element from J/K, reduced overlap, nuclear term     (noci.py:249-262, 277-283,
                                                                           293-304, 95-99)

followed by the same generalised eigenproblem and the same printed output (noci.py:101-122).
The unused N^4 Python loop of two_zeros (noci.py:296-300, its result `active_coulomb` is never
read) is not reproduced, so the mirror also works in direct mode where
``molecule.CoulombIntegrals`` is None.  Activate with ``install(Methods.noci)``.
"""
import numpy as np

from . import hartree_fock as hf_gpu

_REF = None


class _Pair:
    __slots__ = ("i", "j", "state_overlap", "reduced_overlap", "num_zeros", "zeros", "state",
                 "W_alpha", "W_beta", "P_alpha", "P_beta", "alpha_overlaps", "beta_overlaps",
                 "alpha_core", "beta_core")


def _setup_pair(ref, molecule, i, j, state1, state2):
    """noci.py:61-87 and the co-density construction of no_zeros / one_zero / two_zeros."""
    Spin = ref.Spin
    nA, nB = molecule.NAlphaElectrons, molecule.NBetaElectrons
    p = _Pair()
    p.i, p.j = i, j
    alpha = ref.biorthogonalize(state1.Alpha.MOs, state2.Alpha.MOs, molecule.Overlap, nA)
    beta = ref.biorthogonalize(state1.Beta.MOs, state2.Beta.MOs, molecule.Overlap, nB)
    p.alpha_core = alpha[0].T.dot(molecule.Core).dot(alpha[1])
    p.beta_core = beta[0].T.dot(molecule.Core).dot(beta[1])
    alpha_overlaps = np.diagonal(alpha[0].T.dot(molecule.Overlap).dot(alpha[1]))
    beta_overlaps = np.diagonal(beta[0].T.dot(molecule.Overlap).dot(beta[1]))
    p.state_overlap = np.prod(alpha_overlaps) * np.prod(beta_overlaps)
    reduced, zeros = ref.process_overlaps(1, [], alpha_overlaps, Spin.Alpha)
    reduced, zeros = ref.process_overlaps(reduced, zeros, beta_overlaps, Spin.Beta)
    if nA > nB:
        beta_overlaps = ref.resize_array(beta_overlaps, alpha_overlaps, fill=1)
        beta[0] = ref.resize_array(beta[0], alpha[0])
        beta[1] = ref.resize_array(beta[1], alpha[1])
    p.reduced_overlap, p.zeros, p.num_zeros = reduced, zeros, len(zeros)
    p.alpha_overlaps, p.beta_overlaps = alpha_overlaps, beta_overlaps
    p.state = None
    p.W_alpha = p.W_beta = p.P_alpha = p.P_beta = None
    N = molecule.NOrbitals
    if p.num_zeros <= 1:
        p.W_alpha = ref.make_weighted_density(alpha, alpha_overlaps)
        p.W_beta = ref.make_weighted_density(beta, beta_overlaps)
        p.state = ref.CoDensityState(N, p.W_alpha, p.W_beta)
        if p.num_zeros == 1:
            k = zeros[0].index
            p.P_alpha = np.outer(alpha[0][:, k], alpha[1][:, k])
            p.P_beta = np.outer(beta[0][:, k], beta[1][:, k])
    elif p.num_zeros == 2:
        k = zeros[0].index
        p.P_alpha = np.outer(alpha[0][:, k], alpha[1][:, k])
        p.P_beta = np.outer(beta[0][:, k], beta[1][:, k])
        p.state = ref.CoDensityState(N, p.P_alpha, p.P_beta)
    return p


def _element(ref, molecule, p):
    """The matrix element of one pair from its digested co-density state."""
    const = ref.const
    Spin = ref.Spin
    ip = ref.inner_product
    if p.num_zeros == 0:                                   # noci.py:249-262
        st = p.state
        elem = ip(p.W_alpha + p.W_beta, st.Total.Coulomb)
        elem += ip(p.W_alpha, st.Alpha.Exchange)
        elem += ip(p.W_beta, st.Beta.Exchange)
        elem *= 0.5
        for k in range(molecule.NAlphaElectrons):
            if p.alpha_overlaps[k] > const.NOCI_thresh:
                elem += p.alpha_core[k, k] / p.alpha_overlaps[k]
        for k in range(molecule.NBetaElectrons):
            if p.beta_overlaps[k] > const.NOCI_thresh:
                elem += p.beta_core[k, k] / p.beta_overlaps[k]
        return elem
    if p.num_zeros == 1:                                   # noci.py:277-283
        zero = p.zeros[0]
        st = p.state
        is_alpha = zero.spin == Spin.Alpha
        active_exchange = st.Alpha.Exchange if is_alpha else st.Beta.Exchange
        P_active = p.P_alpha if is_alpha else p.P_beta
        active_core = p.alpha_core if is_alpha else p.beta_core
        elem = ip(P_active, st.Total.Coulomb) + ip(P_active, active_exchange)
        elem += active_core[zero.index, zero.index]
        return elem
    if p.num_zeros == 2:                                   # noci.py:293-304
        _, spin = p.zeros[0]
        st = p.state
        is_alpha = spin == Spin.Alpha
        active_exchange = st.Alpha.Exchange if is_alpha else st.Beta.Exchange
        active_P = p.P_alpha if is_alpha else p.P_beta
        return ip(active_P, st.Total.Coulomb) + ip(active_P, active_exchange)
    return 0                                               # noci.py:92-93


def ci_matrices(molecule, ref=None):
    """(CI_matrix, CI_overlap) of molecule.States -- noci.py:52-99 with one batched J/K."""
    ref = ref or _REF
    if ref is None:
        raise RuntimeError("pychem_b200.noci: call install(Methods.noci) first")
    dims = len(molecule.States)
    CI_matrix = np.zeros((dims, dims))
    CI_overlap = np.zeros((dims, dims))
    pairs = []
    for i, state1 in enumerate(molecule.States):
        for j, state2 in enumerate(molecule.States[:i + 1]):
            pairs.append(_setup_pair(ref, molecule, i, j, state1, state2))
    hf_gpu.make_coulomb_exchange_matrices_batch(molecule, [p.state for p in pairs if p.state is not None])
    for p in pairs:
        elem = _element(ref, molecule, p)
        elem *= p.reduced_overlap
        elem += molecule.NuclearRepulsion * p.state_overlap
        CI_matrix[p.i, p.j] = CI_matrix[p.j, p.i] = elem
        CI_overlap[p.i, p.j] = CI_overlap[p.j, p.i] = p.state_overlap
    return CI_matrix, CI_overlap


def do(settings, molecule):
    """Drop-in for Methods/noci.py:46 `do(settings, molecule)`: same molecule attributes
    (NOCIEnergies, NOCIWavefunction, reordered States for spin-flip excitations) and the same
    text written to settings.OutFile."""
    ref = _REF
    if ref is None:
        raise RuntimeError("pychem_b200.noci: call install(Methods.noci) first")
    if "SF" in molecule.ExcitationType:
        ref.reorder_orbitals(molecule)
    CI_matrix, CI_overlap = ci_matrices(molecule, ref)
    energies, wavefunctions = ref.gen_eig(CI_matrix, CI_overlap)
    molecule.NOCIEnergies = energies
    molecule.NOCIWavefunction = wavefunctions
    printf = ref.printf
    printf.delimited_text(settings.OutFile, " NOCI output ")
    printf.text_value(settings.OutFile, " States ", wavefunctions, " NOCI Energies ", energies)
    printf.text_value(settings.OutFile, " Hamiltonian ", CI_matrix, " State overlaps ", CI_overlap)


def install(reference_noci):
    """Rebind `do` inside the reference's Methods.noci module (pychem.py:129 calls noci.do).
    Returns a callable that undoes the patch."""
    global _REF
    saved, saved_ref = reference_noci.do, _REF
    _REF = reference_noci
    reference_noci.do = do

    def uninstall():
        global _REF
        reference_noci.do = saved
        _REF = saved_ref
    return uninstall
