/*
 * pychem_b200.h -- C ABI of the B200-native replacement for pychem's two-electron hot path.
 *
 * Plain C, caller-owned buffers, no torch/numpy types.  Every function returns 0 on success and
 * a non-zero code otherwise; pc_last_error() then describes the failure (CUDA error string or
 * argument problem).  Nothing here ever calls exit().  Buffers documented "host or device" are
 * classified with cudaPointerGetAttributes, so a torch tensor's data_ptr() can be passed as is.
 *
 * Each entry point names the reference interface it replaces (paths relative to the reference
 * checkout).  The reference's eleven fine-grained `_c_ints` functions (Methods/_c_ints.c:68-81)
 * are per-recursion-step micro-calls driven from Python; this ABI replaces their CALLERS
 * (SURVEY.md section 8(b)), which is where >90 % of the reference's time goes.
 */
#ifndef PYCHEM_B200_H
#define PYCHEM_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pc_basis pc_basis; /* opaque: shell table, shell-pair tables, Boys table, plans */

/* digestion variants for pc_jk_* (what make_coulomb_exchange_matrices is asked to do) */
#define PC_JK_AUTO 0 /* pc_jk_direct only: classify the densities on the device, pick the cheapest exact variant */
#define PC_JK_RHF 2 /* symmetric densities, D_alpha == D_beta (one exchange matrix computed)      */
#define PC_JK_UHF 3 /* symmetric densities, two spins                                             */
#define PC_JK_GEN 4 /* general non-symmetric densities (NOCI co-densities, Methods/noci.py:204-211) */

/* Last error message of the calling thread ("" if none). */
const char* pc_last_error(void);

/* Number of visible CUDA devices; fails (non-zero) when there is no usable device. */
int pc_device_count(int* count);

/*
 * Build the device-resident basis from a flat shell table.
 *   Replaces: ContractedGaussian constants (Util/structures.py:834-856), the eager ShellPair
 *   construction (Util/structures.py:510-523, 918-956) and _c_ints.shellpair_quantities
 *   (Methods/_c_ints.c:120-155, Methods/c_ints/shellpair_quantities.c:5-38).
 *   l[s] <= 3 (s, p, d, f); d shells either all spherical or all Cartesian (is_cart, Cartesian_L);
 *   f shells spherical.  Quartets of s/p/d shells run the generated class kernels, quartets with
 *   an f shell the generic kernel (csrc/pc_generic.cuh); g and higher are refused.
 *   scc = cc*(2a)^((l+1.5)/2) (Util/structures.py:843); first_fn = index of the shell's first
 *   basis function (atoms -> shells -> functions, Util/structures.py:511-520).
 */
int pc_basis_create(int device, int nshell, const int* l, const int* K, const int* is_cart,
                    const int* first_fn, const double* centres /*[nshell][3] bohr*/,
                    const double* exps, const double* scc, pc_basis** out);
int pc_basis_destroy(pc_basis* h);
/*
 * Integral type of the handle: 0 = electron repulsion (default), 1 = electron-scattering kernel at
 * grid value S -- the `ints_type, grid_value` arguments of integrals.two_electron /
 * hartree_fock.evaluate_2e_ints (Methods/_c_ints.c:236-240 -> two_electron_scattering.c,
 * spherical_bessel_j.c; caller: Methods/properties.py:19-23).  Switching invalidates the Schwarz
 * factors and the plan (they are recomputed from the new integrals, as the reference does).
 * pc_schwarz, pc_plan, pc_eri_quartets and pc_eri_tensor honour it; J/K digestion is type 0 only.
 */
int pc_basis_set_ints_type(pc_basis* h, int ints_type, double grid_value);
int pc_basis_nbf(const pc_basis* h, int* nbf);
/* cudaStream_t the handle launches on (for CUDA-event timing on the launching stream). */
int pc_basis_stream(const pc_basis* h, void** stream);

/*
 * Schwarz factors.  Replaces pass 1 of hartree_fock.evaluate_2e_ints
 * (Methods/hartree_fock.py:244-254): bounds[p][m*nfb+n] = sqrt((mn|mn)) for shell pair p=(a<=b)
 * in upper-triangular order p = a*nshell - a(a-1)/2 + (b-a); pmax[p] = max over the block.
 * Both outputs are host buffers (bounds: npair*49 doubles, pmax: npair doubles; either may be
 * NULL).  Must be called before any screened entry point; sorts the pair tables by pmax.
 */
int pc_schwarz(pc_basis* h, double* bounds, double* pmax);

/*
 * Build the screened, class-bucketed shell-quartet schedule.
 *   Replaces the loop nest + screen of pass 2 (Methods/hartree_fock.py:276-295): a unique
 *   quartet (ab|cd) survives iff max(B_ab)*max(B_cd) > thresh (strict), diagonal quartets
 *   (ab|ab) always do.  Static cost-balanced multi-GPU partition: the task range of every
 *   (bra bucket, ket bucket) is cut into nranks equal contiguous slices (cost is uniform inside a
 *   bucket pair, so the schedule is exactly balanced) and this handle keeps slice `rank`; all
 *   bucket pairs of one angular-momentum class run in ONE fused kernel launch.
 *   Outputs (may be NULL): quartets/eris kept by this rank, and in total.
 */
int pc_plan(pc_basis* h, double thresh, int rank, int nranks, long long* my_quartets,
            long long* my_eris, long long* all_quartets, long long* all_eris);

/*
 * Arbitrary shell quartets, for parity tests: out receives, for quartet q, the block
 * (nfa, nfb, nfc, nfd) in C order starting at out[offsets[q]] -- exactly what
 * integrals.two_electron(shell_pair1, shell_pair2, 0, -1.0) returns
 * (Methods/integrals.py:427-555).  abcd = [n][4] shell indices with a<=b, c<=d.  Host buffers.
 */
int pc_eri_quartets(pc_basis* h, int n, const int* abcd, const long long* offsets, double* out);

/*
 * Dense tensor G[N][N][N][N] (C order) with the reference's screening and 8-fold scatter.
 *   Replaces hartree_fock.evaluate_2e_ints pass 2 + the scatter loops
 *   (Methods/hartree_fock.py:266-273, 314-325).  Needs pc_schwarz + pc_plan(rank 0 of 1).
 *   G_dev: device buffer of N^4 doubles (zeroed here); G_host (may be NULL) receives a copy.
 */
int pc_eri_tensor(pc_basis* h, double* G_dev, double* G_host);

/*
 * J/K from the stored tensor: one streaming pass over G (HBM-bound).
 *   Replaces the three einsums of hartree_fock.make_coulomb_exchange_matrices
 *   (Methods/hartree_fock.py:345-347): J = einsum("cd,abcd->ab", Dt, G),
 *   Xa = einsum("cb,abcd->ad", -Da, G), Xb likewise (Exchange carries the minus sign).
 *   Dt/Da/Db/J/Xa/Xb: N*N doubles each, host or device.  Densities may be non-symmetric.
 */
int pc_jk_stored(pc_basis* h, const double* G_dev, const double* Dt, const double* Da,
                 const double* Db, double* J, double* Xa, double* Xb);

/*
 * Integral-direct J/K: ERIs of this rank's slice of the plan are regenerated and digested on
 * the fly (no N^4 store).  Same outputs as pc_jk_stored.  variant = PC_JK_RHF/UHF/GEN.
 *   pc_jk_direct_accumulate: acc_dev = device buffer of 3*N*N doubles (zeroed here) receiving
 *     this rank's partial half-accumulators [J | Ka | Kb]; sum it over ranks (NCCL all-reduce),
 *   pc_jk_finalize: turns the summed accumulators into J, Xa, Xb (symmetrisation and sign).
 *   pc_jk_direct: both steps on one GPU with an internal accumulator.
 */
/* variant 5 (measurement only): generate every ERI of the slice and discard it -- the pure
 * ERI-generation time of the same schedule, no digestion. */
#define PC_ERI_ONLY 5
int pc_jk_direct_accumulate(pc_basis* h, int variant, const double* Dt, const double* Da,
                            const double* Db, double* acc_dev);
int pc_jk_finalize(pc_basis* h, int variant, const double* acc_dev, double* J, double* Xa,
                   double* Xb);
/* pc_jk_classify + pc_jk_direct_accumulate with a single upload of host densities; *variant
 * receives the variant that was used (pass it to pc_jk_finalize). */
int pc_jk_direct_accumulate_auto(pc_basis* h, const double* Dt, const double* Da, const double* Db,
                                 double* acc_dev, int* variant);
/* pc_jk_direct with the variant picked on the device and reported in *variant.  When the densities
 * turn out closed-shell (PC_JK_RHF: symmetric, Da == Db) X_beta equals X_alpha and Xb is NOT
 * written: the caller aliases it (one device->host copy less per Fock build). */
int pc_jk_direct_auto(pc_basis* h, const double* Dt, const double* Da, const double* Db, double* J,
                      double* Xa, double* Xb, int* variant);
int pc_jk_direct(pc_basis* h, int variant, const double* Dt, const double* Da, const double* Db,
                 double* J, double* Xa, double* Xb);
/* variant (PC_JK_RHF/UHF/GEN) the densities allow: symmetric & Da==Db / symmetric / general.
 * Exact element-wise tests on the device; host or device inputs. */
int pc_jk_classify(pc_basis* h, const double* Dt, const double* Da, const double* Db, int* variant);

/*
 * Batched J/K for `nset` sets of general (possibly non-symmetric) densities in ONE pass over the
 * integrals: NOCI builds J/K for every determinant pair with its own co-density matrices
 * (Methods/noci.py:247,275,291 -> hartree_fock.make_coulomb_exchange_matrices per pair,
 * Methods/hartree_fock.py:329-347); in direct mode that would regenerate all ERIs per pair.
 *   D:   nset x [Dt | Da | Db], 3*N*N doubles per set, host or device
 *   out: nset x [J | Xa | Xb],  3*N*N doubles per set, host or device (Exchange carries the sign)
 *   pc_jk_stored_batch: streams the stored tensor once per 4 sets.
 *   pc_jk_direct_batch: every ERI block of this rank's slice is generated once and digested
 *     with all sets (general variant).  _accumulate/_finalize_batch split it for multi-GPU:
 *     acc_dev = device buffer of nset*3*N*N doubles to be summed over ranks in between.
 */
int pc_jk_stored_batch(pc_basis* h, const double* G_dev, int nset, const double* D, double* out);
int pc_jk_direct_batch(pc_basis* h, int nset, const double* D, double* out);
int pc_jk_direct_batch_accumulate(pc_basis* h, int nset, const double* D, double* acc_dev);
int pc_jk_finalize_batch(pc_basis* h, int nset, const double* acc_dev, double* out);

/*
 * Measurement hooks (bench.py): with profiling on, pc_jk_direct_accumulate brackets every kernel
 * launch (one fused launch per angular-momentum class; its time is booked on the first bucket
 * pair of the class) with CUDA events on the launching stream and synchronises at the end.  pc_plan_items reports, per item k: cls[4k..] = (lx1,ly1,lx2,ly2),
 * kprim[2k..] = primitive pairs per bra/ket shell pair, tasks[2k..] = (all quartets, this rank's
 * quartets), ms[k] = device time of the item in the last profiled accumulate, prim_exec[k] =
 * primitive quartets the whole bucket pair actually visits (after the primitive-pair cut-off;
 * kprim gives the unscreened contraction depths the reference loops over).  Arrays may be NULL.
 */
int pc_set_profiling(pc_basis* h, int on);
int pc_plan_items(pc_basis* h, int max_items, int* n_items, int* cls, int* kprim,
                  long long* tasks, float* ms, double* prim_exec);

/* Kernel launches issued by this handle since creation (bench.py's gpu_launches). */
int pc_launch_count(const pc_basis* h, long long* n);

/*
 * One-electron matrices (the step before the hot path): Core = kinetic energy +
 * nuclear attraction and Overlap, N x N each (host or device), for all shell pairs in one launch.
 *   Replaces hartree_fock.make_core_matrices' loop over integrals.one_electron
 *   (Methods/hartree_fock.py:207-222, Methods/integrals.py:220-370) and the five
 *   _c_ints.one_electron_* micro-step functions (Methods/_c_ints.c:400-640).
 *   Z[natom] nuclear charges, R[natom][3] nuclear positions in bohr.
 */
int pc_one_electron(pc_basis* h, int natom, const double* Z, const double* R, double* core,
                    double* overlap);

/*
 * MP2: AO->MO four-index transform on the FP64 tensor cores (DMMA, mma.sync m8n8k4 f64) and the
 * UMP2 energy sums.  Replaces the O(N^6) Python loops of Methods/mp2.py:37-94 (half transforms
 * :43-69, energy sums :77-94).  G_dev: the dense tensor from pc_eri_tensor (device, N^4 doubles);
 * Ca/Cb: MO coefficients (N x N, row-major, AO index first, as this_state.Alpha.MOs);
 * Ea/Eb: orbital energies; na/nb: molecule.NAlphaElectrons / NBetaElectrons (mp2.py:78-91);
 * same_spin = 0 skips the alpha-alpha and beta-beta sums ("P2-SOS", mp2.py:77).  Host or device
 * pointers for C and E.  Outputs: the three unscaled sums MP2_Eaa, MP2_Eab, MP2_Ebb.
 * Errors of these two-letter entry points are reported through pc_mp2_last_error().
 */
int pc_mp2_energy(int device, int N, const double* G_dev, const double* Ca, const double* Cb,
                  const double* Ea, const double* Eb, int na, int nb, int same_spin, double* Eaa,
                  double* Eab, double* Ebb);
const char* pc_mp2_last_error(void);
/* C (M x N) = A (M x K) . B (K x N), row-major device buffers, through the same DMMA kernel. */
/* frees the device scratch pc_mp2_energy keeps between calls (four transform intermediates) */
int pc_mp2_release(void);
int pc_dgemm_dmma(int device, int M, int N, int K, const double* A, const double* B, double* C);

/*
 * Test hook, pure host code (no device needed): the segment builder pc_plan applies to one
 * (bra bucket, ket bucket) -- the loop nest and Schwarz screen of Methods/hartree_fock.py:276-295.
 * pm_*: Schwarz maxima by position (groups of pairs sharing their primary shell, descending inside
 * a group, groups by descending maximum), gstart_*: [groups+1] group starts.  Returns per segment
 * seg_ij = (first bra pair | run << 24 | forced << 28, first ket pair), the exclusive prefixes
 * seg_off [n_seg+1] (tasks = ket pairs) and seg_quartets [n_seg+1].  A task's thread keeps the bra
 * pairs i of the run with pm_bra[i]*pm_ket[j] > thresh (and i <= j when same), or the single
 * forced pair.
 */
int pc_plan_segments_host(int nb, const double* pm_bra, int nbg, const int* gstart_bra, int nk,
                          const double* pm_ket, int nkg, const int* gstart_ket, int same, int run,
                          double thresh, int max_seg, int* n_seg, long long* seg_off, int* seg_ij,
                          long long* seg_quartets);

/* Register-resident DFMA micro-benchmark: measured FP64 FMA peak of `device` in TFLOP/s
 * (the roofline denominator for the ERI kernels; MEASURED_PEAKS.json has no FP64 entry). */
int pc_fp64_peak(int device, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* PYCHEM_B200_H */
