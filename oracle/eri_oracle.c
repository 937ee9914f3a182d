/*
 * eri_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference's (dlc62/pychem) two-electron hot path, used only by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as the CHECKER for the CUDA
 * path.  Nothing under pychem_b200/ may call into this file.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against (i) the golden
 * vectors in tests/golden/ that were minted by running the reference's own C extension + Python
 * driver (oracle/build_ref.py, oracle/make_golden.py) and (ii) when oracle/_ref is present, the
 * reference itself, shell quartet by shell quartet.
 *
 * Reference sites restated here (all paths relative to /root/reference):
 *   Methods/c_ints/shellpair_quantities.c:5-38      -> orc_pair_setup
 *   Methods/c_ints/two_electron_fundamentals.c:6-95 -> fundamentals()
 *   Methods/c_ints/interpolation_table.h:4-14       -> boys table (regenerated, see below)
 *   Methods/c_ints/two_electron_vrr.c:4-141         -> vrr_ket(), vrr_bra()
 *   Methods/c_ints/two_electron_contract.c:3-57     -> contraction inside orc_eri_quartet
 *   Methods/c_ints/two_electron_hrr.c:4-94          -> hrr()
 *   Methods/c_ints/angmom_index.c:3-15              -> cidx()
 *   Methods/integrals.py:427-555 (two_electron)     -> orc_eri_quartet
 *   Methods/hartree_fock.py:241-325                 -> orc_schwarz, orc_eri_tensor
 *   Methods/hartree_fock.py:329-347                 -> orc_jk
 *   Util/structures.py:834-856, 918-956             -> shell / shell-pair constants
 *   Data/transform_basis.py:3-30                    -> cart->spherical matrices (l<=3)
 *   Methods/c_ints/two_electron_scattering.c:6-84   -> fundamentals_scatter()  (ints_type = 1)
 *   Methods/c_ints/spherical_bessel_j.c:5-81        -> bessel_j()
 *
 * The Boys interpolation table is missing from the reference checkout (.MISSING_LARGE_BLOBS);
 * like oracle/build_ref.py we regenerate it: cubic in sT=T/(2d) per interval of width 2d from a
 * third-order Taylor expansion about the interval centre, d = 0.002, n_points = 7750.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define ORC_LMAX 3                 /* highest shell angular momentum handled by the oracle   */
#define ORC_KMAX 576               /* max primitive pairs per shell pair (24 x 24)             */
#define ORC_LPAIR (2 * ORC_LMAX)   /* highest pair angular momentum                          */
#define ORC_MMAX (4 * ORC_LMAX + 1)
#define NCART(l) ((((l) + 1) * ((l) + 2)) / 2)
#define NCUM(l) ((((l) + 1) * ((l) + 2) * ((l) + 3)) / 6) /* # cartesians with total <= l    */

/* ------------------------------------------------------------------------------------------ */
/* Boys table                                                                                 */
/* ------------------------------------------------------------------------------------------ */
#define TAB_N 7750
#define TAB_M (ORC_MMAX + 1)
static const double tab_d = 0.002;
static double *tab_f = NULL; /* [4][TAB_M][TAB_N] */

static void boys_ld(int mmax, long double T, long double *F) {
  /* series for the top order, then stable downward recursion */
  long double term = 1.0L / (2 * mmax + 1), acc = term;
  for (int k = 1; k < 600; k++) {
    term = term * (2 * T) / (2 * mmax + 2 * k + 1);
    acc += term;
    if (term < 1e-25L * acc) break;
  }
  long double eT = expl(-T);
  F[mmax] = eT * acc;
  for (int m = mmax; m > 0; m--) F[m - 1] = (2 * T * F[m] + eT) / (2 * m - 1);
}

static void build_table(void) {
  if (tab_f) return;
  tab_f = (double *)malloc(sizeof(double) * 4 * TAB_M * TAB_N);
  long double h = 2.0L * (long double)tab_d;
  long double F[TAB_M + 4];
  for (int j = 0; j < TAB_N; j++) {
    long double a = j + 0.5L;
    boys_ld(TAB_M + 3, a * h, F);
    for (int m = 0; m < TAB_M; m++) {
      long double c0 = F[m], c1 = -F[m + 1], c2 = F[m + 2] / 2, c3 = -F[m + 3] / 6;
      long double h2 = h * h, h3 = h * h * h;
      tab_f[(0 * TAB_M + m) * TAB_N + j] = (double)(c0 - c1 * h * a + c2 * h2 * a * a - c3 * h3 * a * a * a);
      tab_f[(1 * TAB_M + m) * TAB_N + j] = (double)(c1 * h - 2 * c2 * h2 * a + 3 * c3 * h3 * a * a);
      tab_f[(2 * TAB_M + m) * TAB_N + j] = (double)(c2 * h2 - 3 * c3 * h3 * a);
      tab_f[(3 * TAB_M + m) * TAB_N + j] = (double)(c3 * h3);
    }
  }
}

/* export for tests: returns coefficient k of order m at interval j */
double orc_boys_coeff(int k, int m, int j) {
  build_table();
  return tab_f[(k * TAB_M + m) * TAB_N + j];
}

/* ------------------------------------------------------------------------------------------ */
/* basis description                                                                          */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  int l, K, is_cart, first_fn, nfn;
  double A[3];
  const double *exps; /* [K] */
  const double *scc;  /* [K] cc*(2a)^((l+1.5)/2)   (structures.py:843) */
} orc_shell;

typedef struct {
  int nshell, nbf;
  orc_shell *sh;
  double *exps, *scc;
} orc_basis;

orc_basis *orc_basis_new(int nshell, const int *l, const int *K, const int *is_cart,
                         const int *first_fn, const double *centres, const double *exps,
                         const double *scc) {
  build_table();
  orc_basis *b = (orc_basis *)calloc(1, sizeof(orc_basis));
  b->nshell = nshell;
  b->sh = (orc_shell *)calloc(nshell, sizeof(orc_shell));
  int ntot = 0;
  for (int i = 0; i < nshell; i++) ntot += K[i];
  b->exps = (double *)malloc(sizeof(double) * ntot);
  b->scc = (double *)malloc(sizeof(double) * ntot);
  memcpy(b->exps, exps, sizeof(double) * ntot);
  memcpy(b->scc, scc, sizeof(double) * ntot);
  int off = 0;
  b->nbf = 0;
  for (int i = 0; i < nshell; i++) {
    orc_shell *s = &b->sh[i];
    if (l[i] > ORC_LMAX) { free(b->sh); free(b); return NULL; }
    s->l = l[i]; s->K = K[i]; s->is_cart = is_cart[i]; s->first_fn = first_fn[i];
    s->nfn = is_cart[i] ? NCART(l[i]) : 2 * l[i] + 1;
    for (int k = 0; k < 3; k++) s->A[k] = centres[3 * i + k];
    s->exps = b->exps + off; s->scc = b->scc + off;
    off += K[i];
    if (s->first_fn + s->nfn > b->nbf) b->nbf = s->first_fn + s->nfn;
  }
  return b;
}

void orc_basis_free(orc_basis *b) {
  if (!b) return;
  free(b->sh); free(b->exps); free(b->scc); free(b);
}
int orc_basis_nbf(const orc_basis *b) { return b->nbf; }

/* ------------------------------------------------------------------------------------------ */
/* cartesian bookkeeping                                                                      */
/* ------------------------------------------------------------------------------------------ */
/* position of (lx,ly,lz) inside its shell: lx descending, then ly descending
   (two_electron_vrr.c:27-29; angmom_index.c:3-15) */
static inline int cidx(int lx, int ly, int lz) { (void)lx; return (ly + lz) * (ly + lz + 1) / 2 + lz; }

static void cart_list(int l, int (*out)[3]) {
  int n = 0;
  for (int lx = l; lx >= 0; lx--)
    for (int ly = l - lx; ly >= 0; ly--) { out[n][0] = lx; out[n][1] = ly; out[n][2] = l - lx - ly; n++; }
}

/* cart -> real spherical (Data/transform_basis.py:3-30), rows = spherical functions */
static void c2s_matrix(int l, int is_cart, double *M /*[nfn][ncart]*/, int *nfn) {
  int nc = NCART(l);
  if (is_cart || l < 2) {
    *nfn = nc;
    for (int i = 0; i < nc * nc; i++) M[i] = 0.0;
    for (int i = 0; i < nc; i++) M[i * nc + i] = 1.0;
    return;
  }
  *nfn = 2 * l + 1;
  for (int i = 0; i < (*nfn) * nc; i++) M[i] = 0.0;
  if (l == 2) {
    /* cart order: xx xy xz yy yz zz */
    M[0 * 6 + 0] = sqrt(3.0) / 2; M[0 * 6 + 3] = -sqrt(3.0) / 2;
    M[1 * 6 + 1] = 1.0;
    M[2 * 6 + 2] = 1.0;
    M[3 * 6 + 4] = 1.0;
    M[4 * 6 + 0] = -0.5; M[4 * 6 + 3] = -0.5; M[4 * 6 + 5] = 1.0;
  } else if (l == 3) {
    /* cart order: xxx xxy xxz xyy xyz xzz yyy yyz yzz zzz */
    M[0 * 10 + 0] = sqrt(5.0 / 2) / 2; M[0 * 10 + 3] = -3 / (2 * sqrt(2.0));
    M[1 * 10 + 1] = 3 / (2 * sqrt(2.0)); M[1 * 10 + 6] = -sqrt(5.0 / 2) / 2;
    M[2 * 10 + 2] = sqrt(3.0) / 2; M[2 * 10 + 7] = -sqrt(3.0) / 2;
    M[3 * 10 + 4] = 1.0;
    M[4 * 10 + 0] = -sqrt(3.0 / 2) / 2; M[4 * 10 + 3] = -sqrt(3.0 / 10) / 2; M[4 * 10 + 5] = sqrt(6.0 / 5);
    M[5 * 10 + 1] = -sqrt(3.0 / 10) / 2; M[5 * 10 + 6] = -sqrt(3.0 / 2) / 2; M[5 * 10 + 8] = sqrt(6.0 / 5);
    M[6 * 10 + 2] = -3 / (2 * sqrt(5.0)); M[6 * 10 + 7] = -3 / (2 * sqrt(5.0)); M[6 * 10 + 9] = 1.0;
  }
}

/* angular part of the normalisation, (G(lx+1/2)G(ly+1/2)G(lz+1/2))^-1/2  (structures.py:850-856) */
static void cart_norms(int l, double *nm) {
  int c[28][3];
  cart_list(l, c);
  for (int i = 0; i < NCART(l); i++)
    nm[i] = 1.0 / sqrt(tgamma(c[i][0] + 0.5) * tgamma(c[i][1] + 0.5) * tgamma(c[i][2] + 0.5));
}

/* ------------------------------------------------------------------------------------------ */
/* primitive-pair quantities  (shellpair_quantities.c:23-36, structures.py:934-937)           */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  int K;                /* Ka*Kb, a-major */
  double sigma[ORC_KMAX], U[ORC_KMAX], P[ORC_KMAX][3], zeta[ORC_KMAX], cc[ORC_KMAX];
  double kappa[ORC_KMAX];     /* 2 * exponent of the SECONDARY (lower-l) centre, per primitive pair */
  double Rx[3];         /* primary - secondary centre */
} orc_pair;

/* X = primary (higher-l) shell, Y = secondary */
static void orc_pair_setup(const orc_shell *X, const orc_shell *Y, orc_pair *p) {
  double r2 = 0;
  for (int i = 0; i < 3; i++) { p->Rx[i] = X->A[i] - Y->A[i]; r2 += p->Rx[i] * p->Rx[i]; }
  p->K = X->K * Y->K;
  int n = 0;
  for (int ia = 0; ia < X->K; ia++)
    for (int ib = 0; ib < Y->K; ib++, n++) {
      double a = X->exps[ia], b = Y->exps[ib];
      double sigma = 1.0 / (a + b);
      p->sigma[n] = sigma;
      p->U[n] = pow(M_PI * sigma, 1.5) * exp(-a * b * sigma * r2);
      for (int i = 0; i < 3; i++) p->P[n][i] = (a * X->A[i] + b * Y->A[i]) * sigma;
      p->zeta[n] = 0.5 * sigma;
      p->kappa[n] = 2.0 * b;
      p->cc[n] = X->scc[ia] * Y->scc[ib];
    }
}

/* ------------------------------------------------------------------------------------------ */
/* fundamentals  [00|00]^(m)  (two_electron_fundamentals.c:41-89)                             */
/* ------------------------------------------------------------------------------------------ */
static void fundamentals(double sP, double UP, const double *P, double sQ, double UQ, const double *Q,
                         int lmax, double *F /*[lmax+1]*/, double *R /*[3]*/) {
  const double pf = pow(2 / M_PI, 0.5);
  double U = UP * UQ;
  double theta_sq = 1 / (sP + sQ);
  double two_theta_sq = 2 * theta_sq;
  double R2 = 0;
  for (int i = 0; i < 3; i++) { R[i] = P[i] - Q[i]; R2 += R[i] * R[i]; }
  if (R2 < 1.e-14) {
    for (int m = 0; m <= lmax; m++) F[m] = pf * U * pow(two_theta_sq, m + 0.5) * (1 / (2 * (double)m + 1));
  } else {
    double T = theta_sq * R2;
    double sT = T / (2 * tab_d);
    int j = (int)sT;
    if (j < TAB_N) {
      for (int m = lmax; m > -1; m--) {
        double f = tab_f[(0 * TAB_M + m) * TAB_N + j] +
                   sT * (tab_f[(1 * TAB_M + m) * TAB_N + j] +
                         sT * (tab_f[(2 * TAB_M + m) * TAB_N + j] + sT * tab_f[(3 * TAB_M + m) * TAB_N + j]));
        F[m] = pf * U * pow(two_theta_sq, m + 0.5) * f;
      }
    } else {
      for (int m = 0; m <= lmax; m++) {
        double f = tgamma(m + 0.5) / (2 * pow(T, m + 0.5));
        F[m] = pf * U * pow(two_theta_sq, m + 0.5) * f;
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* scattering fundamentals (ints_type = 1): U exp(-S^2/(4 theta^2)) S^(2m) z^-m j_m(z), z = S R  */
/* (two_electron_scattering.c:21-80, spherical_bessel_j.c:5-81)                                 */
/* ------------------------------------------------------------------------------------------ */
static int g_ints_type = 0;
static double g_grid = -1.0;
void orc_set_ints_type(int ints_type, double grid_value) { g_ints_type = ints_type; g_grid = grid_value; }

static void bessel_j(double *j, double z, int l_max) {
  const int l_thresh = 16;
  const double half_pi = 0.5 * M_PI;
  if (z < 1.e1) {                       /* series about z = 0 */
    int k_max = z < 1.e-3 ? 1 : (z < 1.e-1 ? 3 : (z < 1.e0 ? 6 : 20));
    double z2 = z * z, df = 1.0;
    for (int m = 0; m <= l_max; m++) {
      double mm1 = 2 * m + 1;
      df *= mm1;
      double jsum = 1.0, mm_df = 1.0, mm = mm1, zz = 1.0, sign = 1.0, denom = 1.0;
      for (int k = 1; k <= k_max; k++) {
        mm += 2; mm_df *= mm; zz *= z2; sign *= -1; denom *= 2 * k;
        jsum += sign * zz / (denom * mm_df);
      }
      j[m] = jsum / df;
    }
  } else if (z > 1.e2) {                /* asymptotic expansion */
    double zinv = 1 / z, zinv2 = zinv * zinv, zi = 1.0, zoff = z;
    j[0] = sin(z) * zinv;
    for (int m = 1; m <= l_max; m++) {
      double mm1 = (double)(m * (m + 1) / 2);
      zi *= zinv; zoff -= half_pi;
      j[m] = zi * (zinv * sin(zoff) + mm1 * zinv2 * cos(zoff));
    }
  } else {                              /* upward recursion */
    double jj[32];
    double zinv = 1 / z, zinv2 = zinv * zinv;
    jj[0] = sin(z) * zinv;
    jj[1] = (sin(z) - z * cos(z)) * zinv2;
    for (int m = 2; m <= l_thresh; m++) jj[m] = ((2 * m - 1) * jj[m - 1] * zinv - jj[m - 2]);
    zinv2 = 1.0;
    for (int m = 0; m <= l_thresh; m++) { jj[m] = jj[m] * zinv2; zinv2 *= zinv; }
    for (int m = 0; m <= l_max; m++) j[m] = m <= l_thresh ? jj[m] : 0.0;
  }
  for (int m = 0; m <= l_max; m++) if (fabs(j[m]) < 1.e-16) j[m] = 0.0;
}

static void fundamentals_scatter(double sP, double UP, const double *P, double sQ, double UQ, const double *Q,
                                 int lmax, double S, double *F, double *R) {
  double f[ORC_MMAX + 1], S2[ORC_MMAX + 1], j[ORC_MMAX + 1];
  double Ssq = S * S, df = 1.0, Spow = 1.0, quarter_Ssq = 0.25 * Ssq;
  f[0] = 1.0; S2[0] = 1.0;
  for (int m = 1; m <= lmax; m++) { Spow *= Ssq; df *= (double)(2 * m + 1); S2[m] = Spow; f[m] = 1.0 / df; }
  double U = UP * UQ;
  double theta_sq_inv = sP + sQ;
  double eS = exp(-quarter_Ssq * theta_sq_inv);
  double R2 = 0;
  for (int i = 0; i < 3; i++) { R[i] = P[i] - Q[i]; R2 += R[i] * R[i]; }
  if (S < 1.e-14) {
    F[0] = U;
    for (int m = 1; m <= lmax; m++) F[m] = 0;
  } else if (R2 < 1.e-14) {
    for (int m = 0; m <= lmax; m++) F[m] = U * eS * S2[m] * f[m];
  } else {
    bessel_j(j, S * sqrt(R2), lmax);
    for (int m = lmax; m > -1; m--) F[m] = U * eS * S2[m] * j[m];
  }
}

/* ------------------------------------------------------------------------------------------ */
/* VRR on one primitive quartet (two_electron_vrr.c:92-108, Gill-scaled form)                  */
/*   V[(a_cum * NCc + c_cum) * (L+1) + m],  a_cum/c_cum = cumulative cartesian index          */
/* ------------------------------------------------------------------------------------------ */
static inline int cum(int l, int i) { return NCUM(l - 1) + i; }

static void vrr_quartet(int La, int Lc, const orc_pair *bra, int ib, const orc_pair *ket, int ik,
                        const double *F, const double *R, double *V) {
  const int L = La + Lc, M1 = L + 1, NCc = NCUM(Lc);
  const double zeta = bra->zeta[ib], eta = ket->zeta[ik];
  int comp[28][3];
#define VV(a, c, m) V[((a) * NCc + (c)) * M1 + (m)]
  for (int m = 0; m <= L; m++) VV(0, 0, m) = F[m];
  /* step 2 of integrals.py:514 -- ket build on the s bra: sign_Rx=-1, sign_R=+1 */
  for (int lc = 1; lc <= Lc; lc++) {
    cart_list(lc, comp);
    for (int ic = 0; ic < NCART(lc); ic++) {
      int x = comp[ic][0], y = comp[ic][1], z = comp[ic][2];
      int dir = x ? 0 : (y ? 1 : 2);
      int d0[3] = {x, y, z};
      d0[dir] -= 1;
      int i0 = cum(lc - 1, cidx(d0[0], d0[1], d0[2]));
      int nval = d0[dir], i1 = -1;
      if (nval > 0) { int d1[3] = {d0[0], d0[1], d0[2]}; d1[dir] -= 1; i1 = cum(lc - 2, cidx(d1[0], d1[1], d1[2])); }
      double c0 = -ket->Rx[dir] * ket->kappa[ik] * eta;
      double c1 = R[dir] * eta;
      for (int m = 0; m <= L - lc; m++) {
        double v = c0 * VV(0, i0, m) + c1 * VV(0, i0, m + 1);
        if (i1 >= 0) v += nval * eta * (VV(0, i1, m) - eta * VV(0, i1, m + 1));
        VV(0, cum(lc, ic), m) = v;
      }
    }
  }
  /* step 1 -- bra build for every ket class: sign_Rx=-1, sign_R=-1 */
  for (int la = 1; la <= La; la++) {
    cart_list(la, comp);
    for (int ia = 0; ia < NCART(la); ia++) {
      int x = comp[ia][0], y = comp[ia][1], z = comp[ia][2];
      int dir = x ? 0 : (y ? 1 : 2);
      int d0[3] = {x, y, z};
      d0[dir] -= 1;
      int a0 = cum(la - 1, cidx(d0[0], d0[1], d0[2]));
      int aval = d0[dir], a1 = -1;
      if (aval > 0) { int d1[3] = {d0[0], d0[1], d0[2]}; d1[dir] -= 1; a1 = cum(la - 2, cidx(d1[0], d1[1], d1[2])); }
      double c0 = -bra->Rx[dir] * bra->kappa[ib] * zeta;
      double c1 = -R[dir] * zeta;
      int at = cum(la, ia);
      for (int lc = 0; lc <= Lc; lc++) {
        int kc[28][3];
        cart_list(lc, kc);
        for (int ic = 0; ic < NCART(lc); ic++) {
          int ct = cum(lc, ic);
          int cval = kc[ic][dir], c1i = -1;
          if (cval > 0) { int e[3] = {kc[ic][0], kc[ic][1], kc[ic][2]}; e[dir] -= 1; c1i = cum(lc - 1, cidx(e[0], e[1], e[2])); }
          for (int m = 0; m <= L - la - lc; m++) {
            double v = c0 * VV(a0, ct, m) + c1 * VV(a0, ct, m + 1);
            if (a1 >= 0) v += aval * zeta * (VV(a1, ct, m) - zeta * VV(a1, ct, m + 1));
            if (c1i >= 0) v += cval * zeta * eta * VV(a0, c1i, m + 1);
            VV(at, ct, m) = v;
          }
        }
      }
    }
  }
#undef VV
}

/* ------------------------------------------------------------------------------------------ */
/* HRR on contracted integrals (two_electron_hrr.c:82):  (x,y+1_i| = (x+1_i,y| + (X-Y)_i (x y| */
/* Works on one side; "other" is the flattened index of everything else.                      */
/*   in : S[e_cum_rel][other], e = lx..lx+ly (cumulative over those shells)                   */
/*   out: T[(ix*ncart(ly)+iy)][other]                                                         */
/* ------------------------------------------------------------------------------------------ */
static void hrr_side(int lx, int ly, const double *Rx, const double *S, int nother, double *T) {
  /* level k holds (e, k| for e = lx .. lx+ly-k, stored as blocks per e */
  int Ltot = lx + ly;
  size_t maxsz = 0;
  for (int k = 0; k <= ly; k++) {
    size_t sz = 0;
    for (int e = lx; e <= Ltot - k; e++) sz += (size_t)NCART(e) * NCART(k);
    if (sz > maxsz) maxsz = sz;
  }
  double *cur = (double *)malloc(sizeof(double) * maxsz * nother);
  double *nxt = (double *)malloc(sizeof(double) * maxsz * nother);
  size_t off_cur[2 * ORC_LPAIR + 2], off_nxt[2 * ORC_LPAIR + 2];
  /* level 0 = input */
  {
    size_t o = 0;
    for (int e = lx; e <= Ltot; e++) { off_cur[e] = o; o += (size_t)NCART(e); }
    memcpy(cur, S, sizeof(double) * o * nother);
  }
  int cy[28][3], cx[28][3];
  for (int k = 1; k <= ly; k++) {
    size_t o = 0;
    for (int e = lx; e <= Ltot - k; e++) { off_nxt[e] = o; o += (size_t)NCART(e) * NCART(k); }
    cart_list(k, cy);
    for (int e = lx; e <= Ltot - k; e++) {
      cart_list(e, cx);
      for (int ix = 0; ix < NCART(e); ix++)
        for (int iy = 0; iy < NCART(k); iy++) {
          int bx = cy[iy][0], by = cy[iy][1], bz = cy[iy][2];
          int dir = bx ? 0 : (by ? 1 : 2);
          int b0[3] = {bx, by, bz};
          b0[dir] -= 1;
          int a1[3] = {cx[ix][0], cx[ix][1], cx[ix][2]};
          a1[dir] += 1;
          int iy0 = cidx(b0[0], b0[1], b0[2]);
          int ix1 = cidx(a1[0], a1[1], a1[2]);
          const double *s0 = cur + (off_cur[e + 1] + (size_t)ix1 * NCART(k - 1) + iy0) * nother;
          const double *s1 = cur + (off_cur[e] + (size_t)ix * NCART(k - 1) + iy0) * nother;
          double *t = nxt + (off_nxt[e] + (size_t)ix * NCART(k) + iy) * nother;
          for (int q = 0; q < nother; q++) t[q] = s0[q] + Rx[dir] * s1[q];
        }
    }
    double *tmp = cur; cur = nxt; nxt = tmp;
    memcpy(off_cur, off_nxt, sizeof(off_cur));
  }
  memcpy(T, cur, sizeof(double) * (size_t)NCART(lx) * NCART(ly) * nother);
  free(cur); free(nxt);
}

/* ------------------------------------------------------------------------------------------ */
/* one contracted shell quartet (integrals.py:427-555); out[nfnA][nfnB][nfnC][nfnD]           */
/* ------------------------------------------------------------------------------------------ */
static void transpose(const double *in, int r, int c, double *out) {
  for (int i = 0; i < r; i++) for (int j = 0; j < c; j++) out[(size_t)j * r + i] = in[(size_t)i * c + j];
}

static void pair_transform(const orc_shell *A, const orc_shell *B, double *M, int *nsph, int *ncart) {
  /* normalised kron(c2sA, c2sB): rows (ma,mb), cols (ca,cb)   (structures.py:938-956, integrals.py:541-547) */
  double cA[7 * 10], cB[7 * 10], nA[10], nB[10];
  int nfa, nfb, nca = NCART(A->l), ncb = NCART(B->l);
  c2s_matrix(A->l, A->is_cart, cA, &nfa);
  c2s_matrix(B->l, B->is_cart, cB, &nfb);
  cart_norms(A->l, nA); cart_norms(B->l, nB);
  for (int ma = 0; ma < nfa; ma++) for (int mb = 0; mb < nfb; mb++)
    for (int ca = 0; ca < nca; ca++) for (int cb = 0; cb < ncb; cb++)
      M[(size_t)(ma * nfb + mb) * (nca * ncb) + ca * ncb + cb] = cA[ma * nca + ca] * cB[mb * ncb + cb] * (nA[ca] * nB[cb]);
  *nsph = nfa * nfb; *ncart = nca * ncb;
}

void orc_eri_quartet(const orc_basis *bs, int a, int b, int c, int d, double *out) {
  const orc_shell *A = &bs->sh[a], *B = &bs->sh[b], *C = &bs->sh[c], *D = &bs->sh[d];
  /* primary centre of each pair = higher l ("Goofy_bra/ket", integrals.py:79-89) */
  int gb = B->l > A->l, gk = D->l > C->l;
  const orc_shell *X1 = gb ? B : A, *Y1 = gb ? A : B, *X2 = gk ? D : C, *Y2 = gk ? C : D;
  orc_pair *bra = (orc_pair *)malloc(sizeof(orc_pair)), *ket = (orc_pair *)malloc(sizeof(orc_pair));
  orc_pair_setup(X1, Y1, bra);
  orc_pair_setup(X2, Y2, ket);
  const int lx1 = X1->l, ly1 = Y1->l, lx2 = X2->l, ly2 = Y2->l;
  const int La = lx1 + ly1, Lc = lx2 + ly2, L = La + Lc;
  const int NCa = NCUM(La), NCc = NCUM(Lc);
  /* contracted (e0|f0), e = lx1..La, f = lx2..Lc, m = 0 (integrals.py:601-612) */
  const int ea0 = NCUM(lx1 - 1), fc0 = NCUM(lx2 - 1);
  const int ne = NCa - ea0, nf = NCc - fc0;
  double *V = (double *)malloc(sizeof(double) * NCa * NCc * (L + 1));
  double *S = (double *)calloc((size_t)ne * nf, sizeof(double));
  double F[ORC_MMAX + 1], R[3];
  for (int ib = 0; ib < bra->K; ib++)
    for (int ik = 0; ik < ket->K; ik++) {
      if (g_ints_type == 1)
        fundamentals_scatter(bra->sigma[ib], bra->U[ib], bra->P[ib], ket->sigma[ik], ket->U[ik], ket->P[ik], L, g_grid, F, R);
      else
        fundamentals(bra->sigma[ib], bra->U[ib], bra->P[ib], ket->sigma[ik], ket->U[ik], ket->P[ik], L, F, R);
      vrr_quartet(La, Lc, bra, ib, ket, ik, F, R, V);
      double w = bra->cc[ib] * ket->cc[ik];
      for (int e = 0; e < ne; e++)
        for (int f = 0; f < nf; f++) S[(size_t)e * nf + f] += w * V[((size_t)(ea0 + e) * NCc + (fc0 + f)) * (L + 1)];
    }
  /* ket HRR first (on transposed arrays), then bra HRR (integrals.py:531-536) */
  const int nk = NCART(lx2) * NCART(ly2), nb = NCART(lx1) * NCART(ly1);
  double *St = (double *)malloc(sizeof(double) * ne * nf);
  transpose(S, ne, nf, St);                                   /* [f][e] */
  double *T1 = (double *)malloc(sizeof(double) * (size_t)nk * ne);
  hrr_side(lx2, ly2, ket->Rx, St, ne, T1);                    /* [(x2,y2)][e] */
  double *T1t = (double *)malloc(sizeof(double) * (size_t)nk * ne);
  transpose(T1, nk, ne, T1t);                                 /* [e][(x2,y2)] */
  double *T2 = (double *)malloc(sizeof(double) * (size_t)nb * nk);
  hrr_side(lx1, ly1, bra->Rx, T1t, nk, T2);                   /* [(x1,y1)][(x2,y2)] */
  /* put each pair back into (first shell, second shell) order */
  const int nca = NCART(A->l), ncb = NCART(B->l), ncc = NCART(C->l), ncd = NCART(D->l);
  double *G = (double *)malloc(sizeof(double) * (size_t)nb * nk);
  for (int ia = 0; ia < nca; ia++) for (int ibb = 0; ibb < ncb; ibb++)
    for (int ic = 0; ic < ncc; ic++) for (int id = 0; id < ncd; id++) {
      int rb = gb ? (ibb * nca + ia) : (ia * ncb + ibb);
      int rk = gk ? (id * ncc + ic) : (ic * ncd + id);
      G[(size_t)(ia * ncb + ibb) * nk + (ic * ncd + id)] = T2[(size_t)rb * nk + rk];
    }
  /* normalise + cart->spherical: cs_P . (ints o outer(nm_P,nm_Q)) . cs_Q^T  (integrals.py:541-547) */
  double *MP = (double *)malloc(sizeof(double) * 49 * 100), *MQ = (double *)malloc(sizeof(double) * 49 * 100);
  int nsP, ncP, nsQ, ncQ;
  pair_transform(A, B, MP, &nsP, &ncP);
  pair_transform(C, D, MQ, &nsQ, &ncQ);
  double *H = (double *)calloc((size_t)ncP * nsQ, sizeof(double));
  for (int i = 0; i < ncP; i++) for (int q = 0; q < nsQ; q++) {
    double s = 0;
    for (int j = 0; j < ncQ; j++) s += G[(size_t)i * ncQ + j] * MQ[(size_t)q * ncQ + j];
    H[(size_t)i * nsQ + q] = s;
  }
  for (int p = 0; p < nsP; p++) for (int q = 0; q < nsQ; q++) {
    double s = 0;
    for (int i = 0; i < ncP; i++) s += MP[(size_t)p * ncP + i] * H[(size_t)i * nsQ + q];
    out[(size_t)p * nsQ + q] = s;
  }
  free(bra); free(ket); free(V); free(S); free(St); free(T1); free(T1t); free(T2); free(G); free(MP); free(MQ); free(H);
}

/* ------------------------------------------------------------------------------------------ */
/* Schwarz factors, per shell pair a<=b: sqrt((mn|mn))  (hartree_fock.py:244-254)              */
/*   bounds: [npair][49] row-major (nfa x nfb used), pair index p = a*nshell - a(a-1)/2 + b-a  */
/*   pmax:   [npair] max over the block                                                       */
/* ------------------------------------------------------------------------------------------ */
static size_t pair_index(int n, int a, int b) { return (size_t)a * n - (size_t)a * (a - 1) / 2 + (b - a); }

void orc_schwarz(const orc_basis *bs, double *bounds, double *pmax) {
  int n = bs->nshell;
  double *blk = (double *)malloc(sizeof(double) * 49 * 49);
  for (int a = 0; a < n; a++)
    for (int b = a; b < n; b++) {
      int na = bs->sh[a].nfn, nb = bs->sh[b].nfn;
      orc_eri_quartet(bs, a, b, a, b, blk);
      size_t p = pair_index(n, a, b);
      double mx = 0;
      for (int m = 0; m < na; m++) for (int q = 0; q < nb; q++) {
        double d = blk[(((size_t)m * nb + q) * na + m) * nb + q];
        double v = d > 0 ? sqrt(d) : 0.0;   /* scattering diagonals can dip below zero: numpy gives nan, the block is skipped either way */
        bounds[p * 49 + m * nb + q] = v;
        if (v > mx) mx = v;
      }
      pmax[p] = mx;
    }
  free(blk);
}

/* ------------------------------------------------------------------------------------------ */
/* dense tensor with the reference's screening and 8-fold scatter (hartree_fock.py:241-325)   */
/* returns the number of off-diagonal shell quartets that survived the screen                 */
/* ------------------------------------------------------------------------------------------ */
static void scatter8(const orc_basis *bs, int a, int b, int c, int d, const double *blk, double *G) {
  const orc_shell *A = &bs->sh[a], *B = &bs->sh[b], *C = &bs->sh[c], *D = &bs->sh[d];
  size_t N = bs->nbf;
  for (int m = 0; m < A->nfn; m++) for (int n = 0; n < B->nfn; n++)
    for (int l = 0; l < C->nfn; l++) for (int s = 0; s < D->nfn; s++) {
      double v = blk[(((size_t)m * B->nfn + n) * C->nfn + l) * D->nfn + s];
      size_t i = A->first_fn + m, j = B->first_fn + n, k = C->first_fn + l, q = D->first_fn + s;
      G[((i * N + j) * N + k) * N + q] = v; G[((j * N + i) * N + k) * N + q] = v;
      G[((i * N + j) * N + q) * N + k] = v; G[((j * N + i) * N + q) * N + k] = v;
      G[((k * N + q) * N + i) * N + j] = v; G[((k * N + q) * N + j) * N + i] = v;
      G[((q * N + k) * N + i) * N + j] = v; G[((q * N + k) * N + j) * N + i] = v;
    }
}

long orc_eri_tensor(const orc_basis *bs, double thresh, double *G) {
  int n = bs->nshell;
  size_t npair = (size_t)n * (n + 1) / 2;
  double *bounds = (double *)calloc(npair * 49, sizeof(double));
  double *pmax = (double *)calloc(npair, sizeof(double));
  double *blk = (double *)malloc(sizeof(double) * 49 * 49);
  long nsurv = 0;
  orc_schwarz(bs, bounds, pmax);
  for (int a = 0; a < n; a++)
    for (int b = a; b < n; b++) { orc_eri_quartet(bs, a, b, a, b, blk); scatter8(bs, a, b, a, b, blk, G); }
  for (int a = 0; a < n; a++)
    for (int b = a; b < n; b++)
      for (int c = a; c < n; c++)
        for (int d = c; d < n; d++) {
          if (a == c && b == d) continue;
          /* amax(outer(B_ab, B_cd)) = max(B_ab)*max(B_cd); strict > (hartree_fock.py:293-294) */
          if (pmax[pair_index(n, a, b)] * pmax[pair_index(n, c, d)] > thresh) {
            orc_eri_quartet(bs, a, b, c, d, blk);
            scatter8(bs, a, b, c, d, blk, G);
            nsurv++;
          }
        }
  free(bounds); free(pmax); free(blk);
  return nsurv;
}

/* ------------------------------------------------------------------------------------------ */
/* J/K  (hartree_fock.py:345-347): J = einsum("cd,abcd->ab", Dt, G);                          */
/*      Xa = einsum("cb,abcd->ad", -Da, G); Xb likewise.  Densities may be non-symmetric.     */
/* ------------------------------------------------------------------------------------------ */
void orc_jk(int N, const double *G, const double *Dt, const double *Da, const double *Db,
            double *J, double *Xa, double *Xb) {
  size_t n = N;
  memset(J, 0, sizeof(double) * n * n);
  memset(Xa, 0, sizeof(double) * n * n);
  memset(Xb, 0, sizeof(double) * n * n);
  for (size_t a = 0; a < n; a++)
    for (size_t b = 0; b < n; b++)
      for (size_t c = 0; c < n; c++) {
        const double *g = G + ((a * n + b) * n + c) * n;
        double j = 0, da = Da[c * n + b], db = Db[c * n + b];
        for (size_t d = 0; d < n; d++) {
          j += Dt[c * n + d] * g[d];
          Xa[a * n + d] -= da * g[d];
          Xb[a * n + d] -= db * g[d];
        }
        J[a * n + b] += j;
      }
}

/* ------------------------------------------------------------------------------------------ */
/* Sampled J blocks and X rows for molecules whose N^4 tensor cannot be stored ((H2O)16/32).   */
/* Same semantics as orc_eri_tensor + orc_jk restricted to the requested outputs: a block      */
/* (ab|cd) of the reference's tensor is non-zero iff (ab) == (cd) or max(B_ab) max(B_cd) >     */
/* thresh (hartree_fock.py:244-250, 293-294; the 8-fold scatter :314-325 makes that symmetric  */
/* in all index permutations), and                                                             */
/*   J[m,n]  =  sum_{l,s} Dt[l,s] G[m,n,l,s]       for (m,n) in the listed shell pairs  (:345)  */
/*   X[m,s]  = -sum_{n,l} D[l,n]  G[m,n,l,s]       for m in the listed shells, all s    (:346)  */
/* J, Xa, Xb are N x N, zeroed here; only the sampled blocks / rows are filled (J also at the  */
/* transposed position).  pmax: Schwarz maxima from orc_schwarz.  Ket pairs are spread over    */
/* `nthreads` POSIX threads (the reference is single-threaded; this is only the checker).      */
/* ------------------------------------------------------------------------------------------ */
#include <pthread.h>

typedef struct {
  const orc_basis *bs;
  double thresh;
  const double *pmax, *Dt, *Da, *Db;
  int a, b, is_k;            /* J block (a,b) or X rows of shell a */
  volatile int *next;        /* shared counter over the first ket shell c */
  pthread_mutex_t *mu;
  double *jblk;              /* [49] shared */
  double *xa, *xb;           /* [7][N] shared */
  long *nq;
} orc_job;

static void *orc_sample_worker(void *arg) {
  orc_job *jb = (orc_job *)arg;
  const orc_basis *bs = jb->bs;
  const int n = bs->nshell, a = jb->a;
  const size_t N = bs->nbf;
  const orc_shell *A = &bs->sh[a];
  double loc[49], *blk = (double *)malloc(sizeof(double) * 49 * 49);
  double *ra = NULL, *rb = NULL;
  long my = 0;
  memset(loc, 0, sizeof(loc));
  if (jb->is_k) { ra = (double *)calloc(7 * N, sizeof(double)); rb = (double *)calloc(7 * N, sizeof(double)); }
  for (;;) {
    const int c = __sync_fetch_and_add(jb->next, 1);
    if (c >= n) break;
    for (int d = c; d < n; d++) {
      const size_t pcd = pair_index(n, c, d);
      const orc_shell *C = &bs->sh[c], *D = &bs->sh[d];
      if (!jb->is_k) {
        const int b = jb->b;
        const orc_shell *B = &bs->sh[b];
        const size_t pab = pair_index(n, a < b ? a : b, a < b ? b : a);
        if (!(pcd == pab || jb->pmax[pab] * jb->pmax[pcd] > jb->thresh)) continue;
        orc_eri_quartet(bs, a, b, c, d, blk);
        my++;
        for (int m = 0; m < A->nfn; m++) for (int q = 0; q < B->nfn; q++) {
          double s = 0;
          for (int l = 0; l < C->nfn; l++) for (int r = 0; r < D->nfn; r++) {
            const double g = blk[(((size_t)m * B->nfn + q) * C->nfn + l) * D->nfn + r];
            double dd = jb->Dt[(C->first_fn + l) * N + D->first_fn + r];
            if (c != d) dd += jb->Dt[(D->first_fn + r) * N + C->first_fn + l];
            s += dd * g;
          }
          loc[m * B->nfn + q] += s;
        }
      } else {
        for (int b = 0; b < n; b++) {
          const size_t pab = pair_index(n, a < b ? a : b, a < b ? b : a);
          if (!(pcd == pab || jb->pmax[pab] * jb->pmax[pcd] > jb->thresh)) continue;
          const orc_shell *B = &bs->sh[b];
          orc_eri_quartet(bs, a, b, c, d, blk);
          my++;
          for (int m = 0; m < A->nfn; m++) for (int q = 0; q < B->nfn; q++)
            for (int l = 0; l < C->nfn; l++) for (int r = 0; r < D->nfn; r++) {
              const double g = blk[(((size_t)m * B->nfn + q) * C->nfn + l) * D->nfn + r];
              const size_t fb = B->first_fn + q, fc = C->first_fn + l, fd = D->first_fn + r;
              /* (m b | c d): X[m,d] -= D[c,b] g ; its image (m b | d c): X[m,c] -= D[d,b] g */
              ra[m * N + fd] -= jb->Da[fc * N + fb] * g;
              rb[m * N + fd] -= jb->Db[fc * N + fb] * g;
              if (c != d) {
                ra[m * N + fc] -= jb->Da[fd * N + fb] * g;
                rb[m * N + fc] -= jb->Db[fd * N + fb] * g;
              }
            }
        }
      }
    }
  }
  pthread_mutex_lock(jb->mu);
  if (!jb->is_k) {
    for (int k = 0; k < 49; k++) jb->jblk[k] += loc[k];
  } else {
    for (size_t k = 0; k < (size_t)A->nfn * N; k++) { jb->xa[k] += ra[k]; jb->xb[k] += rb[k]; }
  }
  *jb->nq += my;
  pthread_mutex_unlock(jb->mu);
  free(blk); free(ra); free(rb);
  return NULL;
}

void orc_jk_sample(const orc_basis *bs, double thresh, const double *pmax, int nJ, const int *jab,
                   int nK, const int *ka, const double *Dt, const double *Da, const double *Db,
                   double *J, double *Xa, double *Xb, int nthreads, long *nquartets) {
  const size_t N = bs->nbf;
  memset(J, 0, sizeof(double) * N * N);
  memset(Xa, 0, sizeof(double) * N * N);
  memset(Xb, 0, sizeof(double) * N * N);
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 64) nthreads = 64;
  long nq = 0;
  pthread_mutex_t mu;
  pthread_mutex_init(&mu, NULL);
  double *xa = (double *)malloc(sizeof(double) * 7 * N), *xb = (double *)malloc(sizeof(double) * 7 * N);
  for (int t = 0; t < nJ + nK; t++) {
    double jblk[49];
    volatile int next = 0;
    orc_job jb;
    memset(jblk, 0, sizeof(jblk));
    memset(xa, 0, sizeof(double) * 7 * N);
    memset(xb, 0, sizeof(double) * 7 * N);
    jb.bs = bs; jb.thresh = thresh; jb.pmax = pmax; jb.Dt = Dt; jb.Da = Da; jb.Db = Db;
    jb.is_k = t >= nJ;
    jb.a = jb.is_k ? ka[t - nJ] : jab[2 * t];
    jb.b = jb.is_k ? 0 : jab[2 * t + 1];
    jb.next = &next; jb.mu = &mu; jb.jblk = jblk; jb.xa = xa; jb.xb = xb; jb.nq = &nq;
    pthread_t th[64];
    for (int k = 0; k < nthreads; k++) pthread_create(&th[k], NULL, orc_sample_worker, &jb);
    for (int k = 0; k < nthreads; k++) pthread_join(th[k], NULL);
    const orc_shell *A = &bs->sh[jb.a];
    if (!jb.is_k) {
      const orc_shell *B = &bs->sh[jb.b];
      for (int m = 0; m < A->nfn; m++) for (int q = 0; q < B->nfn; q++) {
        J[(A->first_fn + m) * N + B->first_fn + q] = jblk[m * B->nfn + q];
        J[(B->first_fn + q) * N + A->first_fn + m] = jblk[m * B->nfn + q];
      }
    } else {
      for (int m = 0; m < A->nfn; m++) for (size_t s = 0; s < N; s++) {
        Xa[(A->first_fn + m) * N + s] = xa[m * N + s];
        Xb[(A->first_fn + m) * N + s] = xb[m * N + s];
      }
    }
  }
  free(xa); free(xb);
  pthread_mutex_destroy(&mu);
  if (nquartets) *nquartets = nq;
}
