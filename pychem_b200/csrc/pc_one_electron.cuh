// pc_one_electron.cuh -- overlap, kinetic-energy and nuclear-attraction integrals on the device
// (the step BEFORE the hot path).
//
// Replaces integrals.one_electron (Methods/integrals.py:220-370) and its five per-step C kernels
// (Methods/c_ints/one_electron_{fundamentals,vrr,hrr,kinetic,contract}.c) for all shell pairs at
// once: one thread per shell pair a <= b, loops over its primitive pairs and over the nuclei.
// Same quantities as the reference -- primitive prefactor U = (pi sigma)^1.5 exp(-ab sigma r^2)
// (shellpair_quantities.c:30), nuclear fundamentals -Z sqrt(2/pi) U sqrt(2 zeta) F_m(zeta R_PC^2)
// with the same three Boys branches (one_electron_fundamentals.c:47-91), kinetic energy from the
// overlap integrals with b +- 2 (one_electron_kinetic.c:62-66), normalisation and cart->spherical
// as in integrals.py:345-357 -- evaluated with Hermite expansion coefficients instead of the
// reference's class-by-class VRR/HRR (O(N^2 Natom) work: clarity over speed here).
#pragma once
#include "pc_common.cuh"

struct PcShellTable {
  const int* l;         // [nshell]
  const int* K;         // [nshell]
  const int* poff;      // [nshell] offset into exps / scc
  const int* first_fn;  // [nshell]
  const double* A;      // [nshell][3]
  const double* exps;
  const double* scc;    // cc * (2a)^((l+1.5)/2)
  const int* pair_a;    // [npair]
  const int* pair_b;    // [npair]
  int npair;
  int cart_d;
  int nbf;
};

__device__ __forceinline__ int pc1_ncart(int l) { return (l + 1) * (l + 2) / 2; }

// k-th Cartesian component of shell l in the reference's order (lx descending, then ly)
__device__ __forceinline__ void pc1_comp(int l, int k, int& lx, int& ly, int& lz) {
  int idx = 0;
  for (int x = l; x >= 0; --x)
    for (int y = l - x; y >= 0; --y, ++idx)
      if (idx == k) { lx = x; ly = y; lz = l - x - y; return; }
  lx = ly = lz = 0;
}

// packed index of the Hermite triple (t, u, v): triples ordered by their total, then like the
// Cartesian components of a shell
__device__ __forceinline__ int pc1_tuv(int t, int u, int v) {
  const int tot = t + u + v, s = u + v;
  return tot * (tot + 1) * (tot + 2) / 6 + s * (s + 1) / 2 + v;
}

// (Gamma(n + 1/2))^-1/2 for n = 0..3   (Util/structures.py:850-856)
__device__ __forceinline__ double pc1_gnorm(int n) {
  return n == 0 ? 0.75112554446494248 : (n == 1 ? 1.0622519320271969 : (n == 2 ? 0.86732507058407751 : 0.5485445389623983));
}

// unscaled Boys function F_m(T), m = 0..L, with the reference's three branches
__device__ __forceinline__ void pc1_boys(int L, double T, double R2, const double* __restrict__ boys, double* F) {
  if (R2 < 1.e-14) {
    for (int m = 0; m <= L; ++m) F[m] = 1.0 / (2.0 * m + 1.0);
    return;
  }
  const double sT = T * PC_BOYS_INV_2D;
  if (sT < (double)PC_BOYS_NPOINTS) {
    const int j = (int)sT;
    for (int m = 0; m <= L; ++m) {
      const double* c = boys + ((size_t)m * PC_BOYS_NPOINTS + j) * 4;
      F[m] = fma(sT, fma(sT, fma(sT, c[3], c[2]), c[1]), c[0]);
    }
  } else {
    const double rT = 1.0 / T;
    double f = 0.88622692545275801365 * sqrt(rT);
    for (int m = 0; m <= L; ++m) {
      F[m] = f;
      f *= (m + 0.5) * rT;
    }
  }
}

// Hermite expansion coefficients of one dimension: E[i][j][t], i <= la, j <= lbmax
// (LM = highest shell angular momentum the instantiation serves: 2 for s, p, d; 3 with f shells)
template <int LM>
__device__ __forceinline__ void pc1_hermite(int la, int lbmax, double PA, double PB, double inv2p,
                                            double (&E)[LM + 1][LM + 3][2 * LM + 4]) {
  for (int i = 0; i < LM + 1; ++i)
    for (int j = 0; j < LM + 3; ++j)
      for (int t = 0; t < 2 * LM + 4; ++t) E[i][j][t] = 0.0;
  E[0][0][0] = 1.0;
  for (int i = 0; i < la; ++i)
    for (int t = 0; t <= i + 1; ++t) {
      double v = PA * E[i][0][t] + (t + 1) * E[i][0][t + 1];
      if (t > 0) v += inv2p * E[i][0][t - 1];
      E[i + 1][0][t] = v;
    }
  for (int i = 0; i <= la; ++i)
    for (int j = 0; j < lbmax; ++j)
      for (int t = 0; t <= i + j + 1; ++t) {
        double v = PB * E[i][j][t] + (t + 1) * E[i][j][t + 1];
        if (t > 0) v += inv2p * E[i][j][t - 1];
        E[i][j + 1][t] = v;
      }
}

template <int LM>
__global__ void __launch_bounds__(64) one_electron_kernel(PcShellTable S, int natom, const double* __restrict__ Z,
                                                         const double* __restrict__ Rc,
                                                         const double* __restrict__ boys,
                                                         double* __restrict__ core, double* __restrict__ overlap) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= S.npair) return;
  int a = S.pair_a[p], b = S.pair_b[p];
  // the reference puts the higher angular momentum first ("Goofy", integrals.py:243-252);
  // here the order only decides which index the +-2 of the kinetic energy acts on -- both give
  // the same integral, we keep (a, b) as given
  const int la = S.l[a], lb = S.l[b];
  const int na = pc1_ncart(la), nb = pc1_ncart(lb);
  const double Ax = S.A[3 * a], Ay = S.A[3 * a + 1], Az = S.A[3 * a + 2];
  const double Bx = S.A[3 * b], By = S.A[3 * b + 1], Bz = S.A[3 * b + 2];
  const double r2 = (Ax - Bx) * (Ax - Bx) + (Ay - By) * (Ay - By) + (Az - Bz) * (Az - Bz);
  const int L = la + lb;
  constexpr int NC2 = ((LM + 1) * (LM + 2) / 2) * ((LM + 1) * (LM + 2) / 2);     // Cartesian block
  constexpr int NR = 2 * LM + 1;
  double Sc[NC2], Tc[NC2], Vc[NC2];
  for (int k = 0; k < NC2; ++k) Sc[k] = Tc[k] = Vc[k] = 0.0;
  double Ex[LM + 1][LM + 3][2 * LM + 4], Ey[LM + 1][LM + 3][2 * LM + 4], Ez[LM + 1][LM + 3][2 * LM + 4];
  constexpr int NTUV = NR * (NR + 1) * (NR + 2) / 6;       // (t, u, v) with t + u + v <= 2 LM
  double Rt[NR][NTUV];           // R^n_{tuv}, n + t + u + v <= L, (t, u, v) packed by pc1_tuv
  for (int ia = 0; ia < S.K[a]; ++ia)
    for (int ib = 0; ib < S.K[b]; ++ib) {
      const double al = S.exps[S.poff[a] + ia], be = S.exps[S.poff[b] + ib];
      const double w = S.scc[S.poff[a] + ia] * S.scc[S.poff[b] + ib];
      const double sigma = 1.0 / (al + be);
      const double zeta = al + be;
      const double U = pow(M_PI * sigma, 1.5) * exp(-al * be * sigma * r2);     // incl. (pi/p)^1.5
      const double Px = (al * Ax + be * Bx) * sigma, Py = (al * Ay + be * By) * sigma, Pz = (al * Az + be * Bz) * sigma;
      const double inv2p = 0.5 * sigma;
      pc1_hermite<LM>(la, lb + 2, Px - Ax, Px - Bx, inv2p, Ex);
      pc1_hermite<LM>(la, lb + 2, Py - Ay, Py - By, inv2p, Ey);
      pc1_hermite<LM>(la, lb + 2, Pz - Az, Pz - Bz, inv2p, Ez);
      // ---- overlap and kinetic energy (one_electron_kinetic.c:62-66) ----
      for (int ka = 0; ka < na; ++ka) {
        int ax, ay, az;
        pc1_comp(la, ka, ax, ay, az);
        for (int kb = 0; kb < nb; ++kb) {
          int bx, by, bz;
          pc1_comp(lb, kb, bx, by, bz);
          const double sx = Ex[ax][bx][0], sy = Ey[ay][by][0], sz = Ez[az][bz][0];
          const double s = sx * sy * sz;
          double t = be * (2 * lb + 3) * s;
          t -= 2.0 * be * be * (Ex[ax][bx + 2][0] * sy * sz + sx * Ey[ay][by + 2][0] * sz + sx * sy * Ez[az][bz + 2][0]);
          if (bx > 1) t -= 0.5 * bx * (bx - 1) * Ex[ax][bx - 2][0] * sy * sz;
          if (by > 1) t -= 0.5 * by * (by - 1) * sx * Ey[ay][by - 2][0] * sz;
          if (bz > 1) t -= 0.5 * bz * (bz - 1) * sx * sy * Ez[az][bz - 2][0];
          Sc[ka * nb + kb] += w * U * s;
          Tc[ka * nb + kb] += w * U * t;
        }
      }
      // ---- nuclear attraction: -Z sqrt(2/pi) U sqrt(2 zeta) sum_tuv E E E R_tuv ----
      const double spf = 0.79788456080286536 * U * sqrt(2.0 * zeta);
      for (int c = 0; c < natom; ++c) {
        const double X = Px - Rc[3 * c], Y = Py - Rc[3 * c + 1], Zc = Pz - Rc[3 * c + 2];
        const double R2 = X * X + Y * Y + Zc * Zc;
        double F[NR];
        pc1_boys(L, zeta * R2, R2, boys, F);
        double m2p = 1.0;
        for (int n = 0; n <= L; ++n) {
          Rt[n][0] = m2p * F[n];
          m2p *= -2.0 * zeta;
        }
        for (int tot = 1; tot <= L; ++tot)
          for (int t = 0; t <= tot; ++t)
            for (int u = 0; u <= tot - t; ++u) {
              const int v = tot - t - u;
              for (int n = 0; n <= L - tot; ++n) {
                double val;
                if (t > 0) val = (t > 1 ? (t - 1) * Rt[n + 1][pc1_tuv(t - 2, u, v)] : 0.0) + X * Rt[n + 1][pc1_tuv(t - 1, u, v)];
                else if (u > 0) val = (u > 1 ? (u - 1) * Rt[n + 1][pc1_tuv(t, u - 2, v)] : 0.0) + Y * Rt[n + 1][pc1_tuv(t, u - 1, v)];
                else val = (v > 1 ? (v - 1) * Rt[n + 1][pc1_tuv(t, u, v - 2)] : 0.0) + Zc * Rt[n + 1][pc1_tuv(t, u, v - 1)];
                Rt[n][pc1_tuv(t, u, v)] = val;
              }
            }
        const double pf = -Z[c] * spf * w;
        for (int ka = 0; ka < na; ++ka) {
          int ax, ay, az;
          pc1_comp(la, ka, ax, ay, az);
          for (int kb = 0; kb < nb; ++kb) {
            int bx, by, bz;
            pc1_comp(lb, kb, bx, by, bz);
            double sum = 0.0;
            for (int t = 0; t <= ax + bx; ++t)
              for (int u = 0; u <= ay + by; ++u)
                for (int v = 0; v <= az + bz; ++v)
                  sum += Ex[ax][bx][t] * Ey[ay][by][u] * Ez[az][bz][v] * Rt[0][pc1_tuv(t, u, v)];
            Vc[ka * nb + kb] += pf * sum;
          }
        }
      }
    }
  // ---- angular normalisation, cart -> spherical, scatter (integrals.py:345-357) ----
  // nuclear fundamentals of the reference carry (pi sigma)^1.5 inside U and the Hermite sum gives
  // (2 pi / zeta) K_ab ... : -Z sqrt(2/pi) (pi sigma)^1.5 sqrt(2 zeta) = -Z 2 pi / zeta, as it must
  double core_c[NC2], ov_c[NC2];
  for (int ka = 0; ka < na; ++ka) {
    int ax, ay, az;
    pc1_comp(la, ka, ax, ay, az);
    const double nma = pc1_gnorm(ax) * pc1_gnorm(ay) * pc1_gnorm(az);
    for (int kb = 0; kb < nb; ++kb) {
      int bx, by, bz;
      pc1_comp(lb, kb, bx, by, bz);
      const double nm = nma * pc1_gnorm(bx) * pc1_gnorm(by) * pc1_gnorm(bz);
      core_c[ka * nb + kb] = nm * (Tc[ka * nb + kb] + Vc[ka * nb + kb]);
      ov_c[ka * nb + kb] = nm * Sc[ka * nb + kb];
    }
  }
  // spherical d (Data/transform_basis.py:8-12), rows over cart xx xy xz yy yz zz
  const double C2S[5][6] = {{0.86602540378443865, 0, 0, -0.86602540378443865, 0, 0},
                            {0, 1, 0, 0, 0, 0}, {0, 0, 1, 0, 0, 0}, {0, 0, 0, 0, 1, 0},
                            {-0.5, 0, 0, -0.5, 0, 1}};
  // spherical f (Data/transform_basis.py:13-19), rows over cart xxx xxy xxz xyy xyz xzz yyy yyz yzz zzz
  const double C2SF[7][10] = {
      {0.7905694150420949, 0, 0, -1.0606601717798212, 0, 0, 0, 0, 0, 0},
      {0, 1.0606601717798212, 0, 0, 0, 0, -0.7905694150420949, 0, 0, 0},
      {0, 0, 0.8660254037844386, 0, 0, 0, 0, -0.8660254037844386, 0, 0},
      {0, 0, 0, 0, 1, 0, 0, 0, 0, 0},
      {-0.6123724356957945, 0, 0, -0.27386127875258304, 0, 1.0954451150103321, 0, 0, 0, 0},
      {0, -0.27386127875258304, 0, 0, 0, 0, -0.6123724356957945, 0, 1.0954451150103321, 0},
      {0, 0, -0.6708203932499369, 0, 0, 0, 0, -0.6708203932499369, 0, 1}};
  const bool sa = (la == 2 && !S.cart_d) || (LM >= 3 && la == 3), sb = (lb == 2 && !S.cart_d) || (LM >= 3 && lb == 3);
  const int nfa = sa ? 2 * la + 1 : na, nfb = sb ? 2 * lb + 1 : nb;
  const int fa = S.first_fn[a], fb = S.first_fn[b];
  for (int ma = 0; ma < nfa; ++ma)
    for (int mb = 0; mb < nfb; ++mb) {
      double cv = 0.0, sv = 0.0;
      for (int ka = 0; ka < na; ++ka) {
        const double ca = sa ? ((LM >= 3 && la == 3) ? C2SF[ma][ka] : C2S[ma][ka]) : (ka == ma ? 1.0 : 0.0);
        if (ca == 0.0) continue;
        for (int kb = 0; kb < nb; ++kb) {
          const double cb = sb ? ((LM >= 3 && lb == 3) ? C2SF[mb][kb] : C2S[mb][kb]) : (kb == mb ? 1.0 : 0.0);
          if (cb == 0.0) continue;
          cv += ca * cb * core_c[ka * nb + kb];
          sv += ca * cb * ov_c[ka * nb + kb];
        }
      }
      const size_t i = fa + ma, j = fb + mb;
      core[i * S.nbf + j] = cv; core[j * S.nbf + i] = cv;
      overlap[i * S.nbf + j] = sv; overlap[j * S.nbf + i] = sv;
    }
}
