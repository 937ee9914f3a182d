// pc_generic.cuh -- ERI kernel for ANY angular-momentum class up to (ff|ff) (sm_100a, FP64).
//
// The generated class kernels (csrc/gen/eri_*.cu) cover s, p and d shells, the shapes of the
// benchmark configurations.  The reference's recursion machinery is generic in l
// (Methods/integrals.py:73-191 SetRR2, Data/transform_basis.py:3-67) and its basis library ships
// f functions (cc-pVTZ, 6-311G(3df,3pd), pc-n ...), so shell quartets with an f shell take this
// kernel instead: the same recursion family, driven by loops over the Cartesian components
// instead of generated straight-line code.
//   fundamentals  Methods/c_ints/two_electron_fundamentals.c:41-89  -> pc_fundamentals<L> (shared)
//   VRR           Methods/c_ints/two_electron_vrr.c:92-108            -> pcg_vrr  (ket on the s bra
//                 first, then the bra for every ket shell: integrals.py:514-517)
//   contraction   Methods/c_ints/two_electron_contract.c:45            -> accumulated into S
//   HRR           Methods/c_ints/two_electron_hrr.c:82                -> pcg_hrr  (ket, then bra:
//                 integrals.py:531-536)
//   normalise + cart->spherical  Methods/integrals.py:541-547, Data/transform_basis.py:3-30 -> pcg_c2s_side
//
// One thread owns one shell quartet, as in the generated kernels, and the plan / segment decode,
// the pair tables and the accumulators are shared with them.  What does not fit into registers
// (the VRR table of (ff|ff) has ~25 000 entries) lives in a per-thread scratch column in global
// memory, interleaved over the threads of the launch so that a warp touches consecutive words.
// The launch is persistent: a fixed number of threads (bounded by the scratch budget) walk the
// warps of the fused launch with a stride.  Digestion uses plain loops with one atomic per
// target element and quartet.  This is the completeness path, not the tuned one.
#pragma once
#include "pc_common.cuh"
#include "pc_generic_class.h"

#ifdef PC_HOST_EMU
#define PCG_NOINLINE
#else
#define PCG_NOINLINE __noinline__
#endif

// k-th component of shell l in that order
__device__ __forceinline__ void pcg_comp(int l, int k, int& lx, int& ly, int& lz) {
  int s = 0;                               // s = ly + lz: components with sum s start at s(s+1)/2
  while ((s + 1) * (s + 2) / 2 <= k) ++s;
  lz = k - s * (s + 1) / 2;
  ly = s - lz;
  lx = l - s;
}

// (Gamma(n + 1/2))^-1/2, n = 0..3  (Util/structures.py:850-856)
__device__ __forceinline__ double pcg_gnorm(int n) {
  return n == 0 ? 0.7511255444649425 : (n == 1 ? 1.0622519320271968 : (n == 2 ? 0.8673250705840775 : 0.5485445389623983));
}
// angular normalisation of one component relative to the uniform constant of its shell that the
// pair tables already carry (upload_kind: pi^-3/4 {1, sqrt2, 2, 2 sqrt2} for s, p, d, f)
__device__ __forceinline__ double pcg_norm_ratio(int l, int lx, int ly, int lz) {
  const double inv_uniform = l == 0 ? 2.359730492414697 : (l == 1 ? 1.668581432959103 : (l == 2 ? 1.1798652462073485 : 0.8342907164795516));
  return pcg_gnorm(lx) * pcg_gnorm(ly) * pcg_gnorm(lz) * inv_uniform;
}

// cart -> real spherical coefficient (Data/transform_basis.py:3-30); identity for s, p and for
// Cartesian d
__device__ __forceinline__ double pcg_c2s(int l, bool cart, int m, int c) {
  if (l < 2 || cart) return m == c ? 1.0 : 0.0;
  if (l == 2) {
    const double T[5][6] = {{0.86602540378443865, 0, 0, -0.86602540378443865, 0, 0},
                            {0, 1, 0, 0, 0, 0}, {0, 0, 1, 0, 0, 0}, {0, 0, 0, 0, 1, 0},
                            {-0.5, 0, 0, -0.5, 0, 1}};
    return T[m][c];
  }
  const double T[7][10] = {
      {0.7905694150420949, 0, 0, -1.0606601717798212, 0, 0, 0, 0, 0, 0},
      {0, 1.0606601717798212, 0, 0, 0, 0, -0.7905694150420949, 0, 0, 0},
      {0, 0, 0.8660254037844386, 0, 0, 0, 0, -0.8660254037844386, 0, 0},
      {0, 0, 0, 0, 1, 0, 0, 0, 0, 0},
      {-0.6123724356957945, 0, 0, -0.27386127875258304, 0, 1.0954451150103321, 0, 0, 0, 0},
      {0, -0.27386127875258304, 0, 0, 0, 0, -0.6123724356957945, 0, 1.0954451150103321, 0},
      {0, 0, -0.6708203932499369, 0, 0, 0, 0, -0.6708203932499369, 0, 1}};
  return T[m][c];
}

// fundamentals for a run-time L: dispatch to the shared templates
template <int L>
__device__ __forceinline__ void pcg_fund_L(bool scat, double sP, double UP, double sQ, double UQ, double R2,
                                           const double* __restrict__ boys, double S, double* F) {
  double f[L + 1];
  if (scat) pc_fundamentals_scatter<L>(sP, UP, sQ, UQ, R2, S, f);
  else pc_fundamentals<L>(sP, UP, sQ, UQ, R2, boys, f);
#pragma unroll
  for (int m = 0; m <= L; ++m) F[m] = f[m];
}

__device__ PCG_NOINLINE void pcg_fund(int L, bool scat, double sP, double UP, double sQ, double UQ, double R2,
                                      const double* __restrict__ boys, double S, double* F) {
  switch (L) {
    case 0: pcg_fund_L<0>(scat, sP, UP, sQ, UQ, R2, boys, S, F); break;
    case 1: pcg_fund_L<1>(scat, sP, UP, sQ, UQ, R2, boys, S, F); break;
    case 2: pcg_fund_L<2>(scat, sP, UP, sQ, UQ, R2, boys, S, F); break;
    case 3: pcg_fund_L<3>(scat, sP, UP, sQ, UQ, R2, boys, S, F); break;
    case 4: pcg_fund_L<4>(scat, sP, UP, sQ, UQ, R2, boys, S, F); break;
    case 5: pcg_fund_L<5>(scat, sP, UP, sQ, UQ, R2, boys, S, F); break;
    case 6: pcg_fund_L<6>(scat, sP, UP, sQ, UQ, R2, boys, S, F); break;
    case 7: pcg_fund_L<7>(scat, sP, UP, sQ, UQ, R2, boys, S, F); break;
    case 8: pcg_fund_L<8>(scat, sP, UP, sQ, UQ, R2, boys, S, F); break;
    case 9: pcg_fund_L<9>(scat, sP, UP, sQ, UQ, R2, boys, S, F); break;
    case 10: pcg_fund_L<10>(scat, sP, UP, sQ, UQ, R2, boys, S, F); break;
    case 11: pcg_fund_L<11>(scat, sP, UP, sQ, UQ, R2, boys, S, F); break;
    default: pcg_fund_L<12>(scat, sP, UP, sQ, UQ, R2, boys, S, F); break;
  }
}

// ---------------------------------------------------------------------------------------------
// VRR of one primitive quartet into the thread's table (region A).  Block (la, lc) holds
// [ia][ic][m], m = 0 .. L - la - lc, Gill-scaled (the (2 theta^2)^m factors live in F).
//   ket on the s bra:  [0, c+1_i]^(m) = -CD_i kz_Q [0,c]^(m) + R_i eta [0,c]^(m+1)
//                                       + c_i eta ([0,c-1_i]^(m) - eta [0,c-1_i]^(m+1))
//   bra:               [a+1_i, c]^(m) = -AB_i kz_P [a,c]^(m) - R_i zeta [a,c]^(m+1)
//                                       + a_i zeta ([a-1_i,c]^(m) - zeta [a-1_i,c]^(m+1))
//                                       + c_i zeta eta [a,c-1_i]^(m+1)
// (two_electron_vrr.c:92-108 with the signs of integrals.py:561-569,586; R = P - Q; the
// direction i is the first non-zero power of the target, two_electron_vrr.c:33-48)
// ---------------------------------------------------------------------------------------------
__device__ PCG_NOINLINE void pcg_vrr(const PcGenClass& C, double* __restrict__ V, const size_t st, const double* F,
                                     const double zeta, const double eta, const double* PX, const double* QX,
                                     const double* R) {
  const int La = C.lx1 + C.ly1, Lc = C.lx2 + C.ly2, L = C.L;
  for (int m = 0; m <= L; ++m) V[(size_t)(C.offV[0][0] + m) * st] = F[m];
  for (int lc = 1; lc <= Lc; ++lc) {
    const int nm = L + 1 - lc;                      // orders of the target block
    const int o_t = C.offV[0][lc], o_1 = C.offV[0][lc - 1], o_2 = lc > 1 ? C.offV[0][lc - 2] : 0;
    for (int ic = 0; ic < pcg_ncart(lc); ++ic) {
      int c[3];
      pcg_comp(lc, ic, c[0], c[1], c[2]);
      const int dir = c[0] ? 0 : (c[1] ? 1 : 2);
      c[dir] -= 1;
      const int i1 = pcg_cidx(c[1], c[2]);
      const int nval = c[dir];
      int i2 = 0;
      if (nval > 0) { c[dir] -= 1; i2 = pcg_cidx(c[1], c[2]); }
      const double c0 = QX[dir], c1 = R[dir] * eta, c2 = nval * eta;
      for (int m = 0; m < nm; ++m) {
        double v = c0 * V[(size_t)(o_1 + i1 * (nm + 1) + m) * st] + c1 * V[(size_t)(o_1 + i1 * (nm + 1) + m + 1) * st];
        if (nval > 0)
          v += c2 * (V[(size_t)(o_2 + i2 * (nm + 2) + m) * st] - eta * V[(size_t)(o_2 + i2 * (nm + 2) + m + 1) * st]);
        V[(size_t)(o_t + ic * nm + m) * st] = v;
      }
    }
  }
  const double ze = zeta * eta;
  for (int la = 1; la <= La; ++la)
    for (int ia = 0; ia < pcg_ncart(la); ++ia) {
      int a[3];
      pcg_comp(la, ia, a[0], a[1], a[2]);
      const int dir = a[0] ? 0 : (a[1] ? 1 : 2);
      a[dir] -= 1;
      const int a1 = pcg_cidx(a[1], a[2]);
      const int aval = a[dir];
      int a2 = 0;
      if (aval > 0) { a[dir] -= 1; a2 = pcg_cidx(a[1], a[2]); }
      const double c0 = PX[dir], c1 = -R[dir] * zeta, c2 = aval * zeta;
      for (int lc = 0; lc <= Lc; ++lc) {
        const int nm = L + 1 - la - lc, nc = pcg_ncart(lc);
        const int o_t = C.offV[la][lc] + ia * nc * nm;
        const int o_1 = C.offV[la - 1][lc] + a1 * nc * (nm + 1);
        const int o_2 = la > 1 ? C.offV[la - 2][lc] + a2 * nc * (nm + 2) : 0;
        const int ncm = lc > 0 ? pcg_ncart(lc - 1) : 0;
        const int o_c = lc > 0 ? C.offV[la - 1][lc - 1] + a1 * ncm * (nm + 2) : 0;
        for (int ic = 0; ic < nc; ++ic) {
          int c[3];
          pcg_comp(lc, ic, c[0], c[1], c[2]);
          const int cval = c[dir];
          int ic1 = 0;
          if (cval > 0) { c[dir] -= 1; ic1 = pcg_cidx(c[1], c[2]); }
          const double c3 = cval * ze;
          for (int m = 0; m < nm; ++m) {
            double v = c0 * V[(size_t)(o_1 + ic * (nm + 1) + m) * st] + c1 * V[(size_t)(o_1 + ic * (nm + 1) + m + 1) * st];
            if (aval > 0)
              v += c2 * (V[(size_t)(o_2 + ic * (nm + 2) + m) * st] - zeta * V[(size_t)(o_2 + ic * (nm + 2) + m + 1) * st]);
            if (cval > 0) v += c3 * V[(size_t)(o_c + ic1 * (nm + 2) + m + 1) * st];
            V[(size_t)(o_t + ic * nm + m) * st] = v;
          }
        }
      }
    }
}

// ---------------------------------------------------------------------------------------------
// HRR of one side on one vector (two_electron_hrr.c:82):
//   (x, y+1_i| = (x+1_i, y| + (X - Y)_i (x, y|
// in : buf0 = (e, 0| for e = lx .. lx+ly, shells concatenated;  out: (lx, ly| as [ix][iy] in
// whichever buffer the last level was written to (returned).  Both buffers hold PCG_HRR_BUF words.
// ---------------------------------------------------------------------------------------------
__device__ PCG_NOINLINE double* pcg_hrr(int lx, int ly, const double* XY, double* buf0, double* buf1) {
  double* cur = buf0;
  double* nxt = buf1;
  const int Lt = lx + ly;
  for (int k = 1; k <= ly; ++k) {
    const int nk = pcg_ncart(k), nk1 = pcg_ncart(k - 1);
    int o_t = 0;                       // start of block e in the target level
    int o_e = 0;                       // start of block e in the current level
    for (int e = lx; e <= Lt - k; ++e) {
      const int ne = pcg_ncart(e);
      const int o_e1 = o_e + ne * nk1;            // block e + 1 of the current level
      for (int ix = 0; ix < ne; ++ix) {
        int x[3];
        pcg_comp(e, ix, x[0], x[1], x[2]);
        for (int iy = 0; iy < nk; ++iy) {
          int y[3];
          pcg_comp(k, iy, y[0], y[1], y[2]);
          const int dir = y[0] ? 0 : (y[1] ? 1 : 2);
          y[dir] -= 1;
          const int iy0 = pcg_cidx(y[1], y[2]);
          int xp[3] = {x[0], x[1], x[2]};
          xp[dir] += 1;
          const int ix1 = pcg_cidx(xp[1], xp[2]);
          nxt[o_t + ix * nk + iy] = cur[o_e1 + ix1 * nk1 + iy0] + XY[dir] * cur[o_e + ix * nk1 + iy0];
        }
      }
      o_t += ne * nk;
      o_e = o_e1;
    }
    double* t = cur; cur = nxt; nxt = t;
  }
  return cur;
}

// normalise and transform one pair side of a vector: in[cx][cy] (Cartesian, nxc x nyc) ->
// out[mx][my] (nx x ny functions); tmp holds nxc * ny words
__device__ PCG_NOINLINE void pcg_c2s_side(int lx, int ly, bool cart_d, int nx, int ny, const double* in, double* tmp,
                                          double* out) {
  const int nxc = pcg_ncart(lx), nyc = pcg_ncart(ly);
  const bool cx = (lx == 2 && cart_d), cy = (ly == 2 && cart_d);
  double ny_r[PCG_NCART_MAX];
  for (int iy = 0; iy < nyc; ++iy) {
    int a, b, c;
    pcg_comp(ly, iy, a, b, c);
    ny_r[iy] = pcg_norm_ratio(ly, a, b, c);
  }
  for (int ix = 0; ix < nxc; ++ix) {
    int a, b, c;
    pcg_comp(lx, ix, a, b, c);
    const double rx = pcg_norm_ratio(lx, a, b, c);
    for (int my = 0; my < ny; ++my) {
      double s = 0.0;
      for (int iy = 0; iy < nyc; ++iy) {
        const double t = pcg_c2s(ly, cy, my, iy);
        if (t != 0.0) s = fma(t * ny_r[iy], in[ix * nyc + iy], s);
      }
      tmp[ix * ny + my] = rx * s;
    }
  }
  for (int mx = 0; mx < nx; ++mx)
    for (int my = 0; my < ny; ++my) {
      double s = 0.0;
      for (int ix = 0; ix < nxc; ++ix) {
        const double t = pcg_c2s(lx, cx, mx, ix);
        if (t != 0.0) s = fma(t, tmp[ix * ny + my], s);
      }
      out[mx * ny + my] = s;
    }
}

// ---------------------------------------------------------------------------------------------
// digestion with run-time block sizes: one image = one target block,
//   T[(fr + r) N + fc + c] += fac sum_{u,v} d(u, v) g[r sr + c sc + u su + v sv]
// with d(u,v) = D[(fu+u) N + fv+v]  (dmode 0), twice that (1) or D[u,v] + D[v,u] (2)
// ---------------------------------------------------------------------------------------------
struct PcgAxis { int f, n, s; };   // first function, functions, stride inside the block

__device__ PCG_NOINLINE void pcg_image(double* __restrict__ T, const double* __restrict__ D, const size_t N,
                                       const double fac, const PcgAxis r, const PcgAxis c, const PcgAxis u,
                                       const PcgAxis v, const int dmode, const double* __restrict__ g, const size_t st) {
  for (int ir = 0; ir < r.n; ++ir)
    for (int ic = 0; ic < c.n; ++ic) {
      double s = 0.0;
      for (int iu = 0; iu < u.n; ++iu)
        for (int iv = 0; iv < v.n; ++iv) {
          double d = D[(size_t)(u.f + iu) * N + v.f + iv];
          if (dmode == 1) d += d;
          else if (dmode == 2) d += D[(size_t)(v.f + iv) * N + u.f + iu];
          s = fma(d, g[(size_t)(ir * r.s + ic * c.s + iu * u.s + iv * v.s) * st], s);
        }
      atomicAdd(&T[(size_t)(r.f + ir) * N + c.f + ic], fac * s);
    }
}

// J/K digestion of one block (the images of pc_digest_jk, hartree_fock.py:345-347).  Accumulators
// are the same "half" matrices the generated kernels add into (finalised by jk_finalize_kernel).
template <bool GEN, int NSPIN>
__device__ __forceinline__ void pcg_digest(const PcJkView& V, const PcgAxis a, const PcgAxis b, const PcgAxis c,
                                           const PcgAxis d, const double fac, const double* g, const size_t st) {
  const size_t N = V.nbf;
  const int jm = GEN ? 2 : 1;
  pcg_image(V.Jacc, V.Dj, N, fac, a, b, c, d, jm, g, st);       // J[a,b] += (Dj[c,d] + Dj[d,c]) g
  pcg_image(V.Jacc, V.Dj, N, fac, c, d, a, b, jm, g, st);       // J[c,d] += (Dj[a,b] + Dj[b,a]) g
  for (int spin = 0; spin < NSPIN; ++spin) {
    const double* D = spin ? V.Db : V.Da;
    double* K = spin ? V.Kbacc : V.Kaacc;
    pcg_image(K, D, N, fac, a, d, c, b, 0, g, st);              // K[a,d] += D[c,b] g
    pcg_image(K, D, N, fac, b, d, c, a, 0, g, st);              // K[b,d] += D[c,a] g
    if (GEN) {
      pcg_image(K, D, N, fac, a, c, d, b, 0, g, st);            // K[a,c] += D[d,b] g
      pcg_image(K, D, N, fac, b, c, d, a, 0, g, st);            // K[b,c] += D[d,a] g
      pcg_image(K, D, N, fac, c, b, a, d, 0, g, st);            // K[c,b] += D[a,d] g
      pcg_image(K, D, N, fac, c, a, b, d, 0, g, st);            // K[c,a] += D[b,d] g
      pcg_image(K, D, N, fac, d, b, a, c, 0, g, st);            // K[d,b] += D[a,c] g
      pcg_image(K, D, N, fac, d, a, b, c, 0, g, st);            // K[d,a] += D[b,c] g
    } else {
      pcg_image(K, D, N, fac, a, c, b, d, 0, g, st);            // symmetric D: D[d,b] read as D[b,d]
      pcg_image(K, D, N, fac, b, c, a, d, 0, g, st);            //              D[d,a] read as D[a,d]
    }
  }
}

// ---------------------------------------------------------------------------------------------
// one shell quartet.  Scratch of the thread: region A = sA[k * st], region B = sB[k * st].
//   primitive loop   V (A)                 contracted (e0|f0) S (B)
//   ket HRR          S (B)  -> T1 (A + offT1)   [e][(x2, y2) Cartesian]
//   bra HRR          T1     -> G  (A + offG)    [(x1, y1) Cartesian][(x2, y2) Cartesian]
//   ket transform    G      -> H  (B)           [(x1, y1) Cartesian][(x2, y2) functions]
//   bra transform    H      -> out (A)          [(x1, y1) functions][(x2, y2) functions]
// ---------------------------------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ void pcg_quartet(const PcEriArgs& A, const PcGenClass& C, const PcItem& I, const long long t,
                                            const int i, const int j, double* __restrict__ sA,
                                            double* __restrict__ sB, const size_t st) {
  const int nb = I.bra.n, nk = I.ket.n;
  const int La = C.lx1 + C.ly1, Lc = C.lx2 + C.ly2;
  const int ea0 = pcg_ncum(C.lx1 - 1), fc0 = pcg_ncum(C.lx2 - 1);
  const int ne = pcg_ncum(La) - ea0, nf = pcg_ncum(Lc) - fc0;
  const double AB[3] = {__ldg(I.bra.xy + i), __ldg(I.bra.xy + nb + i), __ldg(I.bra.xy + 2 * (size_t)nb + i)};
  const double CD[3] = {__ldg(I.ket.xy + j), __ldg(I.ket.xy + nk + j), __ldg(I.ket.xy + 2 * (size_t)nk + j)};
  const int KB = __ldg(I.bra.keff + i), KK = __ldg(I.ket.keff + j);
  for (int k = 0; k < ne * nf; ++k) sB[(size_t)k * st] = 0.0;
  const double2* __restrict__ bp = reinterpret_cast<const double2*>(I.bra.prim) + i;
  const double2* __restrict__ kp = reinterpret_cast<const double2*>(I.ket.prim) + j;
  for (int ik = 0; ik < KK; ++ik) {
    const double2 q0 = __ldg(kp + (size_t)(3 * ik) * nk), q1 = __ldg(kp + (size_t)(3 * ik + 1) * nk),
                  q2 = __ldg(kp + (size_t)(3 * ik + 2) * nk);
    const double sQ = q0.x, UQ = q0.y, eta = 0.5 * sQ;
    const double QX[3] = {-CD[0] * q2.y, -CD[1] * q2.y, -CD[2] * q2.y};
    for (int ib = 0; ib < KB; ++ib) {
      const double2 p0 = __ldg(bp + (size_t)(3 * ib) * nb), p1 = __ldg(bp + (size_t)(3 * ib + 1) * nb),
                    p2 = __ldg(bp + (size_t)(3 * ib + 2) * nb);
      const double sP = p0.x, UP = p0.y, zeta = 0.5 * sP;
      const double PX[3] = {-AB[0] * p2.y, -AB[1] * p2.y, -AB[2] * p2.y};
      const double R[3] = {p1.x - q1.x, p1.y - q1.y, p2.x - q2.x};
      const double Rsq = R[0] * R[0] + R[1] * R[1] + R[2] * R[2];
      double F[4 * PCG_LMAX + 1];
      pcg_fund(C.L, C.scat != 0, sP, UP, sQ, UQ, Rsq, A.boys, A.scat_S, F);
      pcg_vrr(C, sA, st, F, zeta, eta, PX, QX, R);
      // contraction: the weights are part of the pair prefactors (two_electron_contract.c:45)
      int e = 0;
      for (int la = C.lx1; la <= La; ++la)
        for (int ia = 0; ia < pcg_ncart(la); ++ia, ++e) {
          int f = 0;
          for (int lc = C.lx2; lc <= Lc; ++lc) {
            const int nm = C.L + 1 - la - lc, nc = pcg_ncart(lc);
            const int o = C.offV[la][lc] + ia * nc * nm;
            for (int ic = 0; ic < nc; ++ic, ++f) sB[(size_t)(e * nf + f) * st] += sA[(size_t)(o + ic * nm) * st];
          }
        }
    }
  }
  double b0[PCG_HRR_BUF], b1[PCG_HRR_BUF];
  const int nkc = pcg_ncart(C.lx2) * pcg_ncart(C.ly2), nbc = pcg_ncart(C.lx1) * pcg_ncart(C.ly1);
  // ket HRR, one bra component at a time
  for (int e = 0; e < ne; ++e) {
    for (int f = 0; f < nf; ++f) b0[f] = sB[(size_t)(e * nf + f) * st];
    const double* r = pcg_hrr(C.lx2, C.ly2, CD, b0, b1);
    for (int q = 0; q < nkc; ++q) sA[(size_t)(C.offT1 + e * nkc + q) * st] = r[q];
  }
  // bra HRR, one ket component at a time
  for (int q = 0; q < nkc; ++q) {
    for (int e = 0; e < ne; ++e) b0[e] = sA[(size_t)(C.offT1 + e * nkc + q) * st];
    const double* r = pcg_hrr(C.lx1, C.ly1, AB, b0, b1);
    for (int p = 0; p < nbc; ++p) sA[(size_t)(C.offG + p * nkc + q) * st] = r[p];
  }
  // normalise + cart -> spherical (integrals.py:541-547): ket side of every bra row ...
  const int nsk = C.nx2 * C.ny2, nsb = C.nx1 * C.ny1;
  for (int p = 0; p < nbc; ++p) {
    for (int q = 0; q < nkc; ++q) b0[q] = sA[(size_t)(C.offG + p * nkc + q) * st];
    pcg_c2s_side(C.lx2, C.ly2, C.cart_d != 0, C.nx2, C.ny2, b0, b1, b1 + 70);
    for (int q = 0; q < nsk; ++q) sB[(size_t)(p * nsk + q) * st] = b1[70 + q];
  }
  // ... then the bra side of every ket column
  for (int q = 0; q < nsk; ++q) {
    for (int p = 0; p < nbc; ++p) b0[p] = sB[(size_t)(p * nsk + q) * st];
    pcg_c2s_side(C.lx1, C.ly1, C.cart_d != 0, C.nx1, C.ny1, b0, b1, b1 + 70);
    for (int p = 0; p < nsb; ++p) sA[(size_t)(p * nsk + q) * st] = b1[70 + p];
  }
  // ---- epilogue: the block is sA[k * st], k = ((m ny1 + n) nx2 + l) ny2 + s ----
  const int nsph = nsb * nsk;
  if (MODE == PC_MODE_NULL) {
    double sum = 0.0;
    for (int k = 0; k < nsph; ++k) sum += sA[(size_t)k * st];
    if (sum == 1.2345678e300) A.out[0] = sum;
    return;
  }
  if (MODE == PC_MODE_BLOCKS || MODE == PC_MODE_BLOCKS_SCAT) {
    for (int k = 0; k < nsph; ++k) A.out[(size_t)k * I.t_count + t] = sA[(size_t)k * st];
    return;
  }
  const int fa = __ldg(I.bra.fx + i), fb = __ldg(I.bra.fy + i), fc = __ldg(I.ket.fx + j), fd = __ldg(I.ket.fy + j);
  if (MODE == PC_MODE_TENSOR || MODE == PC_MODE_TENSOR_SCAT) {
    // dense tensor with 8-fold symmetry (hartree_fock.py:314-325)
    const size_t N = A.nbf;
    double* G = A.G;
    int k = 0;
    for (int m = 0; m < C.nx1; ++m)
      for (int n = 0; n < C.ny1; ++n)
        for (int l = 0; l < C.nx2; ++l)
          for (int s = 0; s < C.ny2; ++s, ++k) {
            const double v = sA[(size_t)k * st];
            const size_t a = fa + m, b = fb + n, c = fc + l, d = fd + s;
            G[((a * N + b) * N + c) * N + d] = v; G[((b * N + a) * N + c) * N + d] = v;
            G[((a * N + b) * N + d) * N + c] = v; G[((b * N + a) * N + d) * N + c] = v;
            G[((c * N + d) * N + a) * N + b] = v; G[((c * N + d) * N + b) * N + a] = v;
            G[((d * N + c) * N + a) * N + b] = v; G[((d * N + c) * N + b) * N + a] = v;
          }
    return;
  }
  // J/K digestion; shell-level degeneracy 1/2 per a==b, c==d, (ab)==(cd)
  double fac = 1.0;
  if (fa == fb) fac *= 0.5;
  if (fc == fd) fac *= 0.5;
  if (__ldg(I.bra.pid + i) == __ldg(I.ket.pid + j)) fac *= 0.5;
  const PcgAxis a = {fa, C.nx1, C.ny1 * nsk}, b = {fb, C.ny1, nsk}, c = {fc, C.nx2, C.ny2}, d = {fd, C.ny2, 1};
  if (MODE == PC_MODE_JK_GEN_BATCH) {
    for (int s = 0; s < A.nset; ++s) {
      const size_t off = (size_t)s * (size_t)A.set_stride;
      const PcJkView V = {A.nbf, A.Dj + off, A.Da + off, A.Db + off, A.Jacc + off, A.Kaacc + off, A.Kbacc + off};
      pcg_digest<true, 2>(V, a, b, c, d, fac, sA, st);
    }
  } else {
    const PcJkView V = {A.nbf, A.Dj, A.Da, A.Db, A.Jacc, A.Kaacc, A.Kbacc};
    if (MODE == PC_MODE_JK_GEN) pcg_digest<true, 2>(V, a, b, c, d, fac, sA, st);
    else if (MODE == PC_MODE_JK_UHF) pcg_digest<false, 2>(V, a, b, c, d, fac, sA, st);
    else pcg_digest<false, 1>(V, a, b, c, d, fac, sA, st);
  }
}

template <int MODE>
__global__ void __launch_bounds__(64) eri_generic_kernel(const __grid_constant__ PcEriArgs A,
                                                        const __grid_constant__ PcGenClass C) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= C.nthreads) return;                          // whole warps only: nthreads % 32 == 0
  const int lane = threadIdx.x & 31;
  const size_t st = (size_t)C.nthreads;
  double* __restrict__ sA = C.scratch + tid;
  double* __restrict__ sB = sA + (size_t)C.sizeA * st;
  const int wstep = C.nthreads >> 5;
  for (int gw = tid >> 5; gw < A.nwarps; gw += wstep) {
    const PcItem& I = A.items[pc_find_item(A, gw)];
    const long long t = (long long)(gw - I.warp0) * 32 + lane;
    int i, j, seg_lo, seg_hi;
    if (pc_decode_task(A, I, t, i, j, seg_lo, seg_hi)) pcg_quartet<MODE>(A, C, I, t, i, j, sA, sB, st);
    __syncwarp();                                         // the decode is warp-cooperative
  }
}
