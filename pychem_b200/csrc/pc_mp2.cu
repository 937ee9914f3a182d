// pc_mp2.cu -- MP2 AO->MO four-index transform on the FP64 tensor cores (DMMA) + energy sums.
//
// Replaces the O(N^6) Python loops of the reference's Methods/mp2.py:37-94:
//   half transform  X[m,b,l,d] = sum_ns C[n,b] G[m,n,l,s] C[s,d]           (mp2.py:43-52)
//   second half     MO[a,b,c,d] = sum_ml C[m,a] X[m,b,l,d] C[l,c]          (mp2.py:56-69)
//   energies        sums over occupied i,j and virtual p,q                  (mp2.py:77-94)
// by four quarter transforms that are plain dense GEMMs (the one true dense contraction of the
// path, hence the one place tensor cores are used), restricted to the index ranges the energy
// formulas read: (i p | j q) with i, j occupied and p, q virtual.
//   T1[i,n,l,s] = sum_m C1[m,i] G[m,n,l,s]          (no1 x N) . (N x N^3)
//   T2[i,p,l,s] = sum_n C1[n,p] T1[i,n,l,s]         batched over i
//   T3[i,p,j,s] = sum_l C2[l,j] T2[i,p,l,s]         batched over (i,p)
//   T4[i,p,j,q] = sum_s T3[i,p,j,s] C2[s,q]         (no1 nv1 no2 x N) . (N x nv2)
// GEMM kernel: mma.sync.aligned.m8n8k4 f64 (DMMA), 64x64x16 CTA tiles staged in shared memory,
// 4 warps x (32x32) register tiles.
#include "../../include/pychem_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <mutex>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace {

thread_local std::string g_mp2_err;
int mp2_fail(const std::string& m) { g_mp2_err = m; return 1; }
#define MP2_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t e_ = (call);                                                        \
    if (e_ != cudaSuccess) return mp2_fail(std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

constexpr int TK = 16;

__device__ __forceinline__ void cp_async8(void* dst, const void* src, bool ok) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  const int bytes = ok ? 8 : 0;                         // 0: the destination is zero-filled
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src, int bytes) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// C[b] (M x N, row-major, ldc) = A[b] (M x K, element (r,k) at A[r*sar + k*sak]) . B[b] (K x N, row-major, ldb)
// 4 warps laid out WARPS_M x WARPS_N, every warp owns (8 WM) x (8 WN) of C as WM x WN DMMA tiles
// (mma.sync m8n8k4 f64 -- the only FP64 tensor shape sm_100a has in hardware: the PTX m16n8k8/k16
// forms lower to the same DMMA.8x8x4, cuobjdump).  Operand tiles are double-buffered in shared
// memory with cp.async (zero-filled at the edges), so the loads of step k+1 overlap the DMMAs of
// step k.  The host picks the tile shape that wastes the least padding on the (skinny) MP2 shapes.
template <int WARPS_M, int WARPS_N, int WM, int WN, bool VEC16>
__global__ void __launch_bounds__(128) dgemm_dmma_kernel(int M, int N, int K, const double* __restrict__ A,
                                                         long long sar, long long sak, long long batch_a,
                                                         const double* __restrict__ B, long long ldb,
                                                         long long batch_b, double* __restrict__ C,
                                                         long long ldc, long long batch_c) {
  constexpr int TM = WARPS_M * 8 * WM, TN = WARPS_N * 8 * WN;
  __shared__ __align__(16) double As[2][TM][TK + 1];
  __shared__ __align__(16) double Bs[2][TK][TN + 8];
  const int bz = blockIdx.z;
  A += (size_t)bz * batch_a;
  B += (size_t)bz * batch_b;
  C += (size_t)bz * batch_c;
  const int row0 = blockIdx.y * TM, col0 = blockIdx.x * TN;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wr = (warp / WARPS_N) * 8 * WM, wc = (warp % WARPS_N) * 8 * WN;   // warp tile origin
  const int g = lane >> 2, t4 = lane & 3;
  double acc[WM][WN][2];
#pragma unroll
  for (int i = 0; i < WM; ++i)
#pragma unroll
    for (int j = 0; j < WN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  auto stage = [&](int buf, int k0) {
    for (int e = tid; e < TM * TK; e += 128) {
      const int r = e / TK, k = e % TK;
      const int gr = row0 + r, gk = k0 + k;
      const bool ok = gr < M && gk < K;
      cp_async8(&As[buf][r][k], ok ? A + (size_t)gr * sar + (size_t)gk * sak : A, ok);
    }
    if (VEC16) {
      for (int e = tid; e < TK * (TN / 2); e += 128) {
        const int k = e / (TN / 2), c = (e % (TN / 2)) * 2;
        const int gk = k0 + k, gc = col0 + c;
        const int nbytes = (gk < K && gc < N) ? (gc + 1 < N ? 16 : 8) : 0;
        cp_async16(&Bs[buf][k][c], nbytes ? B + (size_t)gk * ldb + gc : B, nbytes);
      }
    } else {
      for (int e = tid; e < TK * TN; e += 128) {
        const int k = e / TN, c = e % TN;
        const int gk = k0 + k, gc = col0 + c;
        const bool ok = gk < K && gc < N;
        cp_async8(&Bs[buf][k][c], ok ? B + (size_t)gk * ldb + gc : B, ok);
      }
    }
    cp_async_commit();
  };

  const int nk = (K + TK - 1) / TK;
  stage(0, 0);
  for (int ks = 0; ks < nk; ++ks) {
    const int buf = ks & 1;
    if (ks + 1 < nk) {
      stage(buf ^ 1, (ks + 1) * TK);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; kk += 4) {
      double a[WM], b[WN];
#pragma unroll
      for (int i = 0; i < WM; ++i) a[i] = As[buf][wr + i * 8 + g][kk + t4];
#pragma unroll
      for (int j = 0; j < WN; ++j) b[j] = Bs[buf][kk + t4][wc + j * 8 + g];
#pragma unroll
      for (int i = 0; i < WM; ++i)
#pragma unroll
        for (int j = 0; j < WN; ++j)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                       : "+d"(acc[i][j][0]), "+d"(acc[i][j][1])
                       : "d"(a[i]), "d"(b[j]));
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < WM; ++i)
#pragma unroll
    for (int j = 0; j < WN; ++j) {
      const int r = row0 + wr + i * 8 + g;
      const int c = col0 + wc + j * 8 + t4 * 2;
      if (r < M) {
        if (c < N) C[(size_t)r * ldc + c] = acc[i][j][0];
        if (c + 1 < N) C[(size_t)r * ldc + c + 1] = acc[i][j][1];
      }
    }
}

struct TileCfg { int tm, tn; };
constexpr TileCfg kCfg[5] = {{64, 64}, {24, 128}, {80, 64}, {40, 128}, {24, 96}};

template <int WARPS_M, int WARPS_N, int WM, int WN>
int launch_cfg(cudaStream_t st, bool vec, dim3 grid, int M, int N, int K, const double* A, long long sar, long long sak,
               long long batch_a, const double* B, long long ldb, long long batch_b, double* C, long long ldc,
               long long batch_c) {
  if (vec) dgemm_dmma_kernel<WARPS_M, WARPS_N, WM, WN, true><<<grid, 128, 0, st>>>(M, N, K, A, sar, sak, batch_a, B, ldb, batch_b, C, ldc, batch_c);
  else dgemm_dmma_kernel<WARPS_M, WARPS_N, WM, WN, false><<<grid, 128, 0, st>>>(M, N, K, A, sar, sak, batch_a, B, ldb, batch_b, C, ldc, batch_c);
  MP2_CUDA(cudaGetLastError());
  return 0;
}

int gemm(cudaStream_t st, int M, int N, int K, const double* A, long long sar, long long sak,
         long long batch_a, const double* B, long long ldb, long long batch_b, double* C, long long ldc,
         long long batch_c, int batches) {
  if (M <= 0 || N <= 0 || K <= 0 || batches <= 0) return 0;
  // tile shape with the least padded work
  int best = 0;
  double best_w = 1e300;
  for (int c = 0; c < 5; ++c) {
    const double w = (double)((M + kCfg[c].tm - 1) / kCfg[c].tm) * kCfg[c].tm * (double)((N + kCfg[c].tn - 1) / kCfg[c].tn) * kCfg[c].tn;
    if (w < best_w * 0.999) { best_w = w; best = c; }
  }
  // 16-byte cp.async for B needs 16-byte aligned rows in every batch
  const bool vec = ((uintptr_t)B % 16 == 0) && (ldb % 2 == 0) && (batch_b % 2 == 0);
  for (int b0 = 0; b0 < batches; b0 += 65535) {
    const int nb = std::min(65535, batches - b0);
    dim3 grid((N + kCfg[best].tn - 1) / kCfg[best].tn, (M + kCfg[best].tm - 1) / kCfg[best].tm, nb);
    const double* Ab = A + (size_t)b0 * batch_a;
    const double* Bb = B + (size_t)b0 * batch_b;
    double* Cb = C + (size_t)b0 * batch_c;
    int rc = 0;
    switch (best) {
      case 0: rc = launch_cfg<2, 2, 4, 4>(st, vec, grid, M, N, K, Ab, sar, sak, batch_a, Bb, ldb, batch_b, Cb, ldc, batch_c); break;
      case 1: rc = launch_cfg<1, 4, 3, 4>(st, vec, grid, M, N, K, Ab, sar, sak, batch_a, Bb, ldb, batch_b, Cb, ldc, batch_c); break;
      case 2: rc = launch_cfg<2, 2, 5, 4>(st, vec, grid, M, N, K, Ab, sar, sak, batch_a, Bb, ldb, batch_b, Cb, ldc, batch_c); break;
      case 3: rc = launch_cfg<1, 4, 5, 4>(st, vec, grid, M, N, K, Ab, sar, sak, batch_a, Bb, ldb, batch_b, Cb, ldc, batch_c); break;
      default: rc = launch_cfg<1, 4, 3, 3>(st, vec, grid, M, N, K, Ab, sar, sak, batch_a, Bb, ldb, batch_b, Cb, ldc, batch_c); break;
    }
    if (rc) return rc;
  }
  return 0;
}

// same-spin sum (mp2.py:78-82): i < no, j <= i, p >= no, q in [no, p]:
//   (T[i,p,j,q] - T[i,q,j,p])^2 / (E[i] + E[j] - E[p] - E[q]);  T is [no][nv][no][nv]
__global__ void mp2_same_spin_kernel(int no, int nv, const double* __restrict__ T, const double* __restrict__ E,
                                     double* out) {
  const size_t total = (size_t)no * nv * no * nv;
  double s = 0.0;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int q = idx % nv, j = (idx / nv) % no, p = (idx / ((size_t)nv * no)) % nv, i = idx / ((size_t)nv * no * nv);
    if (j <= i && q <= p) {
      const double d = T[idx] - T[(((size_t)i * nv + q) * no + j) * nv + p];
      s += d * d / (E[i] + E[j] - E[no + p] - E[no + q]);
    }
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}

// opposite-spin sum (mp2.py:90-94): T is [noa][nva][nob][nvb]
__global__ void mp2_opp_spin_kernel(int noa, int nva, int nob, int nvb, const double* __restrict__ T,
                                    const double* __restrict__ Ea, const double* __restrict__ Eb, double* out) {
  const size_t total = (size_t)noa * nva * nob * nvb;
  double s = 0.0;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int q = idx % nvb, j = (idx / nvb) % nob, p = (idx / ((size_t)nvb * nob)) % nva,
              i = idx / ((size_t)nvb * nob * nva);
    const double v = T[idx];
    s += v * v / (Ea[i] + Eb[j] - Ea[noa + p] - Eb[nob + q]);
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}

// Scratch of the transform, kept between calls (grow-only, per process): the four intermediates of
// benzene are 150 MB, allocating and freeing them took longer than the GEMMs (round 1: 17 ms per
// MP2 energy against ~3 ms of kernels).
struct Pool {
  int device = -1;
  cudaStream_t st = nullptr;
  double* p[9] = {nullptr};
  size_t cap[9] = {0};
  cudaError_t get(int k, size_t n, double** out) {
    if (cap[k] < n) {
      if (p[k]) cudaFree(p[k]);
      p[k] = nullptr; cap[k] = 0;
      cudaError_t e = cudaMalloc((void**)&p[k], std::max<size_t>(n, 1) * sizeof(double));
      if (e != cudaSuccess) return e;
      cap[k] = n;
    }
    *out = p[k];
    return cudaSuccess;
  }
  void release() {
    for (int k = 0; k < 9; ++k) { if (p[k]) cudaFree(p[k]); p[k] = nullptr; cap[k] = 0; }
    if (st) cudaStreamDestroy(st);
    st = nullptr;
  }
};
Pool g_pool;
std::mutex g_pool_mu;        // pc_mp2_energy / pc_mp2_release may be called from several host threads

struct Buf { double* p = nullptr; };

bool on_host(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return true; }
  return at.type == cudaMemoryTypeUnregistered || at.type == cudaMemoryTypeHost;
}

// (i p | j q) for i in occ1, p in virt1 (coefficients C1), j in occ2, q in virt2 (C2)
int transform(cudaStream_t st, int N, const double* G, const double* C1, int no1, const double* C2, int no2,
              Buf& T1, Buf& T2, Buf& T3, double* T4) {
  const int nv1 = N - no1, nv2 = N - no2;
  const long long N2 = (long long)N * N, N3 = N2 * N;
  // T1[i,(nls)] = sum_m C1[m,i] G[m,(nls)]
  if (gemm(st, no1, (int)std::min<long long>(N3, 2147483647LL), N, C1, 1, N, 0, G, N3, 0, T1.p, N3, 0, 1)) return 1;
  // T2[i][p,(ls)] = sum_n C1[n,no1+p] T1[i][n,(ls)]
  if (gemm(st, nv1, (int)N2, N, C1 + no1, 1, N, 0, T1.p, N2, N3, T2.p, N2, (long long)nv1 * N2, no1)) return 1;
  // T3[(i,p)][j,s] = sum_l C2[l,j] T2[(i,p)][l,s]
  if (gemm(st, no2, N, N, C2, 1, N, 0, T2.p, N, N2, T3.p, N, (long long)no2 * N, no1 * nv1)) return 1;
  // T4[(i,p,j),q] = sum_s T3[(i,p,j),s] C2[s,no2+q]
  if (gemm(st, no1 * nv1 * no2, nv2, N, T3.p, N, 1, 0, C2 + no2, N, 0, T4, nv2, 0, 1)) return 1;
  return 0;
}

}  // namespace

extern "C" {

const char* pc_mp2_last_error(void) { return g_mp2_err.c_str(); }

int pc_mp2_energy(int device, int N, const double* G_dev, const double* Ca, const double* Cb,
                  const double* Ea, const double* Eb, int na, int nb, int same_spin, double* Eaa,
                  double* Eab, double* Ebb) {
  if (!G_dev || !Ca || !Cb || !Ea || !Eb || !Eaa || !Eab || !Ebb) return mp2_fail("pc_mp2_energy: null");
  if (N <= 0 || na < 0 || nb < 0 || na > N || nb > N) return mp2_fail("pc_mp2_energy: bad sizes");
  if ((long long)N * N * N > 2147483647LL) return mp2_fail("pc_mp2_energy: N too large for this build");
  std::lock_guard<std::mutex> lock(g_pool_mu);
  MP2_CUDA(cudaSetDevice(device));
  if (g_pool.device != device) { g_pool.release(); g_pool.device = device; }
  if (!g_pool.st) MP2_CUDA(cudaStreamCreateWithFlags(&g_pool.st, cudaStreamNonBlocking));
  cudaStream_t st = g_pool.st;
  const size_t NN = (size_t)N * N;
  // restricted orbitals (RHF: the same coefficient and energy arrays for both spins): one
  // transform serves all three sums
  const bool restricted = na == nb && ((Ca == Cb && Ea == Eb) ||
                                       (on_host(Ca) && on_host(Cb) && on_host(Ea) && on_host(Eb) &&
                                        memcmp(Ca, Cb, NN * sizeof(double)) == 0 && memcmp(Ea, Eb, N * sizeof(double)) == 0));
  Buf dCa, dCb, dEa, dEb, T1, T2, T3, T4, out;
  MP2_CUDA(g_pool.get(0, NN, &dCa.p)); MP2_CUDA(g_pool.get(1, NN, &dCb.p));
  MP2_CUDA(g_pool.get(2, N, &dEa.p)); MP2_CUDA(g_pool.get(3, N, &dEb.p));
  MP2_CUDA(g_pool.get(4, 3, &out.p));
  MP2_CUDA(cudaMemcpyAsync(dCa.p, Ca, NN * sizeof(double), cudaMemcpyDefault, st));
  MP2_CUDA(cudaMemcpyAsync(dCb.p, Cb, NN * sizeof(double), cudaMemcpyDefault, st));
  MP2_CUDA(cudaMemcpyAsync(dEa.p, Ea, N * sizeof(double), cudaMemcpyDefault, st));
  MP2_CUDA(cudaMemcpyAsync(dEb.p, Eb, N * sizeof(double), cudaMemcpyDefault, st));
  MP2_CUDA(cudaMemsetAsync(out.p, 0, 3 * sizeof(double), st));
  const int nom = std::max(na, nb), nvm = N - std::min(na, nb);
  MP2_CUDA(g_pool.get(5, (size_t)nom * NN * N, &T1.p));
  MP2_CUDA(g_pool.get(6, (size_t)nom * nvm * NN, &T2.p));
  MP2_CUDA(g_pool.get(7, (size_t)nom * nvm * nom * N, &T3.p));
  MP2_CUDA(g_pool.get(8, (size_t)nom * nvm * nom * nvm, &T4.p));
  const int blocks = 148 * 8;
  if (restricted) {
    if (na > 0 && N - na > 0) {
      if (transform(st, N, G_dev, dCa.p, na, dCa.p, na, T1, T2, T3, T4.p)) return 1;
      if (same_spin) {
        mp2_same_spin_kernel<<<blocks, 256, 0, st>>>(na, N - na, T4.p, dEa.p, out.p);
        MP2_CUDA(cudaGetLastError());
      }
      mp2_opp_spin_kernel<<<blocks, 256, 0, st>>>(na, N - na, na, N - na, T4.p, dEa.p, dEa.p, out.p + 1);
      MP2_CUDA(cudaGetLastError());
    }
  } else {
    if (same_spin && na > 0 && N - na > 0) {
      if (transform(st, N, G_dev, dCa.p, na, dCa.p, na, T1, T2, T3, T4.p)) return 1;
      mp2_same_spin_kernel<<<blocks, 256, 0, st>>>(na, N - na, T4.p, dEa.p, out.p);
      MP2_CUDA(cudaGetLastError());
    }
    if (same_spin && nb > 0 && N - nb > 0) {
      if (transform(st, N, G_dev, dCb.p, nb, dCb.p, nb, T1, T2, T3, T4.p)) return 1;
      mp2_same_spin_kernel<<<blocks, 256, 0, st>>>(nb, N - nb, T4.p, dEb.p, out.p + 2);
      MP2_CUDA(cudaGetLastError());
    }
    if (na > 0 && nb > 0 && N - na > 0 && N - nb > 0) {
      if (transform(st, N, G_dev, dCa.p, na, dCb.p, nb, T1, T2, T3, T4.p)) return 1;
      mp2_opp_spin_kernel<<<blocks, 256, 0, st>>>(na, N - na, nb, N - nb, T4.p, dEa.p, dEb.p, out.p + 1);
      MP2_CUDA(cudaGetLastError());
    }
  }
  double res[3];
  MP2_CUDA(cudaMemcpyAsync(res, out.p, 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
  MP2_CUDA(cudaStreamSynchronize(st));
  *Eaa = res[0]; *Eab = res[1]; *Ebb = restricted ? res[0] : res[2];
  return 0;
}

// frees the transform scratch kept between pc_mp2_energy calls
int pc_mp2_release(void) {
  std::lock_guard<std::mutex> lock(g_pool_mu);
  g_pool.release();
  g_pool.device = -1;
  return 0;
}

// plain C = A.B (row-major, all device) through the DMMA kernel, for tests and ncu
int pc_dgemm_dmma(int device, int M, int N, int K, const double* A, const double* B, double* C) {
  MP2_CUDA(cudaSetDevice(device));
  if (gemm(nullptr, M, N, K, A, K, 1, 0, B, N, 0, C, N, 0, 1)) return 1;
  MP2_CUDA(cudaDeviceSynchronize());
  return 0;
}

}  // extern "C"
