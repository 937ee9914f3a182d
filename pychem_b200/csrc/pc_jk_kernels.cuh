// pc_jk_kernels.cuh -- the kernels of the library that are not ERI class kernels: finalisation of
// the J/K half-accumulators, stored-tensor J/K (single and batched density sets), density
// classification and the FP64 peak micro-benchmark.  Included by pc_api.cu only.
//   J/K from the stored tensor  Methods/hartree_fock.py:345-347 (three einsum passes in the reference)
#pragma once
#include "pc_common.cuh"
#include "pc_async.cuh"

namespace {

// ------------------------------------------------------------------------------------------
// small kernels
// ------------------------------------------------------------------------------------------
// J = Jacc + Jacc^T ; X = -(Kacc + Kacc^T) (symmetric densities) or -Kacc (general)
__global__ void jk_finalize_kernel(int N, int general, int nspin, const double* __restrict__ acc,
                                   double* __restrict__ J, double* __restrict__ Xa,
                                   double* __restrict__ Xb) {
  const size_t nn = (size_t)N * N;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < nn;
       idx += (size_t)gridDim.x * blockDim.x) {
    const size_t r = idx / N, c = idx % N, tr = c * N + r;
    J[idx] = acc[idx] + acc[tr];
    const double ka = general ? acc[nn + idx] : acc[nn + idx] + acc[nn + tr];
    Xa[idx] = -ka;
    if (nspin == 2) {
      const double kb = general ? acc[2 * nn + idx] : acc[2 * nn + idx] + acc[2 * nn + tr];
      Xb[idx] = -kb;
    } else {
      Xb[idx] = -ka;
    }
  }
}

// One CTA per group of NB consecutive slabs G[a,b..b+NB-1,:,:] (each N x N contiguous): streams
// the tensor exactly once and re-uses every Dt[c,d] load for NB slabs.
//   J[a,b]   = sum_cd Dt[c,d] G[a,b,c,d]
//   Xa[a,d] -= sum_c  Da[c,b] G[a,b,c,d]   (and beta)
// 8 warps; a warp covers 32*VEC consecutive columns d of one row c per load (16-byte loads when
// N is even), the 8 warps take rows c, c+8, ...
template <int VEC, int NB>
__global__ void __launch_bounds__(256) jk_stored_kernel(int N, int ngrp, const double* __restrict__ G,
                                                        const double* __restrict__ Dt,
                                                        const double* __restrict__ Da,
                                                        const double* __restrict__ Db,
                                                        double* __restrict__ J,
                                                        double* __restrict__ Xa,
                                                        double* __restrict__ Xb) {
  const int a = blockIdx.x / ngrp, b0 = (blockIdx.x % ngrp) * NB;
  const int nbv = min(NB, N - b0);                      // slabs of this group that exist
  const size_t NN = (size_t)N * N;
  const double* __restrict__ slab = G + ((size_t)a * N + b0) * NN;
  const int lane = threadIdx.x & 31, rg = threadIdx.x >> 5;
  constexpr int RG = 8;
  __shared__ double red[2][RG][32 * VEC + 1];
  double jsum[NB];
#pragma unroll
  for (int k = 0; k < NB; ++k) jsum[k] = 0.0;
  const int c0 = (int)((blockIdx.x * 2654435761u >> 12) % (unsigned)N);
  for (int d0 = 0; d0 < N; d0 += 32 * VEC) {
    const int d = d0 + lane * VEC;
    double xa[VEC], xb[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) xa[v] = xb[v] = 0.0;
    if (d < N) {
#pragma unroll 2
      for (int cc = rg; cc < N; cc += RG) {
        int c = cc + c0;                       // every CTA starts at a different row (HBM channel spread)
        if (c >= N) c -= N;
        double t[VEC];
        if (VEC == 2) {
          const double2 t2 = __ldg(reinterpret_cast<const double2*>(Dt + (size_t)c * N + d));
          t[0] = t2.x; t[VEC - 1] = t2.y;
        } else {
          t[0] = __ldg(Dt + (size_t)c * N + d);
        }
#pragma unroll
        for (int k = 0; k < NB; ++k) {
          if (k < nbv) {
            const double da = __ldg(Da + (size_t)c * N + b0 + k), db = __ldg(Db + (size_t)c * N + b0 + k);
            double g[VEC];
            if (VEC == 2) {
              const double2 g2 = *reinterpret_cast<const double2*>(slab + (size_t)k * NN + (size_t)c * N + d);
              g[0] = g2.x; g[VEC - 1] = g2.y;
            } else {
              g[0] = slab[(size_t)k * NN + (size_t)c * N + d];
            }
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
              jsum[k] = fma(t[v], g[v], jsum[k]);
              xa[v] = fma(da, g[v], xa[v]);
              xb[v] = fma(db, g[v], xb[v]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      red[0][rg][lane * VEC + v] = xa[v];
      red[1][rg][lane * VEC + v] = xb[v];
    }
    __syncthreads();
    const int col = threadIdx.x % (32 * VEC), which = threadIdx.x / (32 * VEC);
    if (which < 2 && d0 + col < N) {
      double sum = 0.0;
#pragma unroll
      for (int k = 0; k < RG; ++k) sum += red[which][k][col];
      atomicAdd((which ? Xb : Xa) + (size_t)a * N + d0 + col, -sum);
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    double v = jsum[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[0][rg][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < nbv) {
    double sum = 0.0;
    for (int k = 0; k < RG; ++k) sum += red[0][k][threadIdx.x];
    J[(size_t)a * N + b0 + threadIdx.x] = sum;
  }
}

// The same contraction with the tensor staged through shared memory by the TMA engine
// (`cp.async.bulk`, pc_async.cuh): a CTA owns NB consecutive slabs G[a, b0..b0+NB-1, :, :]; one
// tile = `rt` rows c of all NB slabs (NB contiguous pieces of rt*N doubles)
// (plus the same rows of Dt), STAGES tiles in
// flight per CTA behind full / empty mbarriers.  A producer warp issues the copies -- no registers and
// no LSU slots are spent on the 8 N^4 bytes -- eight warps consume: a thread owns the column pair
// (d, d+1) of one row group for the whole CTA, so X[a, d] stays in registers until the end and
// the row groups meet once, in shared memory.  Needs N even (16-byte rows); odd N keeps
// jk_stored_kernel.  Dynamic shared memory: stage buffers | Da/Db columns | reduction scratch.
template <int NB, int STAGES>
__global__ void __launch_bounds__(288) jk_stored_tma_kernel(int N, int ngrp, int rt, int ct,
                                                            const double* __restrict__ G,
                                                            const double* __restrict__ Dt,
                                                            const double* __restrict__ Da,
                                                            const double* __restrict__ Db,
                                                            double* __restrict__ J,
                                                            double* __restrict__ Xa,
                                                            double* __restrict__ Xb) {
  PC_DYN_SMEM(smem_raw);
  __shared__ PcMbar full[STAGES], empty[STAGES];
  __shared__ double jred[8][NB];
  const int a = blockIdx.x / ngrp, b0 = (blockIdx.x % ngrp) * NB;
  const int nbv = min(NB, N - b0);
  const size_t NN = (size_t)N * N;
  const double* __restrict__ slab = G + ((size_t)a * N + b0) * NN;
  // 8 consumer warps + 1 producer warp (its first lane issues the bulk copies and never computes,
  // so a refill is not held up by the issuing thread's own share of the previous tile)
  const int tid = threadIdx.x, nthr = 256;
  const bool producer = tid >= 256;
  const size_t tile_doubles = (size_t)(NB + 1) * rt * N;           // one stage: NB slab pieces + the Dt rows
  double* stage0 = reinterpret_cast<double*>(smem_raw);
  double* sD = stage0 + (size_t)STAGES * tile_doubles;            // [2][NB][N]: Da[c, b0+k], Db[c, b0+k]
  double* sX = sD + (size_t)2 * NB * N;                            // [2][rgn][2 ct]
  const int ntiles = (N + rt - 1) / rt;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) { pc_mbar_init(&full[s], 1); pc_mbar_init(&empty[s], nthr); }
    pc_mbar_fence_init();
  }
  __syncthreads();
  // Every CTA walks its slabs from a different first tile (the sums over c do not care): CTAs that
  // run side by side otherwise read the same offsets of slabs that lie a multiple of 8 N^2 bytes
  // apart, and those land on a subset of the HBM channels (ncu: dram__cycles_active min 29 % /
  // max 58 % over the channels, 4.45 TB/s; torch.sum over the same tensor reaches 6.3 TB/s)
  const int t0 = (int)((blockIdx.x * 2654435761u >> 12) % (unsigned)ntiles);
  auto issue = [&](int tl) {                                       // thread 0 only; tl = position in this CTA's order
    const int s = tl % STAGES;
    const int t = (tl + t0) % ntiles;
    const int rows = min(rt, N - t * rt);
    const unsigned bytes = (unsigned)((size_t)rows * N * sizeof(double));
    pc_mbar_arrive_expect_tx(&full[s], bytes * (unsigned)(nbv + 1));
    for (int k = 0; k < nbv; ++k)
      pc_bulk_g2s(stage0 + (size_t)s * tile_doubles + (size_t)k * rt * N, slab + (size_t)k * NN + (size_t)t * rt * N, bytes, &full[s]);
    // the matching rows of Dt ride along (from L2): no global load is left in the consumer loop
    pc_bulk_g2s(stage0 + (size_t)s * tile_doubles + (size_t)NB * rt * N, Dt + (size_t)t * rt * N, bytes, &full[s]);
  };
  if (tid == 256)
    for (int tl = 0; tl < min(STAGES, ntiles); ++tl) issue(tl);
  // the density columns this CTA needs for the exchange part
  for (int idx = tid; idx < N * NB && !producer; idx += nthr) {
    const int c = idx / NB, k = idx % NB;
    const bool ok = k < nbv;
    sD[(size_t)k * N + c] = ok ? __ldg(Da + (size_t)c * N + b0 + k) : 0.0;
    sD[(size_t)(NB + k) * N + c] = ok ? __ldg(Db + (size_t)c * N + b0 + k) : 0.0;
  }
  __syncthreads();
  if (tid == 256) {
    for (int tl = STAGES; tl < ntiles; ++tl) {
      // the k-th use of a stage waits until the consumers have left its (k-1)-th tile
      pc_mbar_wait(&empty[tl % STAGES], (unsigned)((tl / STAGES - 1) & 1));
      issue(tl);
    }
  }
  const int rgn = nthr / ct;                                       // row groups
  const int cg = tid % ct, rgi = tid / ct;
  const int d = 2 * cg;
  const bool active = d < N && rgi < rgn && !producer;
  double jsum[NB], xa0 = 0.0, xa1 = 0.0, xb0 = 0.0, xb1 = 0.0;
#pragma unroll
  for (int k = 0; k < NB; ++k) jsum[k] = 0.0;
  for (int tl = 0; tl < ntiles && !producer; ++tl) {
    const int s = tl % STAGES;
    const int t = (tl + t0) % ntiles;
    pc_mbar_wait(&full[s], (unsigned)((tl / STAGES) & 1));
    if (active) {
      const int rows = min(rt, N - t * rt);
      const double* __restrict__ st = stage0 + (size_t)s * tile_doubles;
      for (int r = rgi; r < rows; r += rgn) {
        const int c = t * rt + r;
        const double2 t2 = *reinterpret_cast<const double2*>(st + ((size_t)NB * rt + r) * N + d);
#pragma unroll
        for (int k = 0; k < NB; ++k) {
          // slabs beyond the tensor (k >= nbv) were never copied: their density factors are 0 and
          // the stage words are whatever an earlier tile left -- skip them (uniform per CTA)
          if (k < nbv) {
            const double2 g2 = *reinterpret_cast<const double2*>(st + ((size_t)k * rt + r) * N + d);
            const double da = sD[(size_t)k * N + c], db = sD[(size_t)(NB + k) * N + c];
            jsum[k] = fma(t2.x, g2.x, jsum[k]);
            jsum[k] = fma(t2.y, g2.y, jsum[k]);
            xa0 = fma(da, g2.x, xa0); xa1 = fma(da, g2.y, xa1);
            xb0 = fma(db, g2.x, xb0); xb1 = fma(db, g2.y, xb1);
          }
        }
      }
    }
    pc_mbar_arrive(&empty[s]);
  }
  // ---- exchange: sum the row groups, one atomic per column (the ngrp CTAs of row a meet here)
  if (rgi < rgn && !producer) {
    sX[((size_t)0 * rgn + rgi) * 2 * ct + 2 * cg] = xa0; sX[((size_t)0 * rgn + rgi) * 2 * ct + 2 * cg + 1] = xa1;
    sX[((size_t)1 * rgn + rgi) * 2 * ct + 2 * cg] = xb0; sX[((size_t)1 * rgn + rgi) * 2 * ct + 2 * cg + 1] = xb1;
  }
  __syncthreads();
  for (int idx = tid; idx < 2 * N && !producer; idx += nthr) {
    const int which = idx / N, col = idx % N;
    double sum = 0.0;
    for (int g = 0; g < rgn; ++g) sum += sX[((size_t)which * rgn + g) * 2 * ct + col];
    atomicAdd((which ? Xb : Xa) + (size_t)a * N + col, -sum);
  }
  // ---- Coulomb: block sum of the NB partial dot products
  const int lane = tid & 31, w = tid >> 5;
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    double v = active ? jsum[k] : 0.0;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0 && !producer) jred[w][k] = v;
  }
  __syncthreads();
  if (tid < nbv) {
    double sum = 0.0;
    for (int k = 0; k < (nthr >> 5); ++k) sum += jred[k][tid];
    J[(size_t)a * N + b0 + tid] = sum;
  }
}

// Batched form of jk_stored_kernel (NOCI co-density pairs, Methods/noci.py:247,275,291): the slab
// values are loaded once and contracted with up to NS density sets, so the tensor is streamed
// once per NS sets instead of once per set.  D: [set][Dt | Da | Db], out: [set][J | Xa | Xb],
// both with `stride` doubles per set.
template <int VEC, int NB, int NS>
__global__ void __launch_bounds__(256) jk_stored_batch_kernel(int N, int ngrp, int nset, size_t stride,
                                                              const double* __restrict__ G,
                                                              const double* __restrict__ D,
                                                              double* __restrict__ out) {
  const int a = blockIdx.x / ngrp, b0 = (blockIdx.x % ngrp) * NB;
  const int nbv = min(NB, N - b0);
  const size_t NN = (size_t)N * N;
  const double* __restrict__ slab = G + ((size_t)a * N + b0) * NN;
  const int lane = threadIdx.x & 31, rg = threadIdx.x >> 5;
  constexpr int RG = 8;
  __shared__ double red[2][RG][32 * VEC + 1];
  double jsum[NS][NB];
#pragma unroll
  for (int s = 0; s < NS; ++s)
#pragma unroll
    for (int k = 0; k < NB; ++k) jsum[s][k] = 0.0;
  for (int d0 = 0; d0 < N; d0 += 32 * VEC) {
    const int d = d0 + lane * VEC;
    double xa[NS][VEC], xb[NS][VEC];
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
      for (int v = 0; v < VEC; ++v) xa[s][v] = xb[s][v] = 0.0;
    if (d < N) {
      for (int c = rg; c < N; c += RG) {
        double g[NB][VEC];
#pragma unroll
        for (int k = 0; k < NB; ++k) {
#pragma unroll
          for (int v = 0; v < VEC; ++v) g[k][v] = 0.0;
          if (k < nbv) {
            if (VEC == 2) {
              const double2 g2 = *reinterpret_cast<const double2*>(slab + (size_t)k * NN + (size_t)c * N + d);
              g[k][0] = g2.x; g[k][VEC - 1] = g2.y;
            } else {
              g[k][0] = slab[(size_t)k * NN + (size_t)c * N + d];
            }
          }
        }
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          if (s < nset) {
            const double* __restrict__ Dt = D + (size_t)s * stride;
            const double* __restrict__ Da = Dt + NN;
            const double* __restrict__ Db = Dt + 2 * NN;
            double t[VEC];
            if (VEC == 2) {
              const double2 t2 = __ldg(reinterpret_cast<const double2*>(Dt + (size_t)c * N + d));
              t[0] = t2.x; t[VEC - 1] = t2.y;
            } else {
              t[0] = __ldg(Dt + (size_t)c * N + d);
            }
#pragma unroll
            for (int k = 0; k < NB; ++k) {
              if (k < nbv) {
                const double da = __ldg(Da + (size_t)c * N + b0 + k), db = __ldg(Db + (size_t)c * N + b0 + k);
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                  jsum[s][k] = fma(t[v], g[k][v], jsum[s][k]);
                  xa[s][v] = fma(da, g[k][v], xa[s][v]);
                  xb[s][v] = fma(db, g[k][v], xb[s][v]);
                }
              }
            }
          }
        }
      }
    }
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      if (s < nset) {                                   // uniform over the block
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          red[0][rg][lane * VEC + v] = xa[s][v];
          red[1][rg][lane * VEC + v] = xb[s][v];
        }
        __syncthreads();
        const int col = threadIdx.x % (32 * VEC), which = threadIdx.x / (32 * VEC);
        if (which < 2 && d0 + col < N) {
          double sum = 0.0;
#pragma unroll
          for (int k = 0; k < RG; ++k) sum += red[which][k][col];
          atomicAdd(out + (size_t)s * stride + (which ? 2 : 1) * NN + (size_t)a * N + d0 + col, -sum);
        }
        __syncthreads();
      }
    }
  }
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    if (s < nset) {
#pragma unroll
      for (int k = 0; k < NB; ++k) {
        double v = jsum[s][k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[0][rg][k] = v;
      }
      __syncthreads();
      if (threadIdx.x < nbv) {
        double sum = 0.0;
        for (int k = 0; k < RG; ++k) sum += red[0][k][threadIdx.x];
        out[(size_t)s * stride + (size_t)a * N + b0 + threadIdx.x] = sum;
      }
      __syncthreads();
    }
  }
}

// flags |= 1 if any of Dt, Da, Db is not symmetric; flags |= 2 if Da != Db (bitwise compare of
// values, the same test the host mirror would make with numpy.array_equal)
__global__ void classify_kernel(int N, const double* __restrict__ Dt, const double* __restrict__ Da,
                                const double* __restrict__ Db, int* flags) {
  const size_t nn = (size_t)N * N;
  int f = 0;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < nn;
       idx += (size_t)gridDim.x * blockDim.x) {
    const size_t r = idx / N, c = idx % N, tr = c * N + r;
    if (r < c && (Dt[idx] != Dt[tr] || Da[idx] != Da[tr] || Db[idx] != Db[tr])) f |= 1;
    if (Da[idx] != Db[idx]) f |= 2;
  }
  if (f) atomicOr(flags, f);
}

// register-resident DFMA loop: 8 independent chains per thread
__global__ void dfma_peak_kernel(double* out, int iters, double seed) {
  double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5,
         a6 = seed + 6, a7 = seed + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
      a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
  }
  const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == 123.456) out[0] = s;
}

}  // namespace
