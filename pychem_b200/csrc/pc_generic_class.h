// pc_generic_class.h -- class descriptor of the generic (f-shell) ERI kernel, shared by the
// kernel (pc_generic.cuh) and the host code that sizes its scratch and launches it (pc_api.cu).
#pragma once
#include <cuda_runtime.h>
#include "pc_common.cuh"

#define PCG_LMAX 3                     // highest shell angular momentum (f)
#define PCG_LPAIR (2 * PCG_LMAX)       // highest pair angular momentum
#define PCG_NCART_MAX 10               // Cartesian components of an f shell
#define PCG_HRR_BUF 150                // largest HRR level of one side: (d+f levels of an (f f| pair)

// Class descriptor of one launch (second kernel argument, constant bank).
struct PcGenClass {
  int lx1, ly1, lx2, ly2;      // shells in kernel order: primary (higher l) first in each pair
  int nx1, ny1, nx2, ny2;      // basis functions per shell: 2l+1, or 6 for Cartesian d
  int cart_d;                  // d shells carry their six Cartesians (Cartesian_L = [2])
  int scat;                    // scattering fundamentals (ints_type = 1)
  int L;                       // lx1 + ly1 + lx2 + ly2
  int offV[PCG_LPAIR + 1][PCG_LPAIR + 1];   // VRR table: first word of block (la, lc)
  int sizeA, sizeB;            // scratch regions per thread (doubles), see pcg_quartet
  int offT1, offG;             // inside region A, after the primitive loop
  double* scratch;             // [(sizeA + sizeB)][nthreads]
  int nthreads;                // threads of the launch (multiple of 32) = scratch stride
};

__host__ __device__ inline int pcg_ncart(int l) { return (l + 1) * (l + 2) / 2; }
__host__ __device__ inline int pcg_ncum(int l) { return (l + 1) * (l + 2) * (l + 3) / 6; }  // components with total <= l
// position of (lx, ly, lz) inside its shell: lx descending, then ly descending
// (two_electron_vrr.c:27-29, angmom_index.c:3-15)
__host__ __device__ inline int pcg_cidx(int ly, int lz) { return (ly + lz) * (ly + lz + 1) / 2 + lz; }

// host launcher (pc_generic.cu)
cudaError_t pc_launch_generic(int mode, const PcEriArgs& args, const PcGenClass& cls, cudaStream_t stream);
