// pc_generic.cu -- instantiations and host launcher of the generic (f-shell) ERI kernel.
#include "pc_generic.cuh"

cudaError_t pc_launch_generic(int mode, const PcEriArgs& A, const PcGenClass& C, cudaStream_t st) {
  if (A.nwarps <= 0) return cudaSuccess;
  if (C.nthreads <= 0 || C.nthreads % 64 != 0 || !C.scratch) return cudaErrorInvalidValue;
  const int block = 64;
  const unsigned grid = (unsigned)(C.nthreads / block);
  switch (mode) {
    case PC_MODE_BLOCKS: eri_generic_kernel<PC_MODE_BLOCKS><<<grid, block, 0, st>>>(A, C); break;
    case PC_MODE_TENSOR: eri_generic_kernel<PC_MODE_TENSOR><<<grid, block, 0, st>>>(A, C); break;
    case PC_MODE_JK_RHF: eri_generic_kernel<PC_MODE_JK_RHF><<<grid, block, 0, st>>>(A, C); break;
    case PC_MODE_JK_UHF: eri_generic_kernel<PC_MODE_JK_UHF><<<grid, block, 0, st>>>(A, C); break;
    case PC_MODE_JK_GEN: eri_generic_kernel<PC_MODE_JK_GEN><<<grid, block, 0, st>>>(A, C); break;
    case PC_MODE_NULL: eri_generic_kernel<PC_MODE_NULL><<<grid, block, 0, st>>>(A, C); break;
    case PC_MODE_BLOCKS_SCAT: eri_generic_kernel<PC_MODE_BLOCKS_SCAT><<<grid, block, 0, st>>>(A, C); break;
    case PC_MODE_TENSOR_SCAT: eri_generic_kernel<PC_MODE_TENSOR_SCAT><<<grid, block, 0, st>>>(A, C); break;
    case PC_MODE_JK_GEN_BATCH: eri_generic_kernel<PC_MODE_JK_GEN_BATCH><<<grid, block, 0, st>>>(A, C); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}
