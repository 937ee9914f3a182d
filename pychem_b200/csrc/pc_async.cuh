// pc_async.cuh -- mbarrier + bulk asynchronous copy (TMA engine, 1-D `cp.async.bulk`) helpers.
//
// sm_100a: a shared-memory mbarrier counts (a) thread arrivals and (b) the bytes of the bulk
// copies that were announced with expect_tx; a phase completes when both reach zero.  One thread
// issues `cp.async.bulk.shared::cluster.global` copies of whole contiguous tiles, the TMA engine
// moves them without occupying registers or LSU issue slots, consumers spin on try_wait.parity.
//
// The host emulation (tests/emu) replaces the PTX between the EMU_SKIP markers by a functional
// model with the same interface: copies are synchronous memcpy's, waits yield the fiber.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef PC_HOST_EMU
// PC_EMU_SKIP_BEGIN
struct alignas(8) PcMbar { unsigned long long state; };

__device__ __forceinline__ uint32_t pc_smem_addr(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void pc_mbar_init(PcMbar* b, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pc_smem_addr(b)), "r"(count) : "memory");
}
// make the initialised barriers visible to the asynchronous proxy (the TMA engine)
__device__ __forceinline__ void pc_mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void pc_mbar_arrive_expect_tx(PcMbar* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pc_smem_addr(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pc_mbar_arrive(PcMbar* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(pc_smem_addr(b)) : "memory");
}
__device__ __forceinline__ void pc_mbar_wait(PcMbar* b, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "PC_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra PC_DONE_%=;\n"
      "bra PC_WAIT_%=;\n"
      "PC_DONE_%=:\n"
      "}\n" ::"r"(pc_smem_addr(b)), "r"(parity) : "memory");
}
// bytes: multiple of 16; dst (shared) and src (global) 16-byte aligned
__device__ __forceinline__ void pc_bulk_g2s(void* dst, const void* src, unsigned bytes, PcMbar* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   pc_smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(pc_smem_addr(b))
               : "memory");
}
#define PC_DYN_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
// PC_EMU_SKIP_END
#else
struct alignas(8) PcMbar { int count, pending, phase; long long tx; };
inline void pc_mbar_check(PcMbar* b) {
  if (b->pending == 0 && b->tx == 0) { b->phase ^= 1; b->pending = b->count; }
}
inline void pc_mbar_init(PcMbar* b, unsigned count) { b->count = b->pending = (int)count; b->phase = 0; b->tx = 0; }
inline void pc_mbar_fence_init() {}
inline void pc_mbar_arrive_expect_tx(PcMbar* b, unsigned bytes) { b->tx += bytes; b->pending -= 1; pc_mbar_check(b); }
inline void pc_mbar_arrive(PcMbar* b) { b->pending -= 1; pc_mbar_check(b); }
inline void pc_mbar_wait(PcMbar* b, unsigned parity) { while ((unsigned)(b->phase & 1) == parity) pcemu::yield(); }
inline void pc_bulk_g2s(void* dst, const void* src, unsigned bytes, PcMbar* b) {
  if ((bytes & 15u) || ((uintptr_t)dst & 15u) || ((uintptr_t)src & 15u)) pcemu::die("cp.async.bulk: size / address not 16-byte aligned");
  std::memcpy(dst, src, bytes);
  b->tx -= bytes;
  pc_mbar_check(b);
}
#define PC_DYN_SMEM(name) unsigned char* name = pcemu::dyn_smem()
#endif
