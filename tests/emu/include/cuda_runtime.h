// TEST INFRASTRUCTURE ONLY -- never part of the product path.
//
// A minimal stand-in for <cuda_runtime.h> that lets g++ compile the UNMODIFIED kernel sources of
// pychem_b200/csrc (pc_common.cuh, pc_api.cu, the generated eri_*.cu) for the host, so that the
// `-m "not gpu"` tests can execute the real device code -- task decode, recursions, warp-level
// reductions, digestion "atomics" -- on a machine without a GPU and compare it with the oracle
// (tests/emu/build_emu.py, tests/test_emu_cpu.py).
//
//   * every CUDA thread of a block is a ucontext fiber; blocks run one after another
//   * warp collectives (__shfl_*_sync, __reduce_*_sync, __activemask, ...) are rendezvous points:
//     a fiber yields until all live lanes of its warp arrived, then the results are computed
//   * __syncthreads() is the same at block level; __shared__ is `static` (one block at a time)
//   * atomics are plain read-modify-writes (one OS thread)
//   * the runtime API is a host-memory shim: cudaMalloc = malloc, copies = memcpy, streams,
//     events and graphs are inert
//
// The product library (pychem_b200/libpychem_b200.so) is built by nvcc from the same sources and
// knows nothing about this directory; pychem_b200/_lib.py never loads the emulation.
#pragma once
#include <ucontext.h>
#include <sys/mman.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

// ---------------------------------------------------------------------------------------------
// keywords
// ---------------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static
#define __restrict__

// ---------------------------------------------------------------------------------------------
// vector types
// ---------------------------------------------------------------------------------------------
struct alignas(16) double2 { double x, y; };
struct alignas(8) int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
inline int4 make_int4(int x, int y, int z, int w) { int4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
struct uint3 { unsigned x, y, z; };
struct dim3 { unsigned x = 1, y = 1, z = 1; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
inline int2 make_int2(int x, int y) { int2 r; r.x = x; r.y = y; return r; }

// ---------------------------------------------------------------------------------------------
// fibers
// ---------------------------------------------------------------------------------------------
namespace pcemu {
enum { ST_RUN = 0, ST_WARP = 1, ST_BLOCK = 2, ST_DONE = 3 };
enum { OP_NONE = 0, OP_SHFL_IDX, OP_SHFL_DOWN, OP_SHFL_UP, OP_SHFL_XOR, OP_RED_OR, OP_RED_AND, OP_RED_MAX_S,
       OP_RED_MIN_S, OP_RED_ADD, OP_BALLOT, OP_ACTIVEMASK };

struct Lane {
  ucontext_t ctx;
  int state, op, iarg;
  uint64_t arg, res;
};

struct Ctx {
  unsigned tid = 0, bid = 0, bdim = 1, gdim = 1;
  std::vector<Lane> lanes;
  std::vector<char*> stacks;
  ucontext_t sched;
  int cur = -1;
  const std::function<void()>* body = nullptr;
  unsigned long long launches = 0, switches = 0;
};

inline Ctx& ctx() {
  static Ctx c;
  return c;
}

constexpr size_t STACK_BYTES = size_t(2) << 20;

inline void die(const char* msg) {
  std::fprintf(stderr, "pcemu: %s\n", msg);
  std::abort();
}

inline void entry() {
  Ctx& c = ctx();
  (*c.body)();
  c.lanes[c.cur].state = ST_DONE;      // returning resumes the scheduler (uc_link)
}

inline void resolve_warp(Ctx& c, unsigned w0, unsigned w1) {
  int op = OP_NONE;
  unsigned alive = 0;
  for (unsigned t = w0; t < w1; ++t)
    if (c.lanes[t].state == ST_WARP) {
      alive |= 1u << (t - w0);
      if (op == OP_NONE) op = c.lanes[t].op;
      else if (op != c.lanes[t].op) die("lanes of one warp wait at different collectives (divergent collective)");
    }
  uint64_t red = 0;
  bool first = true;
  for (unsigned t = w0; t < w1; ++t) {
    if (!(alive >> (t - w0) & 1)) continue;
    const uint64_t a = c.lanes[t].arg;
    switch (op) {
      case OP_RED_OR: red |= a; break;
      case OP_RED_AND: red = first ? a : (red & a); break;
      case OP_RED_ADD: red += a; break;
      case OP_RED_MAX_S: red = first ? a : (uint64_t)std::max((int64_t)red, (int64_t)a); break;
      case OP_RED_MIN_S: red = first ? a : (uint64_t)std::min((int64_t)red, (int64_t)a); break;
      case OP_BALLOT: if (a) red |= 1u << (t - w0); break;
      default: break;
    }
    first = false;
  }
  for (unsigned t = w0; t < w1; ++t) {
    if (!(alive >> (t - w0) & 1)) continue;
    Lane& L = c.lanes[t];
    const int lane = (int)(t - w0);
    int src = lane;
    switch (op) {
      case OP_SHFL_IDX: src = L.iarg & 31; break;
      case OP_SHFL_DOWN: src = lane + L.iarg; break;
      case OP_SHFL_UP: src = lane - L.iarg; break;
      case OP_SHFL_XOR: src = lane ^ L.iarg; break;
      default: break;
    }
    switch (op) {
      case OP_SHFL_IDX: case OP_SHFL_DOWN: case OP_SHFL_UP: case OP_SHFL_XOR:
        // out-of-range or exited source lane: the lane keeps its own value
        L.res = (src >= 0 && src < 32 && (alive >> src & 1)) ? c.lanes[w0 + src].arg : L.arg;
        break;
      case OP_ACTIVEMASK: L.res = alive; break;
      default: L.res = red; break;
    }
  }
  for (unsigned t = w0; t < w1; ++t)
    if (alive >> (t - w0) & 1) c.lanes[t].state = ST_RUN;
}

inline void run_block(Ctx& c) {
  const unsigned n = c.bdim;
  for (;;) {
    bool progressed = false;
    for (unsigned t = 0; t < n; ++t)
      if (c.lanes[t].state == ST_RUN) {
        c.cur = (int)t;
        c.tid = t;
        ++c.switches;
        swapcontext(&c.sched, &c.lanes[t].ctx);
        progressed = true;
      }
    unsigned done = 0, bwait = 0;
    for (unsigned w0 = 0; w0 < n; w0 += 32) {
      const unsigned w1 = std::min(n, w0 + 32);
      unsigned alive = 0, waiting = 0;
      for (unsigned t = w0; t < w1; ++t) {
        const int st = c.lanes[t].state;
        if (st == ST_DONE) ++done;
        else ++alive;
        if (st == ST_WARP) ++waiting;
        if (st == ST_BLOCK) ++bwait;
      }
      if (alive && waiting == alive) {
        resolve_warp(c, w0, w1);
        progressed = true;
      }
    }
    if (done == n) return;
    if (bwait && bwait + done == n) {
      for (unsigned t = 0; t < n; ++t)
        if (c.lanes[t].state == ST_BLOCK) c.lanes[t].state = ST_RUN;
      progressed = true;
    }
    if (!progressed) die("deadlock: some threads wait at a collective the others never reach");
  }
}

inline void launch(unsigned grid, unsigned block, const std::function<void()>& body) {
  Ctx& c = ctx();
  if (c.body) die("nested kernel launch");
  if (block == 0 || block > 1024) die("bad block size");
  c.body = &body;
  c.gdim = grid;
  c.bdim = block;
  ++c.launches;
  while (c.stacks.size() < block) {
    void* p = mmap(nullptr, STACK_BYTES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (p == MAP_FAILED) die("mmap of a fiber stack failed");
    c.stacks.push_back((char*)p);
  }
  if (c.lanes.size() < block) c.lanes.resize(block);     // never resized while fibers are live
  for (unsigned b = 0; b < grid; ++b) {
    c.bid = b;
    for (unsigned t = 0; t < block; ++t) {
      Lane& L = c.lanes[t];
      getcontext(&L.ctx);
      L.ctx.uc_stack.ss_sp = c.stacks[t];
      L.ctx.uc_stack.ss_size = STACK_BYTES;
      L.ctx.uc_link = &c.sched;
      makecontext(&L.ctx, entry, 0);
      L.state = ST_RUN;
      L.op = OP_NONE;
    }
    run_block(c);
  }
  c.body = nullptr;
  c.cur = -1;
}

inline uint64_t collective(int op, uint64_t arg, int iarg) {
  Ctx& c = ctx();
  Lane& L = c.lanes[c.cur];
  L.op = op;
  L.arg = arg;
  L.iarg = iarg;
  L.state = ST_WARP;
  swapcontext(&L.ctx, &c.sched);
  return L.res;
}

// give the other fibers of the block a turn (spin-waits on an emulated mbarrier)
inline void yield() {
  Ctx& c = ctx();
  Lane& L = c.lanes[c.cur];
  L.state = ST_RUN;
  swapcontext(&L.ctx, &c.sched);
}

// dynamic shared memory: one static, 128-byte aligned buffer (blocks run one after another)
inline unsigned char* dyn_smem() {
  alignas(128) static unsigned char buf[232448];
  return buf;
}

inline void block_barrier() {
  Ctx& c = ctx();
  Lane& L = c.lanes[c.cur];
  L.state = ST_BLOCK;
  swapcontext(&L.ctx, &c.sched);
}

template <typename T>
inline uint64_t pack(T v) {
  static_assert(sizeof(T) <= 8, "shuffle operand wider than 64 bits");
  uint64_t a = 0;
  std::memcpy(&a, &v, sizeof(T));
  return a;
}
template <typename T>
inline T unpack(uint64_t a) {
  T v;
  std::memcpy(&v, &a, sizeof(T));
  return v;
}

inline uint3 tid3() { return uint3{ctx().tid, 0, 0}; }
inline uint3 bid3() { return uint3{ctx().bid, 0, 0}; }
inline uint3 bdim3() { return uint3{ctx().bdim, 1, 1}; }
inline uint3 gdim3() { return uint3{ctx().gdim, 1, 1}; }

// MUFU.RSQ64H stand-in: a reciprocal square root good to single precision (the kernels refine it)
inline double rsqrt_seed(double x) { return (double)(float)(1.0 / std::sqrt(x)); }
}  // namespace pcemu

#define threadIdx (pcemu::tid3())
#define blockIdx (pcemu::bid3())
#define blockDim (pcemu::bdim3())
#define gridDim (pcemu::gdim3())

// ---------------------------------------------------------------------------------------------
// device intrinsics
// ---------------------------------------------------------------------------------------------
template <typename T> inline T __ldg(const T* p) { return *p; }
template <typename T> inline T __shfl_sync(unsigned, T v, int src, int = 32) { return pcemu::unpack<T>(pcemu::collective(pcemu::OP_SHFL_IDX, pcemu::pack(v), src)); }
template <typename T> inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) { return pcemu::unpack<T>(pcemu::collective(pcemu::OP_SHFL_DOWN, pcemu::pack(v), (int)d)); }
template <typename T> inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) { return pcemu::unpack<T>(pcemu::collective(pcemu::OP_SHFL_UP, pcemu::pack(v), (int)d)); }
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return pcemu::unpack<T>(pcemu::collective(pcemu::OP_SHFL_XOR, pcemu::pack(v), m)); }
inline unsigned __reduce_or_sync(unsigned, unsigned v) { return (unsigned)pcemu::collective(pcemu::OP_RED_OR, v, 0); }
inline unsigned __reduce_and_sync(unsigned, unsigned v) { return (unsigned)pcemu::collective(pcemu::OP_RED_AND, v, 0); }
inline int __reduce_max_sync(unsigned, int v) { return (int)(int64_t)pcemu::collective(pcemu::OP_RED_MAX_S, (uint64_t)(int64_t)v, 0); }
inline int __reduce_min_sync(unsigned, int v) { return (int)(int64_t)pcemu::collective(pcemu::OP_RED_MIN_S, (uint64_t)(int64_t)v, 0); }
inline int __reduce_add_sync(unsigned, int v) { return (int)(int64_t)pcemu::collective(pcemu::OP_RED_ADD, (uint64_t)(int64_t)v, 0); }
inline unsigned __ballot_sync(unsigned, int pred) { return (unsigned)pcemu::collective(pcemu::OP_BALLOT, pred ? 1 : 0, 0); }
inline int __any_sync(unsigned, int pred) { return pcemu::collective(pcemu::OP_BALLOT, pred ? 1 : 0, 0) != 0; }
inline int __all_sync(unsigned m, int pred) { return pcemu::collective(pcemu::OP_BALLOT, pred ? 0 : 1, 0) == 0; }
inline unsigned __activemask() { return (unsigned)pcemu::collective(pcemu::OP_ACTIVEMASK, 0, 0); }
inline void __syncthreads() { pcemu::block_barrier(); }
inline void __syncwarp(unsigned = 0xffffffffu) { pcemu::collective(pcemu::OP_RED_OR, 0, 0); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline double atomicAdd(double* p, double v) { const double o = *p; *p = o + v; return o; }
inline int atomicAdd(int* p, int v) { const int o = *p; *p = o + v; return o; }
inline int atomicOr(int* p, int v) { const int o = *p; *p = o | v; return o; }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
inline double min(double a, double b) { return a < b ? a : b; }
inline double max(double a, double b) { return a > b ? a : b; }
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }

// ---------------------------------------------------------------------------------------------
// runtime API (host memory, inert streams / events / graphs)
// ---------------------------------------------------------------------------------------------
enum cudaError_t { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorNotSupported = 801 };
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
typedef void* cudaGraph_t;
typedef void* cudaGraphExec_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
enum cudaStreamCaptureMode { cudaStreamCaptureModeGlobal = 0, cudaStreamCaptureModeThreadLocal = 1, cudaStreamCaptureModeRelaxed = 2 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
struct cudaPointerAttributes { cudaMemoryType type; int device; void* devicePointer; void* hostPointer; };
struct cudaDeviceProp { int multiProcessorCount; char name[256]; };

inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { std::memset(p, 0, sizeof(*p)); p->multiProcessorCount = 1; std::strcpy(p->name, "pcemu host emulation"); return cudaSuccess; }
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
// every non-null pointer counts as device memory: host buffers are used in place
inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p) { a->type = p ? cudaMemoryTypeDevice : cudaMemoryTypeUnregistered; a->device = 0; a->devicePointer = (void*)p; a->hostPointer = (void*)p; return cudaSuccess; }
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
template <typename F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = std::malloc(1); return cudaSuccess; }
inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = std::malloc(1); return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { std::free(s); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = std::malloc(n); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = std::malloc(1); return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = std::malloc(1); return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { std::free(e); return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 1.0f; return cudaSuccess; }
// graphs cannot be replayed by the emulation: capture is refused (the library is built with
// PC_HOST_EMU, which switches graph replay off, so these are never reached)
inline cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode) { return cudaErrorNotSupported; }
inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t*) { return cudaErrorNotSupported; }
inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t*, cudaGraph_t, unsigned long long = 0) { return cudaErrorNotSupported; }
inline cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t) { return cudaErrorNotSupported; }
inline cudaError_t cudaGraphDestroy(cudaGraph_t) { return cudaSuccess; }
inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return cudaSuccess; }
