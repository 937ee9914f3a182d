// pc_api.cu -- host side of the C ABI declared in include/pychem_b200.h.
//
// Owns the device-resident basis: shell table, class/contraction-bucketed shell-pair tables
// (SoA, sorted by Schwarz maximum), the Boys interpolation table, and the screened quartet plan.
// Reference sites are cited per function in the header.
#include "../../include/pychem_b200.h"
#include "pc_common.cuh"
#include "pc_one_electron.cuh"
#include "pc_generic_class.h"
#include "pc_jk_kernels.cuh"
#include "pc_boys_table.h"

#include <algorithm>
#include <atomic>
#include <functional>
#include <thread>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <numeric>
#include <string>
#include <vector>


namespace {

thread_local std::string g_err;

int fail(const std::string& msg) {
  g_err = msg;
  return 1;
}
#define PC_CUDA(call)                                                                         \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(std::string(#call) + ": " + cudaGetErrorString(e_));                        \
  } while (0)

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaError_t alloc(size_t count) {
    release();
    n = count;
    if (count == 0) return cudaSuccess;
    return cudaMalloc((void**)&p, count * sizeof(T));
  }
  cudaError_t upload(const std::vector<T>& v, cudaStream_t st) {
    // same size as before (tables rebuilt in a new order): keep the allocation, the copy is
    // ordered after earlier work on the stream
    cudaError_t e = (p && n == v.size()) ? cudaSuccess : alloc(v.size());
    if (e != cudaSuccess || v.empty()) return e;
    return cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st);
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  ~DevBuf() { release(); }
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
};

constexpr double PC_PRIM_EPS = 1.0e-24;   // primitive-pair prefactor cut-off (see upload_kind)

inline int ncart(int l) { return (l + 1) * (l + 2) / 2; }
inline int pair_class(int lx, int ly) { return lx * (lx + 1) / 2 + ly; }  // ss ps pp ds dp dd fs fp fd ff
constexpr int PC_NPC_GEN = 6;      // pair classes below this have generated kernels (s, p, d shells)
constexpr int PC_NPC = 10;         // all pair classes up to (f f|
inline void class_l(int pc, int& lx, int& ly) {
  lx = 0;
  while ((lx + 1) * (lx + 2) / 2 <= pc) ++lx;
  ly = pc - lx * (lx + 1) / 2;
}
inline int pair_L(int pc) { int lx, ly; class_l(pc, lx, ly); return lx + ly; }

struct Shell {
  int l, K, first_fn, nfn, poff;
  double A[3];
};

struct HostPair {
  int a, b;       // a <= b, shell indices as given
  int x, y;       // primary (higher l) / secondary
  int swapped;    // x == b
  int kind, pos;  // bucket and position inside it
  double pmax;
  size_t prim_off;  // first record of this pair in pc_basis::prim_host (6 doubles per primitive pair)
  int keff;         // significant primitive pairs
};

struct Kind {
  int lx, ly, K, pc;
  std::vector<int> pairs;  // pair ids in bucket order
  // after pc_schwarz the bucket is ordered in GROUPS of pairs sharing their primary shell x,
  // groups by descending group maximum, pairs inside a group by descending Schwarz maximum
  std::vector<int> gstart;      // [ngroups + 1] first position of every group
  std::vector<double> pm;       // [n] Schwarz maximum by position
  std::vector<int> keff_h;      // [n] significant primitive pairs by position
  DevBuf<int> fx, fy, pid, keff;
  DevBuf<double> xy, prim, pmd, rec;
  PcPairKind view() const {
    PcPairKind v;
    v.n = (int)pairs.size();
    v.K = K;
    v.fx = fx.p; v.fy = fy.p; v.pid = pid.p; v.keff = keff.p; v.xy = xy.p; v.prim = prim.p;
    v.pm = pmd.p;
    v.rec = rec.p;
    return v;
  }
};

struct PlanItem {
  int kb, kk;           // bra / ket kind
  int same;
  long long total;      // all tasks of the bucket pair
  long long begin, count;  // this rank's slice (tasks; cut at segment boundaries)
  long long total_q, count_q;  // quartets of the bucket pair / of this rank's slice
  double prim_exec;        // primitive quartets actually visited by the whole bucket pair
  int nseg;
  // device tables of the item: views into the plan's three pooled buffers (pc_basis::plan_*)
  const long long* seg_off;
  const int* seg_ij;       // int2 per segment
  const int* warp_s0;
  const int* seg_rec;      // int4 per segment (+ sentinel): {off lo, off hi, ij.x, ij.y}
};

// the bucket pairs of one angular-momentum class that go into one fused kernel launch
struct LaunchGroup {
  int pcb, pck;
  std::vector<int> items;   // indices into pc_basis::plan
  double cost;
  // classes with an f shell (pc_generic.cuh): the launch's own scratch, so that the launches of
  // one build can run side by side
  DevBuf<double>* scratch = nullptr;
  int gen_threads = 0;
};

bool is_device_ptr(const void* p) {
  if (!p) return false;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// ------------------------------------------------------------------------------------------
// Boys table: cubic in sT = T/(2d) per interval, third-order Taylor about the interval centre
// (consumer: Methods/c_ints/two_electron_fundamentals.c:62-75; table blob missing upstream)
// ------------------------------------------------------------------------------------------
std::vector<double> make_boys_table() {
  std::vector<double> tab((size_t)PC_BOYS_NM * PC_BOYS_NPOINTS * 4);
  pcb_make_table(PC_BOYS_NM, tab.data());          // csrc/pc_boys_table.h (shared with the _c_ints shim)
  return tab;
}

}  // namespace

// ==========================================================================================
struct pc_basis {
  int device = 0;
  cudaStream_t stream = nullptr;
  int nshell = 0, nbf = 0;
  int ints_type = 0;                    // 0 electron repulsion, 1 scattering (two_electron_scattering.c)
  double grid = -1.0;                   //   ... at this grid value
  int blocks_mode() const { return ints_type == 1 ? PC_MODE_BLOCKS_SCAT : PC_MODE_BLOCKS; }
  int tensor_mode() const { return ints_type == 1 ? PC_MODE_TENSOR_SCAT : PC_MODE_TENSOR; }
  bool cart_d = false;                  // d shells carry their 6 Cartesians (Cartesian_L = [2])
  int nfun(int l) const { return (l == 2 && cart_d) ? 6 : 2 * l + 1; }
  std::vector<Shell> shells;
  std::vector<double> exps, scc;
  std::vector<HostPair> pairs;  // upper-triangular order
  std::vector<double> prim_host; // per pair, most significant primitive first: {sigma, Ucc, Px, Py, Pz, kz}
  std::vector<Kind*> kinds;
  DevBuf<double> boys;                  // [m][j][4]  (one-electron kernel)
  DevBuf<double> boys_l[4 * PCG_LMAX + 1];   // per total angular momentum L < PC_BOYS_COMPACT_MINL: [j][m = 0..L][4]
  DevBuf<double> boys_c[3];             // compact rows of 8 / 12 / 16 values H_k = F_k h^k (L <= 4 / 8 / 12)
  int max_l = 0;                        // highest shell angular momentum of the molecule
  bool force_generic = false;           // PYCHEM_B200_FORCE_GENERIC=1: every class takes the generic kernel (tests)
  DevBuf<double> gen_scratch;           // scratch of the generic kernel for explicit task lists
  std::vector<DevBuf<double>*> gen_bufs;   // ... and of the plan's launch groups
  bool generic(int pcb, int pck) const { return force_generic || pcb >= PC_NPC_GEN || pck >= PC_NPC_GEN; }
  // flat shell table on the device (one-electron integrals)
  DevBuf<int> d_l, d_K, d_poff, d_fn, d_pa, d_pb;
  DevBuf<double> d_A, d_exps, d_scc;
  bool schwarz_done = false;
  std::vector<double> bounds;   // [npair][49]
  // plan
  bool planned = false;
  double thresh = 0;
  int rank = 0, nranks = 1;
  std::vector<PlanItem> plan;
  std::vector<LaunchGroup> groups;     // launch order (longest first)
  // segment tables of ALL plan items, one allocation and one upload each (hundreds of bucket
  // pairs: per-item buffers cost more in cudaMalloc than the plan costs to build)
  DevBuf<long long> plan_seg_off;
  DevBuf<int> plan_seg_ij, plan_warp_s0, plan_seg_rec;
  long long my_quartets = 0, my_eris = 0, all_quartets = 0, all_eris = 0;
  // scratch
  DevBuf<double> acc, dstage, ostage;
  DevBuf<double> bacc, bdstage, bostage;      // batched J/K (nset x 3 N^2 each)
  DevBuf<int> flags;
  long long launches = 0;
  // side streams: the (bra bucket, ket bucket) launches of one Fock build are independent
  // (they only meet in the atomics), so they are spread round-robin to overlap their tails
  std::vector<cudaStream_t> side;
  cudaEvent_t ev_fork = nullptr;
  cudaEvent_t ev_flag = nullptr;        // classification flag has reached flag_host
  int* flag_host = nullptr;             // page-locked
  int guess_variant = 0;                // variant of the previous auto call (speculative digestion)
  std::vector<cudaEvent_t> ev_join;
  // the launch sequence of one Fock build, captured once per (variant, buffers, plan) and
  // replayed as a CUDA graph: removes the host launch cost of ~200 kernels per build
  struct GraphKey {
    int variant = -1;
    const void *dt = nullptr, *da = nullptr, *db = nullptr, *acc = nullptr;
    long long plan_id = -1;
    bool operator==(const GraphKey& o) const {
      return variant == o.variant && dt == o.dt && da == o.da && db == o.db && acc == o.acc && plan_id == o.plan_id;
    }
  };
  GraphKey graph_key;
  cudaGraphExec_t graph_exec = nullptr;
  long long plan_id = 0;
  long long graph_launches = 0;     // kernels inside the cached graph
#ifdef PC_HOST_EMU
  bool use_graphs = false;          // tests/emu: the host emulation cannot replay a captured graph
#else
  bool use_graphs = true;
#endif
  bool profiling = false;
  std::vector<cudaEvent_t> prof_events;     // one before every plan item + one after the last
  std::vector<float> prof_ms;               // per plan item, from the last accumulate

  ~pc_basis() {
    if (graph_exec) cudaGraphExecDestroy(graph_exec);
    for (auto e : prof_events) cudaEventDestroy(e);
    for (auto e : ev_join) cudaEventDestroy(e);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_flag) cudaEventDestroy(ev_flag);
    if (flag_host) cudaFreeHost(flag_host);
    for (auto st : side) cudaStreamDestroy(st);
    for (auto* k : kinds) delete k;
    for (auto* b : gen_bufs) delete b;
    if (stream) cudaStreamDestroy(stream);
  }
  size_t pair_index(int a, int b) const { return (size_t)a * nshell - (size_t)a * (a - 1) / 2 + (b - a); }
};

namespace {

// run fn(0..n-1) on up to 16 host threads
void parallel_for(size_t n, const std::function<void(size_t)>& fn) {
  const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  if (n < 2 || hw == 1) {
    for (size_t k = 0; k < n; ++k) fn(k);
    return;
  }
  std::atomic<size_t> next(0);
  std::vector<std::thread> pool;
  for (unsigned t = 0; t < hw; ++t)
    pool.emplace_back([&]() {
      for (size_t k = next.fetch_add(1); k < n; k = next.fetch_add(1)) fn(k);
    });
  for (auto& th : pool) th.join();
}

// primitive-pair quantities of every shell pair (Methods/c_ints/shellpair_quantities.c:23-36),
// computed once per geometry on the host threads; upload_kind only gathers them in bucket order
void compute_pair_prims(pc_basis* h) {
  size_t off = 0;
  for (HostPair& p : h->pairs) {
    p.prim_off = off;
    off += (size_t)h->shells[p.a].K * h->shells[p.b].K;
  }
  h->prim_host.assign(off * 6, 0.0);
  // uniform normalisation constants folded into the pair prefactor (structures.py:850-856):
  // s: pi^-3/4, p: sqrt(2) pi^-3/4, d: 2 pi^-3/4 (the xy-type d component; xx-type ratio 1/sqrt3
  // lives in the generated cart->spherical code); sqrt(sqrt(2/pi)) per pair gives the
  // sqrt(2/pi) of two_electron_fundamentals.c:24 for the quartet.
  const double pi34 = std::pow(M_PI, -0.75);
  // f: 2 sqrt(2) pi^-3/4 (the xyz component); the per-component ratios are applied by the
  // generic kernel (pcg_norm_ratio)
  const double lnorm[4] = {pi34, std::sqrt(2.0) * pi34, 2.0 * pi34, 2.0 * std::sqrt(2.0) * pi34};
  const double pf_half = std::pow(2.0 / M_PI, 0.25) * std::pow(2.0, 0.25);   // ... and the sqrt(2) of sqrt(2 theta^2)
  const size_t chunk = 256;
  // PYCHEM_B200_PRIM_EPS: measurement knob for the cut-off (default PC_PRIM_EPS)
  const double prim_eps = []() { const char* e = getenv("PYCHEM_B200_PRIM_EPS"); return e ? atof(e) : PC_PRIM_EPS; }();
  parallel_for((h->pairs.size() + chunk - 1) / chunk, [&](size_t c) {
    struct PP { double sigma, ucc, P[3], kz; };
    std::vector<PP> pp;
    for (size_t ip = c * chunk; ip < std::min(h->pairs.size(), (c + 1) * chunk); ++ip) {
      HostPair& p = h->pairs[ip];
      const Shell& X = h->shells[p.x];
      const Shell& Y = h->shells[p.y];
      double r2 = 0;
      for (int c3 = 0; c3 < 3; ++c3) {
        const double d = X.A[c3] - Y.A[c3];
        r2 += d * d;
      }
      const double cn = lnorm[X.l] * lnorm[Y.l] * pf_half;
      // primitive pairs of this shell pair, most significant first.  Pairs whose prefactor
      // U*cc is below PC_PRIM_EPS contribute < 1e-20 to any integral (two-centre pairs of tight
      // primitives: U = exp(-ab/(a+b) r^2) underflows) and are cut off by keff.  The reference
      // visits them all (no primitive screening); the results differ by < 1e-16.
      pp.clear();
      for (int ia = 0; ia < X.K; ++ia)
        for (int ib = 0; ib < Y.K; ++ib) {
          const double a = h->exps[X.poff + ia], b = h->exps[Y.poff + ib];
          PP e;
          e.sigma = 1.0 / (a + b);
          const double U = std::pow(M_PI * e.sigma, 1.5) * std::exp(-a * b * e.sigma * r2);
          e.ucc = U * h->scc[X.poff + ia] * h->scc[Y.poff + ib] * cn;
          for (int c3 = 0; c3 < 3; ++c3) e.P[c3] = (a * X.A[c3] + b * Y.A[c3]) * e.sigma;
          e.kz = b * e.sigma;  // kappa*zeta = (2b)(sigma/2)
          pp.push_back(e);
        }
      std::stable_sort(pp.begin(), pp.end(), [](const PP& u, const PP& v) { return std::fabs(u.ucc) > std::fabs(v.ucc); });
      const int K = (int)pp.size();
      int ke = 0;
      while (ke < K && std::fabs(pp[ke].ucc) >= prim_eps) ++ke;
      p.keff = std::max(ke, 1);
      double* o = &h->prim_host[p.prim_off * 6];
      for (int q = 0; q < K; ++q, o += 6) {
        o[0] = pp[q].sigma; o[1] = pp[q].ucc;
        o[2] = pp[q].P[0];  o[3] = pp[q].P[1];
        o[4] = pp[q].P[2];  o[5] = pp[q].kz;
      }
    }
  });
}

// (re)build and upload the SoA tables of one bucket in its current pair order
int upload_kind(pc_basis* h, Kind* k) {
  const int n = (int)k->pairs.size();
  const int K = k->K;
  std::vector<int> fx(n), fy(n), pid(n), keff(n);
  std::vector<double> xy((size_t)3 * n), prim((size_t)6 * K * n);
  for (int i = 0; i < n; ++i) {
    const HostPair& p = h->pairs[k->pairs[i]];
    const Shell& X = h->shells[p.x];
    const Shell& Y = h->shells[p.y];
    fx[i] = X.first_fn;
    fy[i] = Y.first_fn;
    pid[i] = k->pairs[i];
    for (int c = 0; c < 3; ++c) xy[(size_t)c * n + i] = X.A[c] - Y.A[c];
    keff[i] = p.keff;
    const double* src = &h->prim_host[p.prim_off * 6];
    for (int q = 0; q < K; ++q, src += 6) {
      // [K][3][n] double2: {sigma, U}, {Px, Py}, {Pz, kz}: three 16-byte loads per primitive pair
      double* o0 = &prim[(((size_t)q * 3 + 0) * n + i) * 2];
      double* o1 = &prim[(((size_t)q * 3 + 1) * n + i) * 2];
      double* o2 = &prim[(((size_t)q * 3 + 2) * n + i) * 2];
      o0[0] = src[0]; o0[1] = src[1];
      o1[0] = src[2]; o1[1] = src[3];
      o2[0] = src[4]; o2[1] = src[5];
    }
  }
  // the bra-side records (PcPairKind::rec): header + primitives of one pair, contiguous
  std::vector<double> rec((size_t)6 * (K + 1) * (n + 1), 0.0);
  for (int i = 0; i < n; ++i) {
    const HostPair& p = h->pairs[k->pairs[i]];
    double* o = &rec[(size_t)6 * (K + 1) * i];
    for (int c = 0; c < 3; ++c) o[c] = xy[(size_t)c * n + i];
    o[3] = ((int)k->pm.size() == n) ? k->pm[i] : 0.0;
    int hdr[4] = {fx[i], fy[i], pid[i], keff[i]};
    memcpy(o + 4, hdr, sizeof(hdr));
    memcpy(o + 6, &h->prim_host[p.prim_off * 6], sizeof(double) * 6 * K);
  }
  PC_CUDA(k->rec.upload(rec, h->stream));
  k->keff_h = keff;
  PC_CUDA(k->keff.upload(keff, h->stream));
  PC_CUDA(k->fx.upload(fx, h->stream));
  PC_CUDA(k->fy.upload(fy, h->stream));
  PC_CUDA(k->pid.upload(pid, h->stream));
  PC_CUDA(k->xy.upload(xy, h->stream));
  PC_CUDA(k->prim.upload(prim, h->stream));
  if ((int)k->pm.size() == n && n > 0) PC_CUDA(k->pmd.upload(k->pm, h->stream));
  PC_CUDA(cudaStreamSynchronize(h->stream));  // host vectors go out of scope
  for (int i = 0; i < n; ++i) h->pairs[k->pairs[i]].pos = i;
  return 0;
}

void fill_item(PcItem& I, const Kind* kb, const Kind* kk) {
  memset(&I, 0, sizeof(I));
  I.bra = kb->view();
  I.ket = kk->view();
}

// flop model per primitive / per contracted quartet: the generator's tables for the
// s, p, d classes, the same counting on the loop form for classes with an f shell
double flop_prim(int pcb, int pck) {
  if (pcb < PC_NPC_GEN && pck < PC_NPC_GEN) return pc_flop_prim_table[pcb][pck];
  const int La = pair_L(pcb), Lc = pair_L(pck), L = La + Lc;
  double elems = 0;
  for (int la = 0; la <= La; ++la)
    for (int lc = 0; lc <= Lc; ++lc) elems += (double)ncart(la) * ncart(lc) * (L + 1 - la - lc);
  return 9.0 * (L + 1) + 12.0 + 2.0 * 3.5 * elems;       // ~3.5 base references per VRR element
}
double flop_cont(int pcb, int pck) {
  if (pcb < PC_NPC_GEN && pck < PC_NPC_GEN) return pc_flop_cont_table[pcb][pck];
  int l[4];
  class_l(pcb, l[0], l[1]);
  class_l(pck, l[2], l[3]);
  const double nbc = (double)ncart(l[0]) * ncart(l[1]), nkc = (double)ncart(l[2]) * ncart(l[3]);
  const double ne = pcg_ncum(l[0] + l[1]) - pcg_ncum(l[0] - 1), nf = pcg_ncum(l[2] + l[3]) - pcg_ncum(l[2] - 1);
  // HRR levels (2 flop per element, ~ly levels of ~block size) + the two transforms
  return 2.0 * (ne * nkc * std::max(1, l[3]) + nkc * nbc * std::max(1, l[1])) + 4.0 * nbc * nkc;
}

// scratch layout of the generic kernel for one class (see pcg_quartet); returns words per thread
size_t gen_class_layout(const pc_basis* h, int pcb, int pck, PcGenClass& C) {
  memset(&C, 0, sizeof(C));
  class_l(pcb, C.lx1, C.ly1);
  class_l(pck, C.lx2, C.ly2);
  C.nx1 = h->nfun(C.lx1); C.ny1 = h->nfun(C.ly1); C.nx2 = h->nfun(C.lx2); C.ny2 = h->nfun(C.ly2);
  C.cart_d = h->cart_d ? 1 : 0;
  C.scat = h->ints_type == 1 ? 1 : 0;
  const int La = C.lx1 + C.ly1, Lc = C.lx2 + C.ly2;
  C.L = La + Lc;
  int off = 0;
  for (int la = 0; la <= La; ++la)
    for (int lc = 0; lc <= Lc; ++lc) {
      C.offV[la][lc] = off;
      off += ncart(la) * ncart(lc) * (C.L + 1 - la - lc);
    }
  const int ne = pcg_ncum(La) - pcg_ncum(C.lx1 - 1), nf = pcg_ncum(Lc) - pcg_ncum(C.lx2 - 1);
  const int nbc = ncart(C.lx1) * ncart(C.ly1), nkc = ncart(C.lx2) * ncart(C.ly2);
  const int nsb = C.nx1 * C.ny1, nsk = C.nx2 * C.ny2;
  C.offT1 = 0;
  C.offG = ne * nkc;
  C.sizeA = std::max(off, std::max(ne * nkc + nbc * nkc, nsb * nsk));
  C.sizeB = std::max(ne * nf, nbc * nsk);
  return (size_t)C.sizeA + C.sizeB;
}

// threads of a generic launch: whole 64-thread blocks, enough for the tasks, bounded by the
// scratch budget (PYCHEM_B200_GENERIC_SCRATCH_MB per launch, default 512)
int gen_thread_count(long long nwarps, size_t words) {
  size_t mb = 512;
  if (const char* e = getenv("PYCHEM_B200_GENERIC_SCRATCH_MB"))
    if (atoll(e) > 0) mb = (size_t)atoll(e);
  long long cap = (long long)((mb << 20) / (words * sizeof(double)));
  cap = std::max<long long>(64, cap / 64 * 64);
  cap = std::min<long long>(cap, 148LL * 16 * 64);            // 16 blocks of 64 per SM
  const long long want = (nwarps * 32 + 63) / 64 * 64;
  return (int)std::max<long long>(64, std::min(cap, want));
}

// `scratch`/`gen_threads`: the generic kernel's scratch (classes with an f shell); null = the
// handle's own buffer, grown on demand (explicit task lists, all on h->stream)
cudaError_t launch_args(pc_basis* h, int mode, int pcb, int pck, PcEriArgs& A, cudaStream_t st,
                        DevBuf<double>* scratch = nullptr, int gen_threads = 0) {
  {
    const int Lq = pair_L(pcb) + pair_L(pck);
    A.boys = Lq < PC_BOYS_COMPACT_MINL ? h->boys_l[Lq].p : h->boys_c[pc_boys_row(Lq) / 4 - 2].p;
  }
  A.nbf = h->nbf;
  A.scat_S = h->grid;
  A.thresh = h->thresh;
  cudaError_t e;
  if (h->generic(pcb, pck)) {
    PcGenClass C;
    const size_t words = gen_class_layout(h, pcb, pck, C);
    if (!scratch) {
      scratch = &h->gen_scratch;
      gen_threads = gen_thread_count(A.nwarps, words);
      if (scratch->n < words * (size_t)gen_threads) {
        e = cudaStreamSynchronize(h->stream);                  // an earlier launch may still use it
        if (e == cudaSuccess) e = scratch->alloc(words * (size_t)gen_threads);
        if (e != cudaSuccess) return e;
      }
    }
    C.scratch = scratch->p;
    C.nthreads = gen_threads;
    e = pc_launch_generic(mode, A, C, st ? st : h->stream);
  } else {
    pc_launch_fn fn = pc_launch_table[h->cart_d ? 1 : 0][pcb][pck];
    if (!fn) return cudaErrorInvalidValue;
    e = fn(mode, A, st ? st : h->stream);
  }
  if (e == cudaSuccess) h->launches += 1;
  return e;
}

// single bucket pair with an explicit (bra, ket) task list -- Schwarz diagonal, pc_eri_quartets
int launch_explicit(pc_basis* h, int mode, const Kind* kb, const Kind* kk, PcEriArgs& A, const int* ex_bra,
                    const int* ex_ket, long long count) {
  fill_item(A.items[0], kb, kk);
  A.items[0].t_count = count;
  A.items[0].warp0 = 0;
  A.nitems = 1;
  A.nwarps = (int)((count + 31) / 32);
  A.ex_bra = ex_bra;
  A.ex_ket = ex_ket;
  cudaError_t e = launch_args(h, mode, kb->pc, kk->pc, A, nullptr);
  if (e != cudaSuccess) return fail(std::string("kernel launch: ") + cudaGetErrorString(e));
  return 0;
}

// one fused launch: all bucket pairs of one class (a LaunchGroup), every item warp-aligned
int launch_group(pc_basis* h, int mode, const LaunchGroup& g, PcEriArgs& A, cudaStream_t st) {
  int warp = 0, n = 0;
  for (int idx : g.items) {
    const PlanItem& it = h->plan[idx];
    if (it.count == 0) continue;
    PcItem& I = A.items[n++];
    fill_item(I, h->kinds[it.kb], h->kinds[it.kk]);
    I.seg_off = it.seg_off; I.seg_ij = (const int2*)it.seg_ij; I.warp_s0 = it.warp_s0;
    I.seg_rec = (const int4*)it.seg_rec;
    I.nseg = it.nseg; I.t_begin = it.begin; I.t_count = it.count;
    I.same = it.same;
    I.warp0 = warp;
    warp += (int)((it.count + 31) / 32);
  }
  if (n == 0) return 0;
  A.nitems = n;
  A.nwarps = warp;
  A.ex_bra = nullptr;
  A.ex_ket = nullptr;
  cudaError_t e = launch_args(h, mode, g.pcb, g.pck, A, st, g.scratch, g.gen_threads);
  if (e != cudaSuccess) return fail(std::string("kernel launch: ") + cudaGetErrorString(e));
  return 0;
}

// The segments of one (bra bucket, ket bucket): pure host code (also exported for the CPU tests
// as pc_plan_segments_host).  pm: Schwarz maximum by position, gstart: group starts, keff:
// significant primitive pairs by position.
void build_segments_host(const std::vector<double>& pmB, const std::vector<int>& gsB, const std::vector<int>& keffB,
                         const std::vector<double>& pmK, const std::vector<int>& gsK, const std::vector<int>& keffK,
                         int same, int run, double thresh, std::vector<long long>& seg_off,
                         std::vector<long long>& seg_q, std::vector<double>& seg_prim, std::vector<int>& ij) {
  const int ng = (int)gsK.size() - 1;
  const int nbg = (int)gsB.size() - 1;
  // One segment = (a RUN of up to `run` consecutive bra pairs of one bra group) x (a prefix of one
  // ket group).  The test is the reference's: max(B_ab)*max(B_cd) > thresh, strict
  // (hartree_fock.py:293-294); unique quartets only (ket position >= bra position inside one
  // bucket); the diagonal (ab|ab) is always kept (hartree_fock.py:244-250).  Inside a group the
  // pairs are ordered by descending Schwarz maximum, so the kets that survive bra pair i0+1 are a
  // prefix of those that survive i0, and for one ket the surviving bra pairs are a prefix of the
  // run: the segment is sized by the run's first bra pair and every thread re-tests the (exactly
  // reproducible) product for the later ones.
  std::vector<double> pre(keffK.size() + 1, 0.0);
  for (size_t q = 0; q < keffK.size(); ++q) pre[q + 1] = pre[q] + keffK[q];
  seg_off.assign(1, 0);
  seg_q.assign(1, 0);
  seg_prim.assign(1, 0.0);
  ij.clear();
  auto emit = [&](int i0, int r, int forced, int j0, int len, long long nq, double nprim) {
    ij.push_back(i0 | (r << 24) | (forced << 28));
    ij.push_back(j0);
    seg_off.push_back(seg_off.back() + len);
    seg_q.push_back(seg_q.back() + nq);
    seg_prim.push_back(seg_prim.back() + nprim);
  };
  int j1[PC_MAX_RUN + 1], j0[PC_MAX_RUN + 1];
  bool diag[PC_MAX_RUN + 1];
  run = std::max(1, std::min(run, PC_MAX_RUN));
  for (int bg = 0; bg < nbg; ++bg)
    for (int i0 = gsB[bg]; i0 < gsB[bg + 1]; i0 += run) {
      const int rr = std::min(run, gsB[bg + 1] - i0);
      for (int q = 0; q < rr; ++q) diag[q] = !same;
      for (int g = 0; g < ng; ++g) {
        const int gs = gsK[g], ge = gsK[g + 1];
        if (!(pmB[i0] * pmK[gs] > thresh)) break;     // groups are ordered by their maximum
        int r_eff = 0;
        long long nq = 0;
        double nprim = 0;
        for (int q = 0; q < rr; ++q) {
          const double pb = pmB[i0 + q];
          j0[q] = same ? std::max(gs, i0 + q) : gs;
          j1[q] = gs;
          if (pb * pmK[gs] > thresh) {
            int lo = gs, hi = (q == 0) ? ge : j1[q - 1];     // first position that fails the test
            if (hi > lo) {
              while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (pb * pmK[mid] > thresh) lo = mid; else hi = mid;
              }
              j1[q] = hi;
            }
          }
          if (j1[q] <= j0[q]) break;                          // ... and so do all later bra pairs
          r_eff = q + 1;
          nq += j1[q] - j0[q];
          nprim += (double)keffB[i0 + q] * (pre[j1[q]] - pre[j0[q]]);
          if (same && i0 + q >= j0[q] && i0 + q < j1[q]) diag[q] = true;
        }
        if (r_eff > 0) emit(i0, r_eff, 0, j0[0], j1[0] - j0[0], nq, nprim);
      }
      for (int q = 0; q < rr; ++q)
        if (!diag[q]) emit(i0 + q, 1, 1, i0 + q, 1, 1, (double)keffB[i0 + q] * keffK[i0 + q]);
    }
}

// Longest bra run of a class: the default of its run kernel (pc_run_table; 1 = the class has the
// one-quartet-per-thread kernel), optionally changed at run time for the run kernels:
// PYCHEM_B200_RUN="4" (all of them) or "psss=4,ssss=8,..." (per class; classes not named keep
// their default).
int run_length(const pc_basis* h, int pcb, int pck) {
  if (h->generic(pcb, pck)) return 1;            // the generic kernel takes one quartet per thread
  int r = pc_run_table[h->cart_d ? 1 : 0][pcb][pck];
  if (r <= 1) return 1;                          // not a run kernel
  const char* e = getenv("PYCHEM_B200_RUN");
  if (e && *e) {
    static const char* pcname[6] = {"ss", "ps", "pp", "ds", "dp", "dd"};
    const std::string spec(e);
    if (spec.find('=') == std::string::npos) {
      if (atoi(e) >= 1) r = atoi(e);
    } else {
      const std::string key = std::string(pcname[pcb]) + pcname[pck] + "=";
      size_t pos = 0;
      while ((pos = spec.find(key, pos)) != std::string::npos) {
        if (pos == 0 || spec[pos - 1] == ',') {
          const int v = atoi(spec.c_str() + pos + key.size());
          if (v >= 1) r = v;
          break;
        }
        pos += key.size();
      }
    }
  }
  return std::max(1, std::min(r, PC_MAX_RUN));
}

// copy a host-or-device N*N matrix into device staging (returns device pointer)
int stage_in(pc_basis* h, const double* src, double* stage, const double** out) {
  if (src == stage || is_device_ptr(src)) {
    *out = src;
    return 0;
  }
  PC_CUDA(cudaMemcpyAsync(stage, src, sizeof(double) * h->nbf * h->nbf, cudaMemcpyHostToDevice, h->stream));
  *out = stage;
  return 0;
}

}  // namespace

// ==========================================================================================
extern "C" {

const char* pc_last_error(void) { return g_err.c_str(); }

int pc_device_count(int* count) {
  int n = 0;
  PC_CUDA(cudaGetDeviceCount(&n));
  if (n <= 0) return fail("no CUDA device visible");
  if (count) *count = n;
  return 0;
}

int pc_basis_create(int device, int nshell, const int* l, const int* K, const int* is_cart,
                    const int* first_fn, const double* centres, const double* exps,
                    const double* scc, pc_basis** out) {
  if (!out || nshell <= 0) return fail("pc_basis_create: bad arguments");
  PC_CUDA(cudaSetDevice(device));
  pc_basis* h = new pc_basis();
  h->device = device;
  cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete h; return fail(cudaGetErrorString(e)); }
  h->nshell = nshell;
  if (const char* fg = getenv("PYCHEM_B200_FORCE_GENERIC")) h->force_generic = atoi(fg) != 0;
  int poff = 0, n_d = 0;
  for (int s = 0; s < nshell; ++s) {
    if (l[s] < 0 || l[s] > PCG_LMAX) { delete h; return fail("pc_basis_create: only s, p, d, f shells are supported by this build"); }
    if (l[s] == 3 && is_cart[s]) { delete h; return fail("pc_basis_create: Cartesian f shells are not supported (Cartesian_L = [2] only)"); }
    h->max_l = std::max(h->max_l, l[s]);
    if (l[s] == 2) {
      // Cartesian_L = [2] (Util/structures.py:844-849) applies to every d shell of the molecule
      if (n_d == 0) h->cart_d = is_cart[s] != 0;
      else if (h->cart_d != (is_cart[s] != 0)) { delete h; return fail("pc_basis_create: mixed Cartesian and spherical d shells are not supported"); }
      ++n_d;
    }
    if (K[s] <= 0) { delete h; return fail("pc_basis_create: empty contraction"); }
    Shell sh;
    sh.l = l[s]; sh.K = K[s]; sh.first_fn = first_fn[s]; sh.nfn = (l[s] == 2 && is_cart[s]) ? 6 : 2 * l[s] + 1; sh.poff = poff;
    for (int c = 0; c < 3; ++c) sh.A[c] = centres[3 * s + c];
    poff += K[s];
    h->shells.push_back(sh);
    h->nbf = std::max(h->nbf, sh.first_fn + sh.nfn);
  }
  h->exps.assign(exps, exps + poff);
  h->scc.assign(scc, scc + poff);
  // shell pairs, bucketed by (lx, ly, Kx*Ky)
  std::map<std::tuple<int, int, int>, int> kind_of;
  h->pairs.reserve((size_t)nshell * (nshell + 1) / 2);
  for (int a = 0; a < nshell; ++a)
    for (int b = a; b < nshell; ++b) {
      HostPair p;
      p.a = a; p.b = b;
      p.swapped = h->shells[b].l > h->shells[a].l;   // "Goofy" pairs: integrals.py:79-89
      p.x = p.swapped ? b : a;
      p.y = p.swapped ? a : b;
      p.pmax = 0;
      const int lx = h->shells[p.x].l, ly = h->shells[p.y].l, KK = h->shells[a].K * h->shells[b].K;
      auto key = std::make_tuple(lx, ly, KK);
      auto it = kind_of.find(key);
      if (it == kind_of.end()) {
        Kind* k = new Kind();
        k->lx = lx; k->ly = ly; k->K = KK; k->pc = pair_class(lx, ly);
        it = kind_of.emplace(key, (int)h->kinds.size()).first;
        h->kinds.push_back(k);
      }
      p.kind = it->second;
      p.pos = (int)h->kinds[p.kind]->pairs.size();
      h->kinds[p.kind]->pairs.push_back((int)h->pairs.size());
      h->pairs.push_back(p);
    }
  compute_pair_prims(h);
  for (Kind* k : h->kinds)
    if (upload_kind(h, k)) { delete h; return 1; }
  {
    std::vector<int> vl, vK, vpo, vfn, vpa, vpb;
    std::vector<double> vA;
    for (const Shell& sh : h->shells) {
      vl.push_back(sh.l); vK.push_back(sh.K); vpo.push_back(sh.poff); vfn.push_back(sh.first_fn);
      for (int c = 0; c < 3; ++c) vA.push_back(sh.A[c]);
    }
    for (const HostPair& p : h->pairs) { vpa.push_back(p.a); vpb.push_back(p.b); }
    cudaError_t e2 = h->d_l.upload(vl, h->stream);
    if (e2 == cudaSuccess) e2 = h->d_K.upload(vK, h->stream);
    if (e2 == cudaSuccess) e2 = h->d_poff.upload(vpo, h->stream);
    if (e2 == cudaSuccess) e2 = h->d_fn.upload(vfn, h->stream);
    if (e2 == cudaSuccess) e2 = h->d_pa.upload(vpa, h->stream);
    if (e2 == cudaSuccess) e2 = h->d_pb.upload(vpb, h->stream);
    if (e2 == cudaSuccess) e2 = h->d_A.upload(vA, h->stream);
    if (e2 == cudaSuccess) e2 = h->d_exps.upload(h->exps, h->stream);
    if (e2 == cudaSuccess) e2 = h->d_scc.upload(h->scc, h->stream);
    if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(h->stream);
    if (e2 != cudaSuccess) { delete h; return fail(cudaGetErrorString(e2)); }
  }
  std::vector<double> tab = make_boys_table();
  e = h->boys.upload(tab, h->stream);
  // The ERI kernels of one class read the orders m = 0..L of ONE interval: per-L copies with the
  // orders of an interval adjacent ((L+1) x 32 bytes contiguous) cost one cache line per lookup
  // instead of L+1 lines 248 KB apart.
  for (int L = 0; L <= 4 * h->max_l && L < PC_BOYS_COMPACT_MINL && e == cudaSuccess; ++L) {
    std::vector<double> t((size_t)PC_BOYS_NPOINTS * (L + 1) * 4);
    for (int j = 0; j < PC_BOYS_NPOINTS; ++j)
      for (int m = 0; m <= L; ++m)
        for (int c = 0; c < 4; ++c)
          t[((size_t)j * (L + 1) + m) * 4 + c] = tab[((size_t)m * PC_BOYS_NPOINTS + j) * 4 + c];
    e = h->boys_l[L].upload(t, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);     // `t` dies here
  }
  // From L = PC_BOYS_COMPACT_MINL on the classes share compact rows: L + 4 scaled values per grid
  // point (one 496 KB table for L <= 4, 744 KB for L <= 8, 992 KB with f shells) instead of per-L
  // copies of 4 (L + 1) coefficients (744 KB ... 3.2 MB each): 2 / 3 / 4 sectors per lookup.
  for (int t = 0; t < 3 && e == cudaSuccess; ++t) {
    const int row = 8 + 4 * t, lmax = 4 * (t + 1);
    if (4 * h->max_l <= lmax - 4 && t > 0) break;                   // no class needs the longer rows
    if (4 * h->max_l < PC_BOYS_COMPACT_MINL) break;
    std::vector<double> v((size_t)PC_BOYS_NPOINTS * row);
    pcb_make_scaled_values(row, v.data());
    e = h->boys_c[t].upload(v, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);     // `v` dies here
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) { delete h; return fail(cudaGetErrorString(e)); }
  *out = h;
  return 0;
}

int pc_basis_set_ints_type(pc_basis* h, int ints_type, double grid_value) {
  if (!h) return fail("pc_basis_set_ints_type: null");
  if (ints_type != 0 && ints_type != 1) return fail("pc_basis_set_ints_type: ints_type must be 0 or 1");
  if (h->ints_type != ints_type || (ints_type == 1 && h->grid != grid_value)) {
    h->ints_type = ints_type;
    h->grid = grid_value;
    h->schwarz_done = false;       // bounds, pair order and plan belong to the integral type
    h->planned = false;
  }
  return 0;
}

int pc_basis_destroy(pc_basis* h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  delete h;
  return 0;
}

int pc_basis_nbf(const pc_basis* h, int* nbf) {
  if (!h || !nbf) return fail("pc_basis_nbf: null");
  *nbf = h->nbf;
  return 0;
}

int pc_basis_stream(const pc_basis* h, void** stream) {
  if (!h || !stream) return fail("pc_basis_stream: null");
  *stream = (void*)h->stream;
  return 0;
}

int pc_launch_count(const pc_basis* h, long long* n) {
  if (!h || !n) return fail("pc_launch_count: null");
  *n = h->launches;
  return 0;
}

// ------------------------------------------------------------------------------------------
int pc_schwarz(pc_basis* h, double* bounds, double* pmax) {
  if (!h) return fail("pc_schwarz: null handle");
  PC_CUDA(cudaSetDevice(h->device));
  const size_t npair = h->pairs.size();
  if (!h->schwarz_done) {
    h->bounds.assign(npair * 49, 0.0);
    for (Kind* k : h->kinds) {
      const int n = (int)k->pairs.size();
      const int nx = h->nfun(k->lx), ny = h->nfun(k->ly), nb = nx * ny;
      std::vector<int> idx(n);
      std::iota(idx.begin(), idx.end(), 0);
      DevBuf<int> didx;
      DevBuf<double> dout;
      PC_CUDA(didx.upload(idx, h->stream));
      PC_CUDA(dout.alloc((size_t)nb * nb * n));
      PcEriArgs A;
      memset(&A, 0, sizeof(A));
      A.out = dout.p;
      if (launch_explicit(h, h->blocks_mode(), k, k, A, didx.p, didx.p, n)) return 1;
      std::vector<double> out((size_t)nb * nb * n);
      PC_CUDA(cudaMemcpyAsync(out.data(), dout.p, out.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
      PC_CUDA(cudaStreamSynchronize(h->stream));
      for (int t = 0; t < n; ++t) {
        HostPair& p = h->pairs[k->pairs[t]];
        const int nfa = h->shells[p.a].nfn, nfb = h->shells[p.b].nfn;
        double mx = 0;
        for (int mxi = 0; mxi < nx; ++mxi)
          for (int myi = 0; myi < ny; ++myi) {
            const int q = mxi * ny + myi;
            // numpy.sqrt of the diagonal (mn|mn), hartree_fock.py:251-254
            // scattering diagonals can dip below zero: numpy.sqrt gives nan there and the block
            // is skipped by the screen; 0 has the same effect
            const double dv = out[((size_t)q * nb + q) * n + t];
            const double v = dv > 0.0 ? std::sqrt(dv) : 0.0;
            const int ma = p.swapped ? myi : mxi, mb = p.swapped ? mxi : myi;
            h->bounds[h->pair_index(p.a, p.b) * 49 + ma * nfb + mb] = v;
            (void)nfa;
            mx = std::max(mx, v);
          }
        p.pmax = mx;
      }
    }
    // order every bucket in groups of pairs sharing the primary shell x (groups by descending
    // group maximum, pairs inside a group by descending Schwarz maximum) and rebuild its tables.
    // Consecutive tasks of one bra pair then share the ket's primary shell c, so the lanes of a
    // warp hit the same K[a,c] / K[b,c] / J[a,b] addresses and can be reduced by shuffles; inside
    // a group the survivors of a bra pair are still a prefix.
    for (Kind* k : h->kinds) {
      std::map<int, double> gmax;
      for (int p : k->pairs) {
        double& g = gmax[h->pairs[p].x];
        g = std::max(g, h->pairs[p].pmax);
      }
      std::stable_sort(k->pairs.begin(), k->pairs.end(), [&](int p, int q) {
        const HostPair& P = h->pairs[p];
        const HostPair& Q = h->pairs[q];
        if (P.x != Q.x) {
          const double gp = gmax[P.x], gq = gmax[Q.x];
          if (gp != gq) return gp > gq;
          return P.x < Q.x;
        }
        return P.pmax > Q.pmax;
      });
      k->gstart.clear();
      k->pm.resize(k->pairs.size());
      for (size_t i = 0; i < k->pairs.size(); ++i) {
        k->pm[i] = h->pairs[k->pairs[i]].pmax;
        if (i == 0 || h->pairs[k->pairs[i]].x != h->pairs[k->pairs[i - 1]].x) k->gstart.push_back((int)i);
      }
      k->gstart.push_back((int)k->pairs.size());
      if (upload_kind(h, k)) return 1;
    }
    h->schwarz_done = true;
    h->planned = false;
  }
  if (bounds) memcpy(bounds, h->bounds.data(), sizeof(double) * npair * 49);
  if (pmax)
    for (size_t p = 0; p < npair; ++p) pmax[p] = h->pairs[p].pmax;
  return 0;
}

// ------------------------------------------------------------------------------------------
int pc_plan(pc_basis* h, double thresh, int rank, int nranks, long long* my_quartets,
            long long* my_eris, long long* all_quartets, long long* all_eris) {
  if (!h) return fail("pc_plan: null handle");
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail("pc_plan: bad rank/nranks");
  if (!h->schwarz_done && pc_schwarz(h, nullptr, nullptr)) return 1;
  PC_CUDA(cudaSetDevice(h->device));
  if (!(h->planned && h->thresh == thresh && h->rank == rank && h->nranks == nranks)) {
    for (auto* b : h->gen_bufs) delete b;
    h->gen_bufs.clear();
    h->plan.clear();
    h->my_quartets = h->my_eris = h->all_quartets = h->all_eris = 0;
    // ---- phase A (host threads): the segments of every (bra bucket, ket bucket) ------------
    struct Work {
      int kb, kk, same, run;
      std::vector<long long> seg_off, seg_q;     // prefix of tasks / of quartets per segment
      std::vector<int> ij, s0;
      std::vector<double> seg_prim;              // prefix of primitive quartets per segment
    };
    std::vector<Work> work;
    const int nk = (int)h->kinds.size();
    for (const Kind* k : h->kinds)
      if (k->pairs.size() > (size_t)PC_SEG_I_MASK) return fail("pc_plan: more than 2^24 shell pairs in one bucket");
    for (int ka = 0; ka < nk; ++ka)
      for (int kb2 = ka; kb2 < nk; ++kb2) {
        Work w;
        w.kb = ka; w.kk = kb2;
        if (h->kinds[w.kb]->pc < h->kinds[w.kk]->pc) std::swap(w.kb, w.kk);
        w.same = (w.kb == w.kk);
        w.run = run_length(h, h->kinds[w.kb]->pc, h->kinds[w.kk]->pc);
        work.push_back(std::move(w));
      }
    auto build_segments = [&](Work& w) {
      const Kind* B = h->kinds[w.kb];
      const Kind* Kt = h->kinds[w.kk];
      build_segments_host(B->pm, B->gstart, B->keff_h, Kt->pm, Kt->gstart, Kt->keff_h, w.same, w.run, thresh,
                          w.seg_off, w.seg_q, w.seg_prim, w.ij);
    };
    parallel_for(work.size(), [&](size_t k) { build_segments(work[k]); });
    // ---- phase B: plan items + static multi-GPU schedule ---------------------
    std::vector<int> widx;       // work index of every plan item
    for (size_t k = 0; k < work.size(); ++k) {
      Work& w = work[k];
      if (w.seg_off.back() == 0) continue;
      PlanItem it;
      it.kb = w.kb; it.kk = w.kk; it.same = w.same;
      it.total = w.seg_off.back();
      it.begin = 0; it.count = it.total;
      it.total_q = w.seg_q.back(); it.count_q = it.total_q;
      it.prim_exec = w.seg_prim.back();
      it.nseg = (int)w.ij.size() / 2;
      it.seg_off = nullptr; it.seg_ij = nullptr; it.warp_s0 = nullptr; it.seg_rec = nullptr;
      h->plan.push_back(it);
      widx.push_back((int)k);
    }
    auto quartet_cost = [&](const PlanItem& it) {
      const Kind* B = h->kinds[it.kb];
      const Kind* Kt = h->kinds[it.kk];
      const double nsph = (double)h->nfun(B->lx) * h->nfun(B->ly) * h->nfun(Kt->lx) * h->nfun(Kt->ly);
      return flop_cont(B->pc, Kt->pc) + 40.0 * nsph + 60.0;
    };
    auto cost_total = [&](const PlanItem& it) {
      return it.prim_exec * flop_prim(h->kinds[it.kb]->pc, h->kinds[it.kk]->pc) +
             (double)it.total_q * quartet_cost(it);
    };
    // every bucket pair is cut into nranks contiguous slices of equal modelled cost, at segment
    // boundaries (the cost per quartet inside a bucket pair is nearly uniform, so the static
    // schedule is balanced to a fraction of a percent); because all bucket pairs of a class share
    // ONE fused launch, the slices of small bucket pairs cost nothing extra
    for (size_t k = 0; k < h->plan.size(); ++k) {
      PlanItem& it = h->plan[k];
      const Work& w = work[widx[k]];
      const double fp = flop_prim(h->kinds[it.kb]->pc, h->kinds[it.kk]->pc), fq = quartet_cost(it);
      auto seg_cost = [&](int sg) { return w.seg_prim[sg] * fp + (double)w.seg_q[sg] * fq; };
      auto cut = [&](int r) {           // first segment of rank r's slice
        if (r <= 0) return 0;
        if (r >= nranks) return it.nseg;
        const double target = seg_cost(it.nseg) * r / nranks;
        int lo = 0, hi = it.nseg;     // smallest sg with seg_cost(sg) >= target
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (seg_cost(mid) < target) lo = mid + 1; else hi = mid;
        }
        return lo;
      };
      const int s_begin = cut(rank), s_end = cut(rank + 1);
      it.begin = w.seg_off[s_begin];
      it.count = w.seg_off[s_end] - it.begin;
      it.count_q = w.seg_q[s_end] - w.seg_q[s_begin];
    }
    // ---- phase C (host threads): segment of every warp's first task ------------------------
    parallel_for(h->plan.size(), [&](size_t k) {
      const PlanItem& it = h->plan[k];
      Work& w = work[widx[k]];
      const long long nwarp = (it.count + 31) / 32;
      w.s0.resize((size_t)nwarp);
      int cur = 0;
      for (long long wi = 0; wi < nwarp; ++wi) {
        const long long g0 = it.begin + wi * 32;
        while (w.seg_off[cur + 1] <= g0) ++cur;
        w.s0[(size_t)wi] = cur;
      }
    });
    // ---- phase D: uploads into the three pooled buffers (one synchronisation at the end) ----
    {
      size_t n_off = 0, n_ij = 0, n_s0 = 0;
      for (size_t k = 0; k < h->plan.size(); ++k) {
        if (h->plan[k].count == 0) continue;
        const Work& w = work[widx[k]];
        n_off += w.seg_off.size(); n_ij += w.ij.size(); n_s0 += w.s0.size();
      }
      PC_CUDA(h->plan_seg_off.alloc(n_off));
      PC_CUDA(h->plan_seg_ij.alloc(n_ij));
      PC_CUDA(h->plan_warp_s0.alloc(n_s0));
      PC_CUDA(h->plan_seg_rec.alloc(4 * n_off));
    }
    size_t o_off = 0, o_ij = 0, o_s0 = 0;       // w.ij holds int2 records: every item starts 8-byte aligned
    std::vector<std::vector<int>> recs(h->plan.size());      // alive until the stream is synchronised below
    for (size_t k = 0; k < h->plan.size(); ++k) {
      PlanItem& it = h->plan[k];
      const Work& w = work[widx[k]];
      const Kind* B = h->kinds[it.kb];
      const Kind* Kt = h->kinds[it.kk];
      if (it.count > 0) {
        it.seg_off = h->plan_seg_off.p + o_off;
        it.seg_ij = h->plan_seg_ij.p + o_ij;
        it.warp_s0 = h->plan_warp_s0.p + o_s0;
        PC_CUDA(cudaMemcpyAsync(h->plan_seg_off.p + o_off, w.seg_off.data(), w.seg_off.size() * sizeof(long long), cudaMemcpyHostToDevice, h->stream));
        PC_CUDA(cudaMemcpyAsync(h->plan_seg_ij.p + o_ij, w.ij.data(), w.ij.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
        PC_CUDA(cudaMemcpyAsync(h->plan_warp_s0.p + o_s0, w.s0.data(), w.s0.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
        // merged decode records (one per seg_off entry, the sentinel included)
        std::vector<int>& rec = recs[k];
        rec.resize(4 * w.seg_off.size());
        for (size_t sg = 0; sg < w.seg_off.size(); ++sg) {
          const unsigned long long off = (unsigned long long)w.seg_off[sg];
          rec[4 * sg] = (int)(unsigned)(off & 0xffffffffu);
          rec[4 * sg + 1] = (int)(unsigned)(off >> 32);
          rec[4 * sg + 2] = sg < (size_t)it.nseg ? w.ij[2 * sg] : 0;
          rec[4 * sg + 3] = sg < (size_t)it.nseg ? w.ij[2 * sg + 1] : 0;
        }
        it.seg_rec = h->plan_seg_rec.p + 4 * o_off;
        PC_CUDA(cudaMemcpyAsync(h->plan_seg_rec.p + 4 * o_off, rec.data(), rec.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
        o_off += w.seg_off.size(); o_ij += w.ij.size(); o_s0 += w.s0.size();
      }
      const long long nsph = (long long)h->nfun(B->lx) * h->nfun(B->ly) * h->nfun(Kt->lx) * h->nfun(Kt->ly);
      h->all_quartets += it.total_q; h->all_eris += it.total_q * nsph;
      h->my_quartets += it.count_q; h->my_eris += it.count_q * nsph;
    }
    PC_CUDA(cudaStreamSynchronize(h->stream));      // host vectors in `work` die below
    // ---- launch groups: all bucket pairs of one class -> one fused launch, longest first ----
    h->groups.clear();
    {
      std::map<std::pair<int, int>, std::vector<int>> by_class;
      for (size_t k = 0; k < h->plan.size(); ++k)
        by_class[{h->kinds[h->plan[k].kb]->pc, h->kinds[h->plan[k].kk]->pc}].push_back((int)k);
      for (auto& kv : by_class) {
        std::vector<int>& v = kv.second;
        std::stable_sort(v.begin(), v.end(), [&](int a, int b) { return cost_total(h->plan[a]) > cost_total(h->plan[b]); });
        for (size_t c0 = 0; c0 < v.size(); c0 += PC_MAX_ITEMS) {
          LaunchGroup g;
          g.pcb = kv.first.first; g.pck = kv.first.second; g.cost = 0;
          for (size_t c = c0; c < std::min(v.size(), c0 + PC_MAX_ITEMS); ++c) {
            g.items.push_back(v[c]);
            g.cost += cost_total(h->plan[v[c]]);
          }
          h->groups.push_back(g);
        }
      }
      std::stable_sort(h->groups.begin(), h->groups.end(),
                       [](const LaunchGroup& a, const LaunchGroup& b) { return a.cost > b.cost; });
      // classes with an f shell: every launch group owns the scratch of its generic kernel
      for (LaunchGroup& g : h->groups) {
        if (!h->generic(g.pcb, g.pck)) continue;
        long long nwarps = 0;
        for (int idx : g.items) nwarps += (h->plan[idx].count + 31) / 32;
        if (nwarps == 0) continue;
        PcGenClass C;
        const size_t words = gen_class_layout(h, g.pcb, g.pck, C);
        g.gen_threads = gen_thread_count(nwarps, words);
        g.scratch = new DevBuf<double>();
        h->gen_bufs.push_back(g.scratch);
        PC_CUDA(g.scratch->alloc(words * (size_t)g.gen_threads));
      }
    }
    h->planned = true;
    h->plan_id += 1;
    h->thresh = thresh; h->rank = rank; h->nranks = nranks;
  }
  if (my_quartets) *my_quartets = h->my_quartets;
  if (my_eris) *my_eris = h->my_eris;
  if (all_quartets) *all_quartets = h->all_quartets;
  if (all_eris) *all_eris = h->all_eris;
  return 0;
}

// ------------------------------------------------------------------------------------------
int pc_eri_quartets(pc_basis* h, int n, const int* abcd, const long long* offsets, double* out) {
  if (!h || n < 0 || (n && (!abcd || !offsets || !out))) return fail("pc_eri_quartets: bad arguments");
  PC_CUDA(cudaSetDevice(h->device));
  // group by (bra kind, ket kind) in class-canonical order
  struct Grp { std::vector<int> q, bi, kj, flip; };
  std::map<std::pair<int, int>, Grp> groups;
  for (int q = 0; q < n; ++q) {
    const int a = abcd[4 * q], b = abcd[4 * q + 1], c = abcd[4 * q + 2], d = abcd[4 * q + 3];
    if (a < 0 || b < a || b >= h->nshell || c < 0 || d < c || d >= h->nshell)
      return fail("pc_eri_quartets: need 0 <= a <= b < nshell and 0 <= c <= d < nshell");
    const HostPair& P = h->pairs[h->pair_index(a, b)];
    const HostPair& Q = h->pairs[h->pair_index(c, d)];
    int flip = h->kinds[P.kind]->pc < h->kinds[Q.kind]->pc;
    const HostPair& B = flip ? Q : P;
    const HostPair& Kt = flip ? P : Q;
    Grp& g = groups[{B.kind, Kt.kind}];
    g.q.push_back(q); g.bi.push_back(B.pos); g.kj.push_back(Kt.pos); g.flip.push_back(flip);
  }
  for (auto& kv : groups) {
    const Kind* B = h->kinds[kv.first.first];
    const Kind* Kt = h->kinds[kv.first.second];
    Grp& g = kv.second;
    const int m = (int)g.q.size();
    const int n1 = h->nfun(B->lx), n2 = h->nfun(B->ly), n3 = h->nfun(Kt->lx), n4 = h->nfun(Kt->ly);
    const int nsph = n1 * n2 * n3 * n4;
    DevBuf<int> dbi, dkj;
    DevBuf<double> dout;
    PC_CUDA(dbi.upload(g.bi, h->stream));
    PC_CUDA(dkj.upload(g.kj, h->stream));
    PC_CUDA(dout.alloc((size_t)nsph * m));
    PcEriArgs A;
    memset(&A, 0, sizeof(A));
    A.out = dout.p;
    if (launch_explicit(h, h->blocks_mode(), B, Kt, A, dbi.p, dkj.p, m)) return 1;
    std::vector<double> res((size_t)nsph * m);
    PC_CUDA(cudaMemcpyAsync(res.data(), dout.p, res.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    PC_CUDA(cudaStreamSynchronize(h->stream));
    for (int t = 0; t < m; ++t) {
      const int q = g.q[t];
      const int a = abcd[4 * q], b = abcd[4 * q + 1], c = abcd[4 * q + 2], d = abcd[4 * q + 3];
      const HostPair& P = h->pairs[h->pair_index(a, b)];
      const HostPair& Q = h->pairs[h->pair_index(c, d)];
      const int nfa = h->shells[a].nfn, nfb = h->shells[b].nfn, nfc = h->shells[c].nfn, nfd = h->shells[d].nfn;
      double* o = out + offsets[q];
      for (int ma = 0; ma < nfa; ++ma)
        for (int mb = 0; mb < nfb; ++mb)
          for (int mc = 0; mc < nfc; ++mc)
            for (int md = 0; md < nfd; ++md) {
              // position in the kernel's (x1 y1 | x2 y2) order
              const int px = P.swapped ? mb : ma, py = P.swapped ? ma : mb;
              const int qx = Q.swapped ? md : mc, qy = Q.swapped ? mc : md;
              int i1, i2, i3, i4;
              if (g.flip[t]) { i1 = qx; i2 = qy; i3 = px; i4 = py; }
              else { i1 = px; i2 = py; i3 = qx; i4 = qy; }
              const size_t k = (((size_t)i1 * n2 + i2) * n3 + i3) * n4 + i4;
              o[((ma * nfb + mb) * nfc + mc) * nfd + md] = res[k * m + t];
            }
    }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------
int pc_eri_tensor(pc_basis* h, double* G_dev, double* G_host) {
  if (!h || !G_dev) return fail("pc_eri_tensor: null");
  if (!h->planned) return fail("pc_eri_tensor: call pc_plan first");
  if (h->nranks != 1) return fail("pc_eri_tensor: needs a single-rank plan");
  if (!is_device_ptr(G_dev)) return fail("pc_eri_tensor: G_dev must be device memory");
  PC_CUDA(cudaSetDevice(h->device));
  const size_t N = h->nbf, n4 = N * N * N * N;
  PC_CUDA(cudaMemsetAsync(G_dev, 0, n4 * sizeof(double), h->stream));
  for (const LaunchGroup& g : h->groups) {
    PcEriArgs A;
    memset(&A, 0, sizeof(A));
    A.G = G_dev;
    if (launch_group(h, h->tensor_mode(), g, A, nullptr)) return 1;
  }
  if (G_host)
    PC_CUDA(cudaMemcpyAsync(G_host, G_dev, n4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  PC_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}

// ------------------------------------------------------------------------------------------
static int ensure_scratch(pc_basis* h) {
  const size_t nn = (size_t)h->nbf * h->nbf;
  if (h->acc.n < 3 * nn) PC_CUDA(h->acc.alloc(3 * nn));
  if (h->dstage.n < 3 * nn) PC_CUDA(h->dstage.alloc(3 * nn));
  if (h->ostage.n < 3 * nn) PC_CUDA(h->ostage.alloc(3 * nn));
  return 0;
}

static int copy_out(pc_basis* h, const double* dev, double* dst) {
  const size_t bytes = sizeof(double) * h->nbf * h->nbf;
  if (!dst || dst == dev) return 0;
  PC_CUDA(cudaMemcpyAsync(dst, dev, bytes, is_device_ptr(dst) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream));
  return 0;
}

int pc_jk_stored(pc_basis* h, const double* G_dev, const double* Dt, const double* Da,
                 const double* Db, double* J, double* Xa, double* Xb) {
  if (!h || !G_dev || !Dt || !Da || !Db) return fail("pc_jk_stored: null");
  PC_CUDA(cudaSetDevice(h->device));
  if (ensure_scratch(h)) return 1;
  const int N = h->nbf;
  const size_t nn = (size_t)N * N;
  const double *dt, *da, *db;
  if (stage_in(h, Dt, h->dstage.p, &dt) || stage_in(h, Da, h->dstage.p + nn, &da) ||
      stage_in(h, Db, h->dstage.p + 2 * nn, &db)) return 1;
  double* o = h->ostage.p;
  PC_CUDA(cudaMemsetAsync(o, 0, 3 * nn * sizeof(double), h->stream));
  {
    constexpr int NBG = 4;                       // slabs per CTA
    const int ngrp = (N + NBG - 1) / NBG;
    // Two kernels stream the tensor (profiles/r2c_stored_mode.txt, N = 192, 10.9 GB): plain 16-byte
    // loads reach 99 % of the measured HBM copy bandwidth, the TMA-staged pipeline 95 %; both
    // needed the per-CTA rotated order (66 % without).  Default: the faster one;
    // PYCHEM_B200_STORED_TMA=1 selects the TMA pipeline (read per call: bench.py times both).
    const char* tma_env = getenv("PYCHEM_B200_STORED_TMA");
    const bool use_tma = tma_env && atoi(tma_env) != 0;
    if (N % 2 == 0 && N >= 8 && use_tma) {
      // TMA-staged kernel: ~24 KB tiles (rt rows of the NBG slabs + Dt), 4 in flight per CTA
      constexpr int STAGES = 4;
      const int rt = std::max(1, (int)(24576 / ((size_t)(NBG + 1) * N * sizeof(double))));
      const int ct = std::min(256, ((N / 2 + 31) / 32) * 32);      // threads along the columns (pairs)
      const int rgn = std::max(1, 256 / ct);
      const size_t smem = ((size_t)STAGES * (NBG + 1) * rt * N + (size_t)2 * NBG * N + (size_t)2 * rgn * 2 * ct) * sizeof(double);
      if (N / 2 <= 256 && smem <= 200 * 1024) {
        static bool attr_set = false;
        if (!attr_set) {
          PC_CUDA(cudaFuncSetAttribute(jk_stored_tma_kernel<NBG, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
          attr_set = true;
        }
        jk_stored_tma_kernel<NBG, STAGES><<<N * ngrp, 288, smem, h->stream>>>(N, ngrp, rt, ct, G_dev, dt, da, db, o, o + nn, o + 2 * nn);
      } else {
        jk_stored_kernel<2, NBG><<<N * ngrp, 256, 0, h->stream>>>(N, ngrp, G_dev, dt, da, db, o, o + nn, o + 2 * nn);
      }
    } else if (N % 2 == 0) {
      jk_stored_kernel<2, NBG><<<N * ngrp, 256, 0, h->stream>>>(N, ngrp, G_dev, dt, da, db, o, o + nn, o + 2 * nn);
    } else {
      jk_stored_kernel<1, NBG><<<N * ngrp, 256, 0, h->stream>>>(N, ngrp, G_dev, dt, da, db, o, o + nn, o + 2 * nn);
    }
  }
  PC_CUDA(cudaGetLastError());
  h->launches += 1;
  if (copy_out(h, o, J) || copy_out(h, o + nn, Xa) || copy_out(h, o + 2 * nn, Xb)) return 1;
  PC_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}

int pc_jk_direct_accumulate(pc_basis* h, int variant, const double* Dt, const double* Da,
                            const double* Db, double* acc_dev) {
  if (!h || !Dt || !Da || !acc_dev) return fail("pc_jk_direct_accumulate: null");
  if (variant != PC_JK_RHF && variant != PC_JK_UHF && variant != PC_JK_GEN && variant != PC_MODE_NULL)
    return fail("pc_jk_direct_accumulate: bad variant");
  if (!h->planned) return fail("pc_jk_direct_accumulate: call pc_plan first");
  if (h->ints_type != 0) return fail("pc_jk_direct_accumulate: J/K digestion is defined for ints_type 0 only");
  if (!is_device_ptr(acc_dev)) return fail("pc_jk_direct_accumulate: acc_dev must be device memory");
  PC_CUDA(cudaSetDevice(h->device));
  if (ensure_scratch(h)) return 1;
  const size_t nn = (size_t)h->nbf * h->nbf;
  const double *dt, *da, *db;
  if (!Db) Db = Da;
  if (stage_in(h, Dt, h->dstage.p, &dt) || stage_in(h, Da, h->dstage.p + nn, &da) ||
      stage_in(h, Db, h->dstage.p + 2 * nn, &db)) return 1;
  pc_basis::GraphKey key;
  key.variant = variant; key.dt = dt; key.da = da; key.db = db; key.acc = acc_dev; key.plan_id = h->plan_id;
  const bool graphs = h->use_graphs && !h->profiling && h->groups.size() > 1;
  if (graphs && h->graph_exec && key == h->graph_key) {
    PC_CUDA(cudaGraphLaunch(h->graph_exec, h->stream));
    h->launches += h->graph_launches;
    return 0;
  }
  const long long launches_before = h->launches;
  if (graphs) PC_CUDA(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed));
  PC_CUDA(cudaMemsetAsync(acc_dev, 0, 3 * nn * sizeof(double), h->stream));
  if (h->profiling) {
    while (h->prof_events.size() < h->groups.size() + 1) {
      cudaEvent_t e;
      PC_CUDA(cudaEventCreate(&e));
      h->prof_events.push_back(e);
    }
  }
  static const int NSIDE = []() {
    const char* e = getenv("PYCHEM_B200_STREAMS");     // tuning knob (default 8 side streams)
    const int v = e ? atoi(e) : 8;
    return v < 1 ? 1 : (v > 64 ? 64 : v);
  }();
  const bool fan = !h->profiling && h->groups.size() > 1;
  if (fan) {
    while ((int)h->side.size() < NSIDE) {
      cudaStream_t st;
      PC_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
      h->side.push_back(st);
      cudaEvent_t e;
      PC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      h->ev_join.push_back(e);
    }
    if (!h->ev_fork) PC_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    PC_CUDA(cudaEventRecord(h->ev_fork, h->stream));
    for (int s2 = 0; s2 < NSIDE; ++s2) PC_CUDA(cudaStreamWaitEvent(h->side[s2], h->ev_fork, 0));
  }
  // measurement aid (tools/run_r2_floor.sh): PYCHEM_B200_DEBUG_LRANGE=lo,hi launches only the classes
  // with lo <= L <= hi -- the result is then incomplete by construction, the timing shows which
  // classes the concurrent step rests on
  static const int LR[2] = {[]() { const char* e = getenv("PYCHEM_B200_DEBUG_LRANGE"); return e ? atoi(e) : 0; }(),
                            []() { const char* e = getenv("PYCHEM_B200_DEBUG_LRANGE"); const char* c = e ? strchr(e, ',') : nullptr; return c ? atoi(c + 1) : 99; }()};
  static const int pair_l[6] = {0, 1, 2, 2, 3, 4};       // ss ps pp ds dp dd
  size_t idx = 0;
  for (const LaunchGroup& g : h->groups) {
    if (h->profiling) PC_CUDA(cudaEventRecord(h->prof_events[idx], h->stream));
    ++idx;
    if (g.pcb < 6 && g.pck < 6) {
      const int Lg = pair_l[g.pcb] + pair_l[g.pck];
      if (Lg < LR[0] || Lg > LR[1]) continue;
    }
    PcEriArgs A;
    memset(&A, 0, sizeof(A));
    A.Dj = dt; A.Da = da; A.Db = db;
    A.Jacc = acc_dev; A.Kaacc = acc_dev + nn; A.Kbacc = acc_dev + 2 * nn;
    A.out = acc_dev;
    if (launch_group(h, variant, g, A, fan ? h->side[idx % NSIDE] : h->stream)) return 1;
  }
  if (fan) {
    for (int s2 = 0; s2 < NSIDE; ++s2) {
      PC_CUDA(cudaEventRecord(h->ev_join[s2], h->side[s2]));
      PC_CUDA(cudaStreamWaitEvent(h->stream, h->ev_join[s2], 0));
    }
  }
  if (graphs) {
    cudaGraph_t graph = nullptr;
    PC_CUDA(cudaStreamEndCapture(h->stream, &graph));
    if (h->graph_exec) { cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; }
    cudaError_t ge = cudaGraphInstantiate(&h->graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ge != cudaSuccess) return fail(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ge));
    h->graph_key = key;
    h->graph_launches = h->launches - launches_before;
    PC_CUDA(cudaGraphLaunch(h->graph_exec, h->stream));
    return 0;
  }
  if (h->profiling) {
    PC_CUDA(cudaEventRecord(h->prof_events[idx], h->stream));
    PC_CUDA(cudaStreamSynchronize(h->stream));
    // one fused launch per group: its time is booked on the group's first bucket pair
    h->prof_ms.assign(h->plan.size(), 0.f);
    for (size_t k = 0; k < h->groups.size(); ++k) {
      float ms = 0.f;
      PC_CUDA(cudaEventElapsedTime(&ms, h->prof_events[k], h->prof_events[k + 1]));
      if (!h->groups[k].items.empty()) h->prof_ms[h->groups[k].items[0]] = ms;
    }
  }
  return 0;
}

int pc_set_profiling(pc_basis* h, int on) {
  if (!h) return fail("pc_set_profiling: null");
  h->profiling = on != 0;
  return 0;
}

int pc_plan_items(pc_basis* h, int max_items, int* n_items, int* cls, int* kprim,
                  long long* tasks, float* ms, double* prim_exec) {
  if (!h || !n_items) return fail("pc_plan_items: null");
  if (!h->planned) return fail("pc_plan_items: call pc_plan first");
  *n_items = (int)h->plan.size();
  for (int k = 0; k < (int)h->plan.size() && k < max_items; ++k) {
    const PlanItem& it = h->plan[k];
    const Kind* B = h->kinds[it.kb];
    const Kind* Kt = h->kinds[it.kk];
    if (cls) { cls[4 * k] = B->lx; cls[4 * k + 1] = B->ly; cls[4 * k + 2] = Kt->lx; cls[4 * k + 3] = Kt->ly; }
    if (kprim) { kprim[2 * k] = B->K; kprim[2 * k + 1] = Kt->K; }
    if (tasks) { tasks[2 * k] = it.total_q; tasks[2 * k + 1] = it.count_q; }
    if (ms) ms[k] = k < (int)h->prof_ms.size() ? h->prof_ms[k] : 0.f;
    if (prim_exec) prim_exec[k] = it.prim_exec;
  }
  return 0;
}

int pc_jk_finalize(pc_basis* h, int variant, const double* acc_dev, double* J, double* Xa,
                   double* Xb) {
  if (!h || !acc_dev) return fail("pc_jk_finalize: null");
  PC_CUDA(cudaSetDevice(h->device));
  if (ensure_scratch(h)) return 1;
  const int N = h->nbf;
  const size_t nn = (size_t)N * N;
  double* o = h->ostage.p;
  double* dj = is_device_ptr(J) ? J : o;
  double* dxa = is_device_ptr(Xa) ? Xa : o + nn;
  double* dxb = is_device_ptr(Xb) ? Xb : o + 2 * nn;
  const int blocks = (int)std::min<size_t>((nn + 255) / 256, 148 * 8);
  jk_finalize_kernel<<<blocks, 256, 0, h->stream>>>(N, variant == PC_JK_GEN, variant == PC_JK_RHF ? 1 : 2,
                                                    acc_dev, dj, dxa, dxb);
  PC_CUDA(cudaGetLastError());
  h->launches += 1;
  if (J && dj != J) { if (copy_out(h, dj, J)) return 1; }
  if (Xa && dxa != Xa) { if (copy_out(h, dxa, Xa)) return 1; }
  if (Xb && dxb != Xb) { if (copy_out(h, dxb, Xb)) return 1; }
  PC_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}

int pc_jk_classify(pc_basis* h, const double* Dt, const double* Da, const double* Db, int* variant) {
  if (!h || !Dt || !Da || !variant) return fail("pc_jk_classify: null");
  PC_CUDA(cudaSetDevice(h->device));
  if (ensure_scratch(h)) return 1;
  if (!Db) Db = Da;
  const size_t nn = (size_t)h->nbf * h->nbf;
  const double *dt, *da, *db;
  if (stage_in(h, Dt, h->dstage.p, &dt) || stage_in(h, Da, h->dstage.p + nn, &da) ||
      stage_in(h, Db, h->dstage.p + 2 * nn, &db)) return 1;
  if (!h->flags.p) PC_CUDA(h->flags.alloc(1));
  PC_CUDA(cudaMemsetAsync(h->flags.p, 0, sizeof(int), h->stream));
  const int blocks = (int)std::min<size_t>((nn + 255) / 256, 148 * 8);
  classify_kernel<<<blocks, 256, 0, h->stream>>>(h->nbf, dt, da, db, h->flags.p);
  PC_CUDA(cudaGetLastError());
  h->launches += 1;
  int f = 0;
  PC_CUDA(cudaMemcpyAsync(&f, h->flags.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  PC_CUDA(cudaStreamSynchronize(h->stream));
  *variant = (f & 1) ? PC_JK_GEN : ((f & 2) ? PC_JK_UHF : PC_JK_RHF);
  return 0;
}

// Classify on the device and digest straight from the staged copies (one upload only).  The
// 4-byte flag needs a host round trip before the right digestion kernels can be chosen; an SCF
// hands over the same kind of densities call after call, so the digestion of the PREVIOUS call's
// variant is queued behind the classification at once and the host reads the flag while it runs.
// A different answer (first call of another kind) queues the right digestion after it -- the
// accumulators are cleared by every pc_jk_direct_accumulate.  PYCHEM_B200_SPECULATE=0: wait first.
static int classify_then_accumulate(pc_basis* h, const double* Dt, const double* Da, const double* Db,
                                    double* acc_dev, int* variant) {
  if (!Db) Db = Da;
  const size_t nn = (size_t)h->nbf * h->nbf;
  const double *dt, *da, *db;
  if (stage_in(h, Dt, h->dstage.p, &dt) || stage_in(h, Da, h->dstage.p + nn, &da) ||
      stage_in(h, Db, h->dstage.p + 2 * nn, &db)) return 1;
  if (!h->flags.p) PC_CUDA(h->flags.alloc(1));
  if (!h->flag_host) PC_CUDA(cudaMallocHost(reinterpret_cast<void**>(&h->flag_host), sizeof(int)));
  if (!h->ev_flag) PC_CUDA(cudaEventCreateWithFlags(&h->ev_flag, cudaEventDisableTiming));
  PC_CUDA(cudaMemsetAsync(h->flags.p, 0, sizeof(int), h->stream));
  const int blocks = (int)std::min<size_t>((nn + 255) / 256, 148 * 8);
  classify_kernel<<<blocks, 256, 0, h->stream>>>(h->nbf, dt, da, db, h->flags.p);
  PC_CUDA(cudaGetLastError());
  h->launches += 1;
  PC_CUDA(cudaMemcpyAsync(h->flag_host, h->flags.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  PC_CUDA(cudaEventRecord(h->ev_flag, h->stream));
  static const bool speculate = []() { const char* e = getenv("PYCHEM_B200_SPECULATE"); return !(e && e[0] == '0'); }();
  const int guess = speculate ? h->guess_variant : 0;
  if (guess && pc_jk_direct_accumulate(h, guess, dt, da, db, acc_dev)) return 1;
  PC_CUDA(cudaEventSynchronize(h->ev_flag));
  const int f = *h->flag_host;
  const int v = (f & 1) ? PC_JK_GEN : ((f & 2) ? PC_JK_UHF : PC_JK_RHF);
  if (v != guess && pc_jk_direct_accumulate(h, v, dt, da, db, acc_dev)) return 1;
  h->guess_variant = v;
  *variant = v;
  return 0;
}

int pc_jk_direct_accumulate_auto(pc_basis* h, const double* Dt, const double* Da, const double* Db,
                                 double* acc_dev, int* variant) {
  if (!h || !variant) return fail("pc_jk_direct_accumulate_auto: null");
  PC_CUDA(cudaSetDevice(h->device));
  if (ensure_scratch(h)) return 1;
  return classify_then_accumulate(h, Dt, Da, Db, acc_dev, variant);
}

int pc_jk_direct(pc_basis* h, int variant, const double* Dt, const double* Da, const double* Db,
                 double* J, double* Xa, double* Xb) {
  if (!h) return fail("pc_jk_direct: null");
  PC_CUDA(cudaSetDevice(h->device));
  if (ensure_scratch(h)) return 1;
  if (variant == PC_JK_AUTO) {
    if (classify_then_accumulate(h, Dt, Da, Db, h->acc.p, &variant)) return 1;
    return pc_jk_finalize(h, variant, h->acc.p, J, Xa, Xb);
  }
  if (pc_jk_direct_accumulate(h, variant, Dt, Da, Db, h->acc.p)) return 1;
  return pc_jk_finalize(h, variant, h->acc.p, J, Xa, Xb);
}

int pc_jk_direct_auto(pc_basis* h, const double* Dt, const double* Da, const double* Db, double* J,
                      double* Xa, double* Xb, int* variant) {
  if (!h || !variant) return fail("pc_jk_direct_auto: null");
  PC_CUDA(cudaSetDevice(h->device));
  if (ensure_scratch(h)) return 1;
  if (classify_then_accumulate(h, Dt, Da, Db, h->acc.p, variant)) return 1;
  // closed-shell densities: X_beta == X_alpha, nothing is copied into Xb
  return pc_jk_finalize(h, *variant, h->acc.p, J, Xa, *variant == PC_JK_RHF ? nullptr : Xb);
}

// ------------------------------------------------------------------------------------------
// batched J/K for general density sets (the NOCI determinant pairs)
// ------------------------------------------------------------------------------------------
static int ensure_batch_scratch(pc_basis* h, int nset, bool acc, bool din, bool dout) {
  const size_t need = (size_t)nset * 3 * h->nbf * h->nbf;
  if (acc && h->bacc.n < need) PC_CUDA(h->bacc.alloc(need));
  if (din && h->bdstage.n < need) PC_CUDA(h->bdstage.alloc(need));
  if (dout && h->bostage.n < need) PC_CUDA(h->bostage.alloc(need));
  return 0;
}

static int stage_in_batch(pc_basis* h, int nset, const double* D, const double** out) {
  if (is_device_ptr(D)) {
    *out = D;
    return 0;
  }
  if (ensure_batch_scratch(h, nset, false, true, false)) return 1;
  PC_CUDA(cudaMemcpyAsync(h->bdstage.p, D, sizeof(double) * nset * 3 * h->nbf * h->nbf, cudaMemcpyHostToDevice, h->stream));
  *out = h->bdstage.p;
  return 0;
}

int pc_jk_direct_batch_accumulate(pc_basis* h, int nset, const double* D, double* acc_dev) {
  if (!h || !D || !acc_dev || nset < 1) return fail("pc_jk_direct_batch_accumulate: bad arguments");
  if (!h->planned) return fail("pc_jk_direct_batch_accumulate: call pc_plan first");
  if (h->ints_type != 0) return fail("pc_jk_direct_batch_accumulate: J/K digestion is defined for ints_type 0 only");
  if (!is_device_ptr(acc_dev)) return fail("pc_jk_direct_batch_accumulate: acc_dev must be device memory");
  PC_CUDA(cudaSetDevice(h->device));
  const size_t nn = (size_t)h->nbf * h->nbf;
  const double* d;
  if (stage_in_batch(h, nset, D, &d)) return 1;
  PC_CUDA(cudaMemsetAsync(acc_dev, 0, sizeof(double) * nset * 3 * nn, h->stream));
  // the class launches are independent (they only meet in the atomics): spread them over side
  // streams like pc_jk_direct_accumulate does
  const int nside = 4;
  const bool fan = h->groups.size() > 1;
  if (fan) {
    while ((int)h->side.size() < nside) {
      cudaStream_t st;
      PC_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
      h->side.push_back(st);
      cudaEvent_t e;
      PC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      h->ev_join.push_back(e);
    }
    if (!h->ev_fork) PC_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    PC_CUDA(cudaEventRecord(h->ev_fork, h->stream));
    for (int s2 = 0; s2 < nside; ++s2) PC_CUDA(cudaStreamWaitEvent(h->side[s2], h->ev_fork, 0));
  }
  size_t idx = 0;
  for (const LaunchGroup& g : h->groups) {
    PcEriArgs A;
    memset(&A, 0, sizeof(A));
    A.Dj = d; A.Da = d + nn; A.Db = d + 2 * nn;
    A.Jacc = acc_dev; A.Kaacc = acc_dev + nn; A.Kbacc = acc_dev + 2 * nn;
    A.out = acc_dev;
    A.nset = nset;
    A.set_stride = (long long)(3 * nn);
    if (launch_group(h, PC_MODE_JK_GEN_BATCH, g, A, fan ? h->side[idx % nside] : h->stream)) return 1;
    ++idx;
  }
  if (fan) {
    for (int s2 = 0; s2 < nside; ++s2) {
      PC_CUDA(cudaEventRecord(h->ev_join[s2], h->side[s2]));
      PC_CUDA(cudaStreamWaitEvent(h->stream, h->ev_join[s2], 0));
    }
  }
  return 0;
}

int pc_jk_finalize_batch(pc_basis* h, int nset, const double* acc_dev, double* out) {
  if (!h || !acc_dev || !out || nset < 1) return fail("pc_jk_finalize_batch: bad arguments");
  PC_CUDA(cudaSetDevice(h->device));
  const int N = h->nbf;
  const size_t nn = (size_t)N * N;
  double* o = out;
  if (!is_device_ptr(out)) {
    if (ensure_batch_scratch(h, nset, false, false, true)) return 1;
    o = h->bostage.p;
  }
  const int blocks = (int)std::min<size_t>((nn + 255) / 256, 148 * 8);
  for (int s = 0; s < nset; ++s) {
    const double* a = acc_dev + (size_t)s * 3 * nn;
    double* os = o + (size_t)s * 3 * nn;
    jk_finalize_kernel<<<blocks, 256, 0, h->stream>>>(N, 1, 2, a, os, os + nn, os + 2 * nn);
    PC_CUDA(cudaGetLastError());
    h->launches += 1;
  }
  if (o != out)
    PC_CUDA(cudaMemcpyAsync(out, o, sizeof(double) * nset * 3 * nn, cudaMemcpyDeviceToHost, h->stream));
  PC_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}

int pc_jk_direct_batch(pc_basis* h, int nset, const double* D, double* out) {
  if (!h || nset < 1) return fail("pc_jk_direct_batch: bad arguments");
  PC_CUDA(cudaSetDevice(h->device));
  if (ensure_batch_scratch(h, nset, true, false, false)) return 1;
  if (pc_jk_direct_batch_accumulate(h, nset, D, h->bacc.p)) return 1;
  return pc_jk_finalize_batch(h, nset, h->bacc.p, out);
}

int pc_jk_stored_batch(pc_basis* h, const double* G_dev, int nset, const double* D, double* out) {
  if (!h || !G_dev || !D || !out || nset < 1) return fail("pc_jk_stored_batch: bad arguments");
  PC_CUDA(cudaSetDevice(h->device));
  const int N = h->nbf;
  const size_t nn = (size_t)N * N;
  const double* d;
  if (stage_in_batch(h, nset, D, &d)) return 1;
  double* o = out;
  if (!is_device_ptr(out)) {
    if (ensure_batch_scratch(h, nset, false, false, true)) return 1;
    o = h->bostage.p;
  }
  PC_CUDA(cudaMemsetAsync(o, 0, sizeof(double) * nset * 3 * nn, h->stream));
  constexpr int NBG = 4, NS = 4;                 // slabs per CTA, density sets per pass over the tensor
  const int ngrp = (N + NBG - 1) / NBG;
  for (int s0 = 0; s0 < nset; s0 += NS) {
    const int ns = std::min(NS, nset - s0);
    const double* ds = d + (size_t)s0 * 3 * nn;
    double* os = o + (size_t)s0 * 3 * nn;
    if (N % 2 == 0) jk_stored_batch_kernel<2, NBG, NS><<<N * ngrp, 256, 0, h->stream>>>(N, ngrp, ns, 3 * nn, G_dev, ds, os);
    else jk_stored_batch_kernel<1, NBG, NS><<<N * ngrp, 256, 0, h->stream>>>(N, ngrp, ns, 3 * nn, G_dev, ds, os);
    PC_CUDA(cudaGetLastError());
    h->launches += 1;
  }
  if (o != out)
    PC_CUDA(cudaMemcpyAsync(out, o, sizeof(double) * nset * 3 * nn, cudaMemcpyDeviceToHost, h->stream));
  PC_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}

int pc_one_electron(pc_basis* h, int natom, const double* Z, const double* R, double* core,
                    double* overlap) {
  if (!h || natom <= 0 || !Z || !R || !core || !overlap) return fail("pc_one_electron: bad arguments");
  PC_CUDA(cudaSetDevice(h->device));
  if (ensure_scratch(h)) return 1;
  const size_t nn = (size_t)h->nbf * h->nbf;
  std::vector<double> zr(4 * (size_t)natom);
  for (int c = 0; c < natom; ++c) {
    zr[c] = Z[c];
    for (int k = 0; k < 3; ++k) zr[natom + 3 * c + k] = R[3 * c + k];
  }
  DevBuf<double> dzr;
  PC_CUDA(dzr.upload(zr, h->stream));
  PcShellTable S;
  S.l = h->d_l.p; S.K = h->d_K.p; S.poff = h->d_poff.p; S.first_fn = h->d_fn.p; S.A = h->d_A.p;
  S.exps = h->d_exps.p; S.scc = h->d_scc.p; S.pair_a = h->d_pa.p; S.pair_b = h->d_pb.p;
  S.npair = (int)h->pairs.size(); S.cart_d = h->cart_d ? 1 : 0; S.nbf = h->nbf;
  double* o = h->ostage.p;
  double* dcore = is_device_ptr(core) ? core : o;
  double* dov = is_device_ptr(overlap) ? overlap : o + nn;
  if (h->max_l <= 2) one_electron_kernel<2><<<(S.npair + 63) / 64, 64, 0, h->stream>>>(S, natom, dzr.p, dzr.p + natom, h->boys.p, dcore, dov);
  else one_electron_kernel<3><<<(S.npair + 63) / 64, 64, 0, h->stream>>>(S, natom, dzr.p, dzr.p + natom, h->boys.p, dcore, dov);
  PC_CUDA(cudaGetLastError());
  h->launches += 1;
  if (dcore != core) { if (copy_out(h, dcore, core)) return 1; }
  if (dov != overlap) { if (copy_out(h, dov, overlap)) return 1; }
  PC_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}

int pc_plan_segments_host(int nb, const double* pm_bra, int nbg, const int* gstart_bra, int nk,
                           const double* pm_ket, int nkg, const int* gstart_ket, int same, int run,
                           double thresh, int max_seg, int* n_seg, long long* seg_off, int* seg_ij,
                           long long* seg_quartets) {
  if (nb < 0 || nk < 0 || !pm_bra || !pm_ket || !gstart_bra || !gstart_ket || !n_seg)
    return fail("pc_plan_segments_host: bad arguments");
  std::vector<double> pB(pm_bra, pm_bra + nb), pK(pm_ket, pm_ket + nk);
  std::vector<int> gB(gstart_bra, gstart_bra + nbg + 1), gK(gstart_ket, gstart_ket + nkg + 1);
  std::vector<int> kB(nb, 1), kK(nk, 1), ij;
  std::vector<long long> off, q;
  std::vector<double> prim;
  build_segments_host(pB, gB, kB, pK, gK, kK, same, run, thresh, off, q, prim, ij);
  *n_seg = (int)ij.size() / 2;
  if (*n_seg > max_seg) return fail("pc_plan_segments_host: max_seg too small");
  if (seg_off) std::copy(off.begin(), off.end(), seg_off);
  if (seg_ij) std::copy(ij.begin(), ij.end(), seg_ij);
  if (seg_quartets) std::copy(q.begin(), q.end(), seg_quartets);
  return 0;
}

int pc_fp64_peak(int device, double* tflops) {
  if (!tflops) return fail("pc_fp64_peak: null");
  PC_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  PC_CUDA(cudaGetDeviceProperties(&prop, device));
  double* d = nullptr;
  PC_CUDA(cudaMalloc((void**)&d, 8));
  cudaEvent_t e0, e1;
  PC_CUDA(cudaEventCreate(&e0));
  PC_CUDA(cudaEventCreate(&e1));
  const int iters = 4096, threads = 256, blocks = prop.multiProcessorCount * 8;
  double best = 0;
  for (int rep = 0; rep < 5; ++rep) {
    PC_CUDA(cudaEventRecord(e0));
    dfma_peak_kernel<<<blocks, threads>>>(d, iters, 1.0 + rep);
    PC_CUDA(cudaEventRecord(e1));
    PC_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    PC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 64.0 * iters * (double)threads * blocks;
    if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  *tflops = best;
  return 0;
}

}  // extern "C"
