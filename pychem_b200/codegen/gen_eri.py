#!/usr/bin/env python
"""Generate the per-class ERI kernels (sm_100a, FP64) as straight-line CUDA.

For every angular-momentum class (lx1 ly1 | lx2 ly2) with lx >= ly inside a pair and
pair-class(bra) >= pair-class(ket) this writes pychem_b200/csrc/gen/eri_<class>.cu holding

  * the primitive-quartet body: Head-Gordon-Pople vertical recursion in the reference's
    Gill-scaled form (Methods/c_ints/two_electron_vrr.c:92-108; ket built first on the s bra,
    then the bra, Methods/integrals.py:514-517), reduction direction = first non-zero of x,y,z of
    the target component (two_electron_vrr.c:33-48), pruned to the components actually needed,
    accumulating the contracted (e0|f0) in registers (two_electron_contract.c:45);
  * the contracted tail: horizontal recursion (two_electron_hrr.c:82; ket first, then bra,
    integrals.py:531-536) and normalisation + cart->spherical (integrals.py:541-547,
    Data/transform_basis.py:8-12) with the constants folded;
  * one __global__ kernel per output mode and a host launcher.

The recursion DAG that the reference rebuilds in Python for every shell quartet
(integrals.SetRR2, integrals.py:73-191) is resolved here once, at code-generation time.

Kernel forms (make_class picks one per class): ClassGen -- straight-line, one quartet per thread
(`source_single`) or one ket pair x a run of bra pairs per thread (`source_run`, RUN_CLASSES);
ClassGenPass -- the primitive loops once per group of bra components, accumulators in registers
(PASS_CLASSES: the six classes whose contracted block does not fit the registers);
ClassGenV2 -- rolled loops for (dd|dd).  CoopGen (PC_GEN_COOP) is the measured-and-shelved
warp-cooperative variant of the same decomposition as ClassGenPass.
"""
import math
import os
import sys

LNAME = "spdf"
PAIR_CLASSES = [(0, 0), (1, 0), (1, 1), (2, 0), (2, 1), (2, 2)]   # (lx, ly), lx >= ly, l <= 2


def ncart(l):
    return (l + 1) * (l + 2) // 2


def ncum(l):
    return (l + 1) * (l + 2) * (l + 3) // 6 if l >= 0 else 0


def comps(l):
    return [(lx, ly, l - lx - ly) for lx in range(l, -1, -1) for ly in range(l - lx, -1, -1)]


def cidx(c):
    lx, ly, lz = c
    return (ly + lz) * (ly + lz + 1) // 2 + lz


def cum(c):
    return ncum(sum(c) - 1) + cidx(c)


def first_dir(c):
    return 0 if c[0] else (1 if c[1] else 2)


def dec(c, d, n=1):
    c = list(c)
    c[d] -= n
    return tuple(c)


def inc(c, d):
    c = list(c)
    c[d] += 1
    return tuple(c)


def nsph(l):
    return 2 * l + 1


# cart -> spherical with the component-dependent part of the normalisation folded in, relative
# to the (uniform) factor that is folded into the pair prefactor on the host:
#   s: 1 ; p: identity ; d: nm_xx/nm_xy = 1/sqrt(3)  (Util/structures.py:850-856)
def c2s_rows(l):
    if l == 0:
        return [[(0, 1.0)]]
    if l == 1:
        return [[(0, 1.0)], [(1, 1.0)], [(2, 1.0)]]
    if l == 2:
        r3 = 1.0 / math.sqrt(3.0)
        # cart order xx xy xz yy yz zz
        return [[(0, 0.5), (3, -0.5)], [(1, 1.0)], [(2, 1.0)], [(4, 1.0)],
                [(0, -0.5 * r3), (3, -0.5 * r3), (5, r3)]]
    raise ValueError(l)


def nfun(l, cart_d=False):
    """functions per shell: 2l+1 real spherical, or the 6 Cartesians for d when Cartesian_L = [2]
    (Util/structures.py:844-849)."""
    return ncart(l) if (cart_d and l == 2) else nsph(l)


def c2s_rows_any(l, cart_d=False):
    """cart->function rows incl. the component-dependent normalisation ratio.  Cartesian d keeps
    all six components (CartToSpher = identity) and only carries nm_xx/nm_xy = 1/sqrt(3)."""
    if cart_d and l == 2:
        r3 = 1.0 / math.sqrt(3.0)
        return [[(0, r3)], [(1, 1.0)], [(2, 1.0)], [(3, r3)], [(4, 1.0)], [(5, r3)]]
    return c2s_rows(l)


class Emit:
    def __init__(self, prefix):
        self.lines = []
        self.n = 0
        self.ops = 0          # multiply/fma operations emitted by lin_comb
        self.prefix = prefix

    def new(self, expr):
        name = "%s%d" % (self.prefix, self.n)
        self.n += 1
        self.lines.append("const double %s = %s;" % (name, expr))
        return name

    def raw(self, line):
        self.lines.append(line)


def fmt(x):
    return repr(float(x))


def lin_comb(em, terms):
    """terms: [(coef, var)] -> expression variable."""
    terms = [(c, v) for c, v in terms if c != 0.0]
    if len(terms) == 1 and terms[0][0] == 1.0:
        return terms[0][1]
    expr = None
    em.ops += len(terms)
    for c, v in terms:
        if expr is None:
            expr = v if c == 1.0 else "%s * %s" % (fmt(c), v)
        else:
            expr = "fma(%s, %s, %s)" % (fmt(c), v, expr)
    return em.new(expr)


class ClassGen:
    def __init__(self, lx1, ly1, lx2, ly2, cart_d=False):
        self.l = (lx1, ly1, lx2, ly2)
        self.cart_d = bool(cart_d) and 2 in self.l
        self.La, self.Lc = lx1 + ly1, lx2 + ly2
        self.L = self.La + self.Lc
        names = "spDf" if self.cart_d else LNAME
        self.name = "".join(names[x] for x in self.l)
        # contracted (e0|f0): e = lx1..La, f = lx2..Lc, all components
        self.e_list = [c for le in range(lx1, self.La + 1) for c in comps(le)]
        self.f_list = [c for lf in range(lx2, self.Lc + 1) for c in comps(lf)]
        self.ne, self.nf = len(self.e_list), len(self.f_list)
        self.nsph = [nfun(x, self.cart_d) for x in self.l]

    # ------------------------------------------------------------------ VRR
    def gen_vrr(self):
        """Vertical recursion of one primitive quartet + accumulation into acc[].  Elements that
        only feed the accumulation (never an operand of a higher element) take acc as the seed of
        their FMA chain: n FMAs instead of a multiply, n-1 FMAs and an add."""
        zero = (0, 0, 0)
        # pass 1: which elements are operands of other elements
        operand = set()
        seen = set()

        def walk(a, c, m, top):
            key = (a, c, m)
            if not top:
                operand.add(key)
            if key in seen:
                return
            seen.add(key)
            if a == zero and c == zero:
                return
            if a == zero:
                d = first_dir(c)
                c0 = dec(c, d)
                walk(zero, c0, m, False); walk(zero, c0, m + 1, False)
                if c0[d] > 0:
                    c1 = dec(c0, d)
                    walk(zero, c1, m, False); walk(zero, c1, m + 1, False)
            else:
                d = first_dir(a)
                a0 = dec(a, d)
                walk(a0, c, m, False); walk(a0, c, m + 1, False)
                if a0[d] > 0:
                    a1 = dec(a0, d)
                    walk(a1, c, m, False); walk(a1, c, m + 1, False)
                if c[d] > 0:
                    walk(a0, dec(c, d), m + 1, False)

        for e in self.e_list:
            for f in self.f_list:
                walk(e, f, 0, True)

        em = Emit("v")
        memo = {}
        self.vrr_refs = 0

        def get(a, c, m, seed=None):
            """seed: name of the accumulator the element is added to (accumulate-only elements)."""
            key = (a, c, m)
            if key in memo:
                return memo[key]
            if a == zero and c == zero:
                val = "F[%d]" % m
            elif a == zero:
                d = first_dir(c)
                c0 = dec(c, d)
                n = c0[d]
                b0, b1 = get(zero, c0, m), get(zero, c0, m + 1)
                tail = "Re%d * %s" % (d, b1) if seed is None else "fma(Re%d, %s, %s)" % (d, b1, seed)
                expr = "fma(QX%d, %s, %s)" % (d, b0, tail)
                self.vrr_refs += 2 + (2 if n > 0 else 0)
                if n > 0:
                    c1 = dec(c0, d)
                    b2, b3 = get(zero, c1, m), get(zero, c1, m + 1)
                    expr = "fma(ne%d, fma(-eta, %s, %s), %s)" % (n, b3, b2, expr)
                if seed is not None:
                    return expr
                val = em.new(expr)
            else:
                d = first_dir(a)
                a0 = dec(a, d)
                n = a0[d]
                b0, b1 = get(a0, c, m), get(a0, c, m + 1)
                tail = "Rz%d * %s" % (d, b1) if seed is None else "fma(Rz%d, %s, %s)" % (d, b1, seed)
                expr = "fma(PX%d, %s, %s)" % (d, b0, tail)
                self.vrr_refs += 2 + (2 if n > 0 else 0) + (1 if c[d] > 0 else 0)
                if n > 0:
                    a1 = dec(a0, d)
                    b2, b3 = get(a1, c, m), get(a1, c, m + 1)
                    expr = "fma(nz%d, fma(-zeta, %s, %s), %s)" % (n, b3, b2, expr)
                if c[d] > 0:
                    b4 = get(a0, dec(c, d), m + 1)
                    expr = "fma(nze%d, %s, %s)" % (c[d], b4, expr)
                if seed is not None:
                    return expr
                val = em.new(expr)
            memo[key] = val
            return val

        n_fused = 0
        for ie, e in enumerate(self.e_list):
            for jf, f in enumerate(self.f_list):
                k = ie * self.nf + jf
                if (e, f, 0) not in operand and not (e == zero and f == zero) and FUSE_ACC:
                    em.raw("acc[%d] = %s;" % (k, get(e, f, 0, seed="acc[%d]" % k)))
                    n_fused += 1
                else:
                    em.raw("acc[%d] += %s;" % (k, get(e, f, 0)))
        self.n_vrr = em.n + n_fused
        return em.lines

    # ------------------------------------------------------------------ HRR + c2s
    def gen_tail(self):
        lx1, ly1, lx2, ly2 = self.l
        em = Emit("h")
        self.hrr_el = 0
        self.c2s_ops = 0
        f_index = {c: i for i, c in enumerate(self.f_list)}
        e_index = {c: i for i, c in enumerate(self.e_list)}
        nqs = nfun(lx2, self.cart_d) * nfun(ly2, self.cart_d)

        # ---- ket HRR for every bra e-component, then ket cart->sph
        ks = {}     # (ie, q) -> var
        for ie in range(self.ne):
            memo = {}

            def hk(cx, cy):
                key = (cx, cy)
                if key in memo:
                    return memo[key]
                if sum(cy) == 0:
                    val = "acc[%d]" % (ie * self.nf + f_index[cx])
                else:
                    d = first_dir(cy)
                    cy0 = dec(cy, d)
                    val = em.new("fma(CD%d, %s, %s)" % (d, hk(cx, cy0), hk(inc(cx, d), cy0)))
                    self.hrr_el += 1
                memo[key] = val
                return val

            cart = {(ix, iy): hk(cx, cy) for ix, cx in enumerate(comps(lx2)) for iy, cy in enumerate(comps(ly2))}
            # transform second index then first
            half = {}
            for ix in range(ncart(lx2)):
                for my, row in enumerate(c2s_rows_any(ly2, self.cart_d)):
                    half[(ix, my)] = lin_comb(em, [(c, cart[(ix, iy)]) for iy, c in row])
            for mx, row in enumerate(c2s_rows_any(lx2, self.cart_d)):
                for my in range(nfun(ly2, self.cart_d)):
                    ks[(ie, mx * nfun(ly2, self.cart_d) + my)] = lin_comb(em, [(c, half[(ix, my)]) for ix, c in row])

        # ---- bra HRR for every ket spherical component, then bra cart->sph
        out = []
        for q in range(nqs):
            memo = {}

            def hb(cx, cy):
                key = (cx, cy)
                if key in memo:
                    return memo[key]
                if sum(cy) == 0:
                    val = ks[(e_index[cx], q)]
                else:
                    d = first_dir(cy)
                    cy0 = dec(cy, d)
                    val = em.new("fma(AB%d, %s, %s)" % (d, hb(cx, cy0), hb(inc(cx, d), cy0)))
                    self.hrr_el += 1
                memo[key] = val
                return val

            cart = {(ix, iy): hb(cx, cy) for ix, cx in enumerate(comps(lx1)) for iy, cy in enumerate(comps(ly1))}
            half = {}
            for ix in range(ncart(lx1)):
                for my, row in enumerate(c2s_rows_any(ly1, self.cart_d)):
                    half[(ix, my)] = lin_comb(em, [(c, cart[(ix, iy)]) for iy, c in row])
            for mx, row in enumerate(c2s_rows_any(lx1, self.cart_d)):
                for my in range(nfun(ly1, self.cart_d)):
                    v = lin_comb(em, [(c, half[(ix, my)]) for ix, c in row])
                    p = mx * nfun(ly1, self.cart_d) + my
                    out.append("g[%d] = %s;" % (p * nqs + q, v))
        self.n_tail = em.n
        self.c2s_ops = em.ops
        return em.lines + out

    # ------------------------------------------------------------------ file
    V2 = False

    def prim_prologue(self, s, nmax):
        """Primitive loops: ket primitives outside (per-lane, coalesced SoA loads, once per ket
        primitive), bra primitives inside.  The lanes of a segment share the bra pair: its
        primitives come from the pair's contiguous RECORD (PcPairKind::rec, 48 bytes each), loaded
        one primitive ahead so that the load latency overlaps the recursion of the current one
        (the prefetch wraps to primitive 0 for the next ket primitive)."""
        ket_load = ["const double2 q0 = __ldg(kp), q1 = __ldg(kp + nk), q2 = __ldg(kp + 2 * (size_t)nk);",
                    "const double sQ = q0.x, UQ = q0.y, Qx = q1.x, Qy = q1.y, Qz = q2.x, kzQ = q2.y;",
                    "const double eta = 0.5 * sQ;",
                    "const double QX0 = -CD0 * kzQ, QX1 = -CD1 * kzQ, QX2 = -CD2 * kzQ;"]
        ket_load += ["const double ne%d = %d.0 * eta;" % (n, n) for n in range(1, nmax + 1)]
        if PREFETCH == 1:
            bra_load = ["const double2 p0 = n0, p1 = n1, p2 = n2;",
                        "{ const double2* __restrict__ nx = (ib + 1 < KB) ? bq + 3 : brec + 3; n0 = __ldg(nx); n1 = __ldg(nx + 1); n2 = __ldg(nx + 2); }"]
        else:
            bra_load = ["const double2 p0 = __ldg(bq), p1 = __ldg(bq + 1), p2 = __ldg(bq + 2);"]
        bra_load += [
                    "const double sP = p0.x, UP = p0.y, Px = p1.x, Py = p1.y, Pz = p2.x, kzP = p2.y;",
                    "const double zeta = 0.5 * sP;",
                    "const double PX0 = -AB0 * kzP, PX1 = -AB1 * kzP, PX2 = -AB2 * kzP;"]
        bra_load += ["const double nz%d = %d.0 * zeta;" % (n, n) for n in range(1, nmax + 1)]
        s.append("  const double2* __restrict__ kp = reinterpret_cast<const double2*>(I.ket.prim) + j;")
        if PREFETCH == 1:
            s.append("  double2 n0 = __ldg(brec + 3), n1 = __ldg(brec + 4), n2 = __ldg(brec + 5);")
        s.append("  for (int ik = 0; ik < KK; ++ik, kp += 3 * (size_t)nk) {")
        s.extend("    " + l for l in ket_load)
        s.append("    const double2* __restrict__ bq = brec + 3;")
        if UNROLL_IB > 1 and self.L <= UNROLL_IB_MAXL and not self.V2:
            s.append("#pragma unroll %d" % UNROLL_IB)
        s.append("    for (int ib = 0; ib < KB; ++ib, bq += 3) {")
        s.extend("      " + l for l in bra_load)
        s.append("      const double R0 = Px - Qx, R1 = Py - Qy, R2_ = Pz - Qz;")
        s.append("      const double Rsq = R0 * R0 + R1 * R1 + R2_ * R2_;")
        s.append("      double F[L + 1];")
        s.append("      if (MODE == PC_MODE_BLOCKS_SCAT || MODE == PC_MODE_TENSOR_SCAT) pc_fundamentals_scatter<L>(sP, UP, sQ, UQ, Rsq, A.scat_S, F); "
                 "else pc_fundamentals<L>(sP, UP, sQ, UQ, Rsq, A.boys, F);")
        s.append("      const double Rz0 = -R0 * zeta, Rz1 = -R1 * zeta, Rz2 = -R2_ * zeta;")
        s.append("      const double Re0 = R0 * eta, Re1 = R1 * eta, Re2 = R2_ * eta;")
        s.append("      const double ze = zeta * eta;")
        for n in range(1, nmax + 1):
            s.append("      const double nze%d = %d.0 * ze;" % (n, n))
        s.append("      (void)QX0; (void)QX1; (void)QX2; (void)PX0; (void)PX1; (void)PX2; (void)ze;")
        s.append("      (void)Rz0; (void)Rz1; (void)Rz2; (void)Re0; (void)Re1; (void)Re2;")

    def close_prim_loops(self, s):
        s.append("    }")
        s.append("  }")

    def bra_record(self, s, ivar, indent="  "):
        """Header of the bra pair's record: X - Y, Schwarz maximum, first functions, pair id, keff."""
        s.append(indent + "const double2* __restrict__ brec = reinterpret_cast<const double2*>(I.bra.rec) + (size_t)%s * (3 * (I.bra.K + 1));" % ivar)
        s.append(indent + "const double2 bh0 = __ldg(brec), bh1 = __ldg(brec + 1);")
        s.append(indent + "const int4 bh2 = __ldg(reinterpret_cast<const int4*>(brec + 2));")
        s.append(indent + "const double AB0 = bh0.x, AB1 = bh0.y, AB2 = bh1.x;")
        s.append(indent + "const int fb = bh2.y, pidb = bh2.z, KB = bh2.w;")
        if PREFETCH == 2:
            # the rest of the record (48 (K + 1) bytes from brec) into L1 while the ket side is loaded
            s.append(indent + "for (int ln = 128; ln < 48 * (KB + 1); ln += 128) pc_prefetch_l1(reinterpret_cast<const char*>(brec) + ln);")

    def block_size(self):
        return 128 if self.L <= 4 else 64

    def min_blocks(self):
        # occupancy floor: the low classes are latency-bound (ncu: long-scoreboard stalls), more
        # resident warps beat a few spilled registers
        override = os.environ.get("PC_GEN_MINB_L%d" % self.L)
        if override:
            return int(override)
        return {0: 8, 1: 8, 2: 6, 3: 4}.get(self.L, 1)   # tuned on (H2O)32, profiles/README.md

    run = 1          # longest bra run one thread walks (1: one quartet per thread)

    def source(self):
        if self.run > 1:
            return self.source_run()
        return self.source_single()

    def launcher(self, block):
        s = []
        s.append("cudaError_t pc_launch_%s(int mode, const PcEriArgs& A, cudaStream_t st) {" % self.name)
        s.append("  if (A.nwarps <= 0) return cudaSuccess;")
        if getattr(self, "coop", None) is not None:
            s.extend(self.coop.launch_lines())
        s.append("  const int block = %d;" % block)
        s.append("  const unsigned grid = (unsigned)(((long long)A.nwarps * 32 + block - 1) / block);")
        # experiment: cap the resident CTAs per SM with unused dynamic shared memory (keeps the
        # local-memory footprint of the spilling classes inside L1)
        cap = int(os.environ.get("PC_GEN_OCC_CAP_L%d" % self.L, "0"))
        dyn = (200 * 1024 // cap) if cap else 0
        s.append("  switch (mode) {")
        for mode in MODES:
            pre = "pc_prefer_l1(eri_%s_kernel<%s>);" % (self.name, mode)
            if cap:
                pre = "cudaFuncSetAttribute(eri_%s_kernel<%s>, cudaFuncAttributeMaxDynamicSharedMemorySize, %d);" % (self.name, mode, dyn)
            s.append("    case %s: %s eri_%s_kernel<%s><<<grid, block, %d, st>>>(A); break;" % (mode, pre, self.name, mode, dyn))
        s.append("    default: return cudaErrorInvalidValue;")
        s.append("  }")
        s.append("  return cudaGetLastError();")
        s.append("}")
        return s

    def run_min_blocks(self):
        override = os.environ.get("PC_GEN_RUN_MINB_L%d" % self.L)
        if override:
            return int(override)
        return {0: 8, 1: 6, 2: 4, 3: 4}.get(self.L, 3)     # L = 2: 4 beats 5 since the bra records (profiles/r2d_*)

    def source_run(self):
        """Run form: one thread = one ket pair x a run of bra pairs sharing their primary shell
        (pc_plan).  The generation body is the same straight-line code, inside a warp-uniform loop
        over the run; symmetric-density digestion keeps the images without b in registers."""
        assert not self.V2
        lx1, ly1, lx2, ly2 = self.l
        vrr = self.gen_vrr()
        tail = self.gen_tail()
        NA, NB, NC, ND = self.nsph
        nsp = NA * NB * NC * ND
        nmax = max(self.La, self.Lc, 1)
        block = self.block_size()
        dims = "%d, %d, %d, %d" % (NA, NB, NC, ND)
        s = []
        s.append("// GENERATED by pychem_b200/codegen/gen_eri.py -- do not edit.")
        s.append("// class (%s%s|%s%s) [run form, up to %d bra pairs per thread]: L=%d, %d x %d contracted (e0|f0), %d VRR temporaries, %d tail temporaries"
                 % (LNAME[lx1], LNAME[ly1], LNAME[lx2], LNAME[ly2], self.run, self.L, self.ne, self.nf, self.n_vrr, self.n_tail))
        s.append('#include "../pc_common.cuh"')
        s.append("")
        s.append("namespace {")
        s.append("constexpr int L = %d, NE = %d, NF = %d, NSPH = %d;" % (self.L, self.ne, self.nf, nsp))
        s.extend(self.tables())
        s.append("")
        s.append("template <int MODE>")
        s.append("__global__ void __launch_bounds__(%d, %d) eri_%s_kernel(const __grid_constant__ PcEriArgs A) {" % (block, self.run_min_blocks(), self.name))
        s.append("  constexpr bool JK = (MODE >= PC_MODE_JK_RHF && MODE <= PC_MODE_JK_GEN) || MODE == PC_MODE_JK_GEN_BATCH;")
        s.append("  constexpr bool JKP = (MODE == PC_MODE_JK_RHF || MODE == PC_MODE_JK_UHF);   // resident images")
        s.append("  typedef PcSegScratch<PcSegNeed<MODE, %s, true>::ROWS> Scratch;" % dims)
        s.append("  __shared__ Scratch seg_scratch[%d];" % (block // 32))
        s.append("  Scratch* S = &seg_scratch[threadIdx.x >> 5];")
        s.append("  const int gw = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);")
        s.append("  if (gw >= A.nwarps) return;")
        s.append("  const PcItem& I = A.items[pc_find_item(A, gw)];")
        s.append("  const long long t = (long long)(gw - I.warp0) * 32 + (threadIdx.x & 31);")
        s.append("  int i0, rseg, j, seg_lo, seg_hi;")
        s.append("  bool forced;")
        s.append("  if (!pc_decode_run(A, I, t, i0, rseg, forced, j, seg_lo, seg_hi)) return;")
        s.append("  const int nb = I.bra.n, nk = I.ket.n, KK = __ldg(I.ket.keff + j);")
        s.append("  const double pk = forced ? 0.0 : __ldg(I.ket.pm + j);")
        s.append("  const double CD0 = __ldg(I.ket.xy + j), CD1 = __ldg(I.ket.xy + nk + j), CD2 = __ldg(I.ket.xy + 2 * nk + j);")
        s.append("  int fa0 = 0, fc = 0, fd = 0;")
        s.append("  if (JK) { fa0 = __ldg(I.bra.fx + i0); fc = __ldg(I.ket.fx + j); fd = __ldg(I.ket.fy + j); }")
        s.append("  const int kpid = JK ? __ldg(I.ket.pid + j) : 0;")
        s.append("  const int rmax = __reduce_max_sync(0xffffffffu, rseg);")
        s.append("  // the records of the run's bra pairs are consecutive: pull their first lines into L1 now, every")
        s.append("  // iteration of the run otherwise starts with an exposed L2 round trip for its header")
        s.append("  for (int r = 1; r < rseg; ++r) pc_prefetch_l1(reinterpret_cast<const double2*>(I.bra.rec) + (size_t)min(i0 + r, nb - 1) * (3 * (I.bra.K + 1)));")
        s.append("  PcRunAcc<JKP ? MODE : PC_MODE_JK_RHF, %s> RA;" % dims)
        s.append("  if (JKP) pc_run_init(A, RA, fa0, fc, fd);")
        s.append("#pragma unroll 1")
        s.append("  for (int r = 0; r < rmax; ++r) {")
        s.append("  const int i = min(i0 + r, nb - 1);")
        s.append("  const bool live = r < rseg;")
        self.bra_record(s, "i")
        s.append("  bool active = live;")
        s.append("  if (active && !forced) active = (bh1.y * pk > A.thresh) && (!I.same || i <= j);")
        s.append("  if (JK) pc_run_prefetch<%s>(A, fa0, fb, fc, fd);" % dims)
        s.append("  double g[NSPH];")
        s.append("  if (active) {")
        s.append("  double acc[NE * NF];")
        s.append("#pragma unroll")
        s.append("  for (int k = 0; k < NE * NF; ++k) acc[k] = 0.0;")
        self.prim_prologue(s, nmax)
        for line in vrr:
            s.append("      " + line)
        self.close_prim_loops(s)
        s.append("  (void)AB0; (void)AB1; (void)AB2;")
        for line in tail:
            s.append("  " + line)
        s.append("  if (!JK) pc_epilogue<MODE, %s>(A, I, t, true, bh2.x, fb, pidb, j, seg_lo, seg_hi, g, S);" % dims)
        s.append("  } else if (JK) {")
        s.append("#pragma unroll")
        s.append("    for (int k = 0; k < NSPH; ++k) g[k] = 0.0;")
        s.append("  }")
        s.append("  if (JK) {")
        s.append("    // shell-level degeneracy: 1/2 per a==b, c==d, (ab)==(cd)")
        s.append("    double fac = 1.0;")
        s.append("    if (fa0 == fb) fac *= 0.5;")
        s.append("    if (fc == fd) fac *= 0.5;")
        s.append("    if (pidb == kpid) fac *= 0.5;")
        s.append("    if (JKP) {")
        s.append("      if (fac != 1.0) {")
        s.append("#pragma unroll")
        s.append("        for (int k = 0; k < NSPH; ++k) g[k] *= fac;")
        s.append("      }")
        s.append("      pc_run_iter(A, RA, fa0, fb, fc, fd, g, active, live, seg_lo, seg_hi, S);")
        s.append("    } else {")
        s.append("      // general densities: all images per quartet (lanes without a quartet add zeros)")
        s.append("      pc_digest_any<MODE, %s>(A, fa0, fb, fc, fd, fac, g, active, live, seg_lo, seg_hi, S);" % dims)
        s.append("    }")
        s.append("  }")
        s.append("  }")
        s.append("  (void)CD0; (void)CD1; (void)CD2;")
        s.append("  if (JKP) pc_run_final(A, RA, fa0, fc, fd, rseg > 0, seg_lo, seg_hi, S);")
        s.append("}")
        s.append("}  // namespace")
        s.append("")
        s.extend(self.launcher(block))
        return "\n".join(s) + "\n"

    def source_single(self):
        lx1, ly1, lx2, ly2 = self.l
        vrr = self.gen_vrr()
        tail = self.gen_tail()
        NA, NB, NC, ND = self.nsph
        nsp = NA * NB * NC * ND
        nmax = max(self.La, self.Lc, 1)
        block = self.block_size()
        dims = "%d, %d, %d, %d" % (NA, NB, NC, ND)
        s = []
        s.append("// GENERATED by pychem_b200/codegen/gen_eri.py -- do not edit.")
        s.append("// class (%s%s|%s%s)%s: L=%d, %d x %d contracted (e0|f0), %d VRR temporaries, %d tail temporaries"
                 % (LNAME[lx1], LNAME[ly1], LNAME[lx2], LNAME[ly2], " [rolled form]" if self.V2 else "",
                    self.L, self.ne, self.nf, self.n_vrr, self.n_tail))
        s.append('#include "../pc_common.cuh"')
        if self.name in COOP_CLASSES and not self.cart_d:
            s.append('#include "../pc_async.cuh"')
        s.append("")
        s.append("namespace {")
        s.append("constexpr int L = %d, NE = %d, NF = %d, NSPH = %d;" % (self.L, self.ne, self.nf, nsp))
        s.extend(self.tables())
        s.append("")
        s.append("template <int MODE>")
        s.append("__global__ void __launch_bounds__(%d, %d) eri_%s_kernel(const __grid_constant__ PcEriArgs A) {" % (block, self.min_blocks(), self.name))
        s.append("  typedef PcSegScratch<PcSegNeed<MODE, %s, false>::ROWS> Scratch;" % dims)
        s.append("  __shared__ Scratch seg_scratch[%d];" % (block // 32))
        s.append("  Scratch* S = &seg_scratch[threadIdx.x >> 5];")
        s.append("  const int gw = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);")
        s.append("  if (gw >= A.nwarps) return;")
        s.append("  const PcItem& I = A.items[pc_find_item(A, gw)];")
        s.append("  const long long t = (long long)(gw - I.warp0) * 32 + (threadIdx.x & 31);")
        s.append("  int i, j, seg_lo, seg_hi;")
        s.append("  bool valid;")
        s.append("  if (!pc_decode_live(A, I, t, i, j, seg_lo, seg_hi, valid)) return;")
        self.bra_record(s, "i")
        s.append("  const int fa = bh2.x;")
        s.append("  // contraction depths after the primitive-pair cut-off (per shell pair)")
        s.append("  const int nk = I.ket.n, KK = __ldg(I.ket.keff + j);")
        s.append("  if (MODE >= PC_MODE_JK_RHF && MODE <= PC_MODE_JK_GEN)")
        s.append("    pc_prefetch_density<%s>(A, fa, fb, __ldg(I.ket.fx + j), __ldg(I.ket.fy + j), MODE != PC_MODE_JK_GEN);" % dims)
        s.append("  const double CD0 = __ldg(I.ket.xy + j), CD1 = __ldg(I.ket.xy + nk + j), CD2 = __ldg(I.ket.xy + 2 * nk + j);")
        self.emit_body(s, nmax, vrr, tail)
        s.append("  pc_epilogue<MODE, %s>(A, I, t, valid, fa, fb, pidb, j, seg_lo, seg_hi, g, S);" % dims)
        s.append("}")
        self.coop = None
        if self.name in COOP_CLASSES and not self.cart_d:
            self.coop = CoopGen(self, COOP_CLASSES[self.name])
            s.extend(self.coop.source())
        s.append("}  // namespace")
        s.append("")
        s.extend(self.launcher(block))
        return "\n".join(s) + "\n"

    def emit_body(self, s, nmax, vrr, tail):
        """primitive loops + contraction + tail: leaves g[NSPH]"""
        s.append("  double acc[NE * NF];")
        s.append("#pragma unroll%s" % (" 1" if self.V2 else ""))
        s.append("  for (int k = 0; k < NE * NF; ++k) acc[k] = 0.0;")
        s.extend(self.scratch_decls())
        self.prim_prologue(s, nmax)
        for line in vrr:
            s.append("      " + line)
        self.close_prim_loops(s)
        s.append("  (void)AB0; (void)AB1; (void)AB2; (void)CD0; (void)CD1; (void)CD2;")
        s.append("  double g[NSPH];")
        for line in tail:
            s.append("  " + line)

    def tables(self):
        return []

    def scratch_decls(self):
        return []


class ClassGenPass(ClassGen):
    """Several passes over the primitive loops, one per group of bra components (the grouping of
    CoopGen: chains of the bra recursion, dealt out by cost).  A pass keeps only ITS ne_p x nf
    contracted accumulators -- in registers -- and ends with the ket HRR + cart->spherical of its
    (e0| rows, written once per quartet to the per-thread array ks[]; the bra HRR, the bra
    cart->spherical and the digestion follow as in the one-pass form.  The one-pass kernels of these
    classes keep 144-279 accumulators in local memory and touch them in every primitive quartet
    ((dp|pp): 126 local stores + 151 loads per primitive quartet, 1.1 GB of DRAM writes per launch);
    here the price is the shared low end of the recursion, computed once per pass."""

    def __init__(self, *cls, **kw):
        self.npass = kw.pop("npass", 3)
        ClassGen.__init__(self, *cls, **kw)
        if isinstance(self.npass, str):           # "c96": as many passes as a cap of 96 accumulators per pass needs
            self.groups = CoopGen(self, 2, ket_split=False)
            self.groups.regroup_by_cap(int(self.npass[1:]))
        else:
            self.groups = CoopGen(self, self.npass, ket_split=False)

    def emit_body(self, s, nmax, vrr, tail):
        co = self.groups
        lx1, ly1, lx2, ly2 = self.l
        NA, NB, NC, ND = self.nsph
        nq = NC * ND
        s.append("  double ks[%d];      // (e0|cd): ket HRR + cart->spherical done, [e][q]" % (self.ne * nq))
        for r, es in enumerate(co.egroups):
            s.append("  {   // pass %d: bra components %s" % (r, " ".join("%d%d%d" % e for e in es)))
            s.append("  double acc[%d];" % (len(es) * self.nf))
            s.append("#pragma unroll")
            s.append("  for (int k = 0; k < %d; ++k) acc[k] = 0.0;" % (len(es) * self.nf))
            self.prim_prologue(s, nmax)
            for line in co.sub(es).gen_vrr():
                s.append("      " + line)
            self.close_prim_loops(s)
            for line in co.ket_tail(r, dest="ks[%d]", lane_stride=1):
                s.append("  " + line)
            s.append("  }")
        s.append("  (void)AB0; (void)AB1; (void)AB2; (void)CD0; (void)CD1; (void)CD2;")
        s.append("  double g[NSPH];")
        # bra HRR + cart->spherical for every ket function pair q, from ks[]
        em = Emit("hb")
        e_glob = {c: i for i, c in enumerate(self.e_list)}
        out = []
        for q in range(nq):
            memo = {}

            def hb(cx, cy):
                key = (cx, cy)
                if key in memo:
                    return memo[key]
                if sum(cy) == 0:
                    val = "ks[%d]" % (e_glob[cx] * nq + q)
                else:
                    d = first_dir(cy)
                    cy0 = dec(cy, d)
                    val = em.new("fma(AB%d, %s, %s)" % (d, hb(cx, cy0), hb(inc(cx, d), cy0)))
                memo[key] = val
                return val

            cart = {(ix, iy): hb(cx, cy) for ix, cx in enumerate(comps(lx1)) for iy, cy in enumerate(comps(ly1))}
            half = {}
            cd = self.cart_d
            for ix in range(ncart(lx1)):
                for my, row in enumerate(c2s_rows_any(ly1, cd)):
                    half[(ix, my)] = lin_comb(em, [(c, cart[(ix, iy)]) for iy, c in row])
            for mx, row in enumerate(c2s_rows_any(lx1, cd)):
                for my in range(nfun(ly1, cd)):
                    v = lin_comb(em, [(c, half[(ix, my)]) for ix, c in row])
                    out.append("g[%d] = %s;" % ((mx * nfun(ly1, cd) + my) * nq + q, v))
        for line in em.lines + out:
            s.append("  " + line)


class ClassGenV2(ClassGen):
    """Rolled form for the large classes.  Same recursion, same arithmetic per element, but
      * the bra build is a rolled loop over the ket Cartesian component (one code body per ket
        shell): the recursion over the bra components runs in registers for one ket component
        at a time; the one term that couples ket components, [a-1|c-1_i]^(m+1)
        (two_electron_vrr.c:104-108), is read from a small per-thread scratch array written
        while the previous ket shell was processed;
      * the ket HRR + cart->spherical is one code body in a rolled loop over the bra (e0| components,
        the bra HRR + cart->spherical one body in a rolled loop over the ket spherical components.
    The fully straight-line form of these classes (up to 8000 FMAs in one basic block with
    thousands of live values) makes ptxas fall back to spilling everything."""
    V2 = True

    def __init__(self, *cls, **kw):
        ClassGen.__init__(self, *cls, **kw)
        self.layout()

    def shell_needs(self):
        lx1, ly1, lx2, ly2 = self.l
        need = {}
        stack = [(la, lc, 0) for la in range(lx1, self.La + 1) for lc in range(lx2, self.Lc + 1)]
        while stack:
            la, lc, m = stack.pop()
            if la < 0 or lc < 0:
                continue
            if m in need.setdefault((la, lc), set()):
                continue
            need[(la, lc)].add(m)
            if la > 0:
                stack += [(la - 1, lc, m), (la - 1, lc, m + 1)]
                if la >= 2:
                    stack += [(la - 2, lc, m), (la - 2, lc, m + 1)]
                if lc >= 1:
                    stack.append((la - 1, lc - 1, m + 1))
            elif lc > 0:
                stack += [(0, lc - 1, m), (0, lc - 1, m + 1)]
                if lc >= 2:
                    stack += [(0, lc - 2, m), (0, lc - 2, m + 1)]
        return need

    def layout(self):
        self.need = self.shell_needs()
        La, Lc = self.La, self.Lc
        # values of ket shell lc that the NEXT ket shell reads through the cross term
        self.store = {}
        for lc in range(Lc):
            for la in range(La):
                src = self.need.get((la + 1, lc + 1), set())
                self.store[(la, lc)] = sorted(m + 1 for m in src)
        self.MS = {}
        for lc in range(Lc):
            mx = [max(self.store[(la, lc)]) for la in range(La) if self.store[(la, lc)]]
            self.MS[lc] = max(mx) if mx else 0
        self.nslot = ncum(La - 1)
        self.xs_size = max([self.nslot * ncart(lc) * self.MS[lc] for lc in range(Lc)] + [1])

    def tables(self):
        out = []
        for lc in range(1, self.Lc + 1):
            cm, cv = [], []
            for c in comps(lc):
                for d in range(3):
                    if c[d] > 0:
                        cm.append(cidx(dec(c, d)))
                        cv.append(float(c[d]))
                    else:
                        cm.append(0)
                        cv.append(0.0)
            out.append("__device__ const int CM%d[%d] = {%s};" % (lc, len(cm), ", ".join(str(x) for x in cm)))
            out.append("__device__ const double CV%d[%d] = {%s};" % (lc, len(cv), ", ".join(repr(x) for x in cv)))
        return out

    def scratch_decls(self):
        return ["  double KV[%d];" % (ncum(self.Lc) * (self.L + 1)),
                "  double XSA[%d], XSB[%d];" % (self.xs_size, self.xs_size)]

    def gen_vrr(self):
        lx1, ly1, lx2, ly2 = self.l
        La, Lc, L = self.La, self.Lc, self.L
        M1 = L + 1
        zero = (0, 0, 0)
        need = self.need
        lines = []
        self.vrr_refs = 0
        ntemp = 0
        # ---- ket build on the s bra, straight-line, results to KV[c*M1+m]
        em = Emit("k")
        kv = {}
        for m in sorted(need.get((0, 0), [])):
            kv[(zero, m)] = "F[%d]" % m
            em.raw("KV[%d] = F[%d];" % (m, m))
        for lc in range(1, Lc + 1):
            for c in comps(lc):
                d = first_dir(c)
                c0 = dec(c, d)
                n = c0[d]
                for m in sorted(need.get((0, lc), [])):
                    expr = "fma(QX%d, %s, Re%d * %s)" % (d, kv[(c0, m)], d, kv[(c0, m + 1)])
                    self.vrr_refs += 2 + (2 if n > 0 else 0)
                    if n > 0:
                        c1 = dec(c0, d)
                        expr = "fma(ne%d, fma(-eta, %s, %s), %s)" % (n, kv[(c1, m + 1)], kv[(c1, m)], expr)
                    v = em.new(expr)
                    kv[(c, m)] = v
                    em.raw("KV[%d] = %s;" % (cum(c) * M1 + m, v))
        ntemp += em.n
        lines += em.lines
        # ---- bra build, one rolled loop per ket shell
        e_index = {c: i for i, c in enumerate(self.e_list)}
        foff = {}
        o = 0
        for lf in range(lx2, Lc + 1):
            foff[lf] = o
            o += ncart(lf)
        for lc in range(0, Lc + 1):
            if not any(need.get((la, lc)) for la in range(1, La + 1)):
                continue
            cur = "XSA" if lc % 2 == 0 else "XSB"
            prev = "XSB" if lc % 2 == 0 else "XSA"
            NCc = ncart(lc)
            em = Emit("w%d_" % lc)
            lines.append("#pragma unroll 1")
            lines.append("for (int ic = 0; ic < %d; ++ic) {" % NCc)
            body = []
            V = {}
            for m in sorted(need.get((0, lc), [])):
                V[(zero, m)] = "b%d_%d" % (lc, m)
                body.append("const double b%d_%d = KV[(%d + ic) * %d + %d];" % (lc, m, ncum(lc - 1), M1, m))
            if lc >= 1:
                for d in range(3):
                    body.append("const int cs%d = CM%d[ic * 3 + %d] * %d;" % (d, lc, d, self.MS[lc - 1]))
                    body.append("const double cz%d = CV%d[ic * 3 + %d] * ze;" % (d, lc, d))
            for la in range(1, La + 1):
                ms = sorted(need.get((la, lc), []))
                if not ms:
                    continue
                for a in comps(la):
                    d = first_dir(a)
                    a0 = dec(a, d)
                    n = a0[d]
                    for m in ms:
                        expr = "fma(PX%d, %s, Rz%d * %s)" % (d, V[(a0, m)], d, V[(a0, m + 1)])
                        # executed once per ket component; the cross term exists for the
                        # components with c_d > 0 only (counted as in the straight-line form)
                        self.vrr_refs += NCc * (2 + (2 if n > 0 else 0)) + sum(1 for c in comps(lc) if c[d] > 0)
                        if n > 0:
                            a1 = dec(a0, d)
                            expr = "fma(nz%d, fma(-zeta, %s, %s), %s)" % (n, V[(a1, m + 1)], V[(a1, m)], expr)
                        if lc >= 1:
                            base = cum(a0) * ncart(lc - 1) * self.MS[lc - 1] + m      # (m+1) - 1
                            expr = "fma(cz%d, %s[%d + cs%d], %s)" % (d, prev, base, d, expr)
                        V[(a, m)] = em.new(expr)
            body += em.lines
            # stores for the next ket shell
            if lc < Lc:
                for la in range(0, La):
                    for a in comps(la):
                        for mp in self.store[(la, lc)]:
                            body.append("%s[%d + ic * %d] = %s;" % (cur, cum(a) * NCc * self.MS[lc] + (mp - 1),
                                                                    self.MS[lc], V[(a, mp)]))
            # contraction
            if lc >= lx2:
                for ie, e in enumerate(self.e_list):
                    body.append("acc[%d + ic] += %s;" % (ie * self.nf + foff[lc], V[(e, 0)]))
            lines += ["  " + b for b in body]
            lines.append("}")
            ntemp += em.n * NCc
        self.n_vrr = ntemp
        return lines

    def gen_tail(self):
        lx1, ly1, lx2, ly2 = self.l
        self.hrr_el = 0
        self.c2s_ops = 0
        f_index = {c: i for i, c in enumerate(self.f_list)}
        e_index = {c: i for i, c in enumerate(self.e_list)}
        nqs = nfun(lx2, self.cart_d) * nfun(ly2, self.cart_d)
        nps = nfun(lx1, self.cart_d) * nfun(ly1, self.cart_d)
        lines = ["double ks[%d];" % (self.ne * nqs)]
        # ---- ket HRR + cart->sph, rolled over the bra (e0| component
        em = Emit("hk")
        memo = {}

        def hk(cx, cy):
            key = (cx, cy)
            if key in memo:
                return memo[key]
            if sum(cy) == 0:
                val = "ar[%d]" % f_index[cx]
            else:
                d = first_dir(cy)
                cy0 = dec(cy, d)
                val = em.new("fma(CD%d, %s, %s)" % (d, hk(cx, cy0), hk(inc(cx, d), cy0)))
                self.hrr_el += self.ne
            memo[key] = val
            return val

        cart = {(ix, iy): hk(cx, cy) for ix, cx in enumerate(comps(lx2)) for iy, cy in enumerate(comps(ly2))}
        half = {}
        for ix in range(ncart(lx2)):
            for my, row in enumerate(c2s_rows_any(ly2, self.cart_d)):
                half[(ix, my)] = lin_comb(em, [(c, cart[(ix, iy)]) for iy, c in row])
        outs = []
        for mx, row in enumerate(c2s_rows_any(lx2, self.cart_d)):
            for my in range(nfun(ly2, self.cart_d)):
                v = lin_comb(em, [(c, half[(ix, my)]) for ix, c in row])
                outs.append("ks[ie * %d + %d] = %s;" % (nqs, mx * nfun(ly2, self.cart_d) + my, v))
        lines.append("#pragma unroll 1")
        lines.append("for (int ie = 0; ie < %d; ++ie) {" % self.ne)
        lines.append("  const double* __restrict__ ar = acc + ie * %d;" % self.nf)
        lines += ["  " + x for x in em.lines + outs]
        lines.append("}")
        ops1 = em.ops * self.ne
        n1 = em.n
        # ---- bra HRR + cart->sph, rolled over the ket spherical component
        em = Emit("hb")
        memo = {}

        def hb(cx, cy):
            key = (cx, cy)
            if key in memo:
                return memo[key]
            if sum(cy) == 0:
                val = "kr[%d]" % (e_index[cx] * nqs)
            else:
                d = first_dir(cy)
                cy0 = dec(cy, d)
                val = em.new("fma(AB%d, %s, %s)" % (d, hb(cx, cy0), hb(inc(cx, d), cy0)))
                self.hrr_el += nqs
            memo[key] = val
            return val

        cart = {(ix, iy): hb(cx, cy) for ix, cx in enumerate(comps(lx1)) for iy, cy in enumerate(comps(ly1))}
        half = {}
        for ix in range(ncart(lx1)):
            for my, row in enumerate(c2s_rows_any(ly1, self.cart_d)):
                half[(ix, my)] = lin_comb(em, [(c, cart[(ix, iy)]) for iy, c in row])
        outs = []
        for mx, row in enumerate(c2s_rows_any(lx1, self.cart_d)):
            for my in range(nfun(ly1, self.cart_d)):
                v = lin_comb(em, [(c, half[(ix, my)]) for ix, c in row])
                outs.append("g[%d + q] = %s;" % ((mx * nfun(ly1, self.cart_d) + my) * nqs, v))
        lines.append("#pragma unroll 1")
        lines.append("for (int q = 0; q < %d; ++q) {" % nqs)
        lines.append("  const double* __restrict__ kr = ks + q;")
        lines += ["  " + x for x in em.lines + outs]
        lines.append("}")
        self.c2s_ops = ops1 + em.ops * nqs
        self.n_tail = n1 + em.n
        return lines


class CoopGen:
    """Warp-cooperative form of a high-L class: the CTA is G warps that hold the SAME 32 quartets
    (lane = quartet, warp = role).  One thread per quartet keeps ne x nf contracted (e0|f0) and a
    few hundred recursion temporaries alive -- (dp|pp) spills 1.8 KB per thread, and ncu shows the
    spill lines travelling to L2 and DRAM (1.1 GB written per launch).  Here
      phase 1  role r runs the primitive loops for ITS bra components e (the chains of the bra
               recursion that end in its d/p/s ancestors: the bra build lowers one fixed direction
               per component, so the chains only share the low end), keeps ne_r x nf accumulators in
               registers, and finishes with the ket HRR + cart->spherical of its rows; the
               (e0|cd) rows go to shared memory [e][q][lane];
      phase 2  role r takes the ket functions c of ITS range: bra HRR + cart->spherical from shared
               memory, then the J/K digestion of the (a b | c_r d) sub-block.
    The redundant work is the shared low end of the recursion (x1.2-1.4 arithmetic in total); no
    accumulator leaves the registers.  Instantiated for the single-set J/K modes and NULL; the other
    modes (tensor / block output, scattering, batched density sets) stay on the one-thread kernel."""

    MODES = ("PC_MODE_JK_RHF", "PC_MODE_JK_UHF", "PC_MODE_JK_GEN", "PC_MODE_NULL")

    def __init__(self, base, G, ket_split=True, depth=0):
        self.b = base
        lx1, ly1, lx2, ly2 = base.l
        NA, NB, NC, ND = base.nsph
        split = os.environ.get("PC_GEN_COOP_SPLIT", "d")          # phase 2 by ket function d (default) or c
        if ND == 1:
            split = "c"
        self.G = G = min(G, (ND if split == "d" else NC) if ket_split else G,
                         sum(ncart(lx1 + k) for k in range(depth + 1)))
        self.nq = NC * ND

        def anc(e):
            # chain of e: its ancestor `depth` levels above the bra shell (0: the lx1-level component)
            while sum(e) > lx1 + depth:
                e = dec(e, first_dir(e))
            return e
        chains = {}
        for e in base.e_list:
            chains.setdefault(anc(e), []).append(e)

        def cost(es):
            return sum(l.count("fma(") + l.count(" * ") for l in self.sub(es).gen_vrr())
        bins = [[] for _ in range(G)]
        for _, es in sorted(chains.items(), key=lambda kv: -cost(kv[1])):
            best = min(range(G), key=lambda r: (cost(bins[r] + es), r))
            bins[best] = bins[best] + es
        order = {e: k for k, e in enumerate(base.e_list)}
        self.egroups = [sorted(b, key=order.get) for b in bins]
        # phase 2: role r digests the sub-block (a b | c0..c0+ncs-1, d0..d0+nds-1).  Split by d where
        # the ket's second shell has functions to split: the per-lane images J[c,d], K[a,d], K[b,d]
        # then belong to one role each (a split by c would repeat K[a,d] and K[b,d] in every role)
        self.sub_blocks = []
        tot = ND if split == "d" else NC
        x0 = 0
        for r in range(G):
            n = tot // G + (1 if r < tot % G else 0)
            self.sub_blocks.append((0, NC, x0, n) if split == "d" else (x0, n, 0, ND))
            x0 += n
        # with whole ket shells c in every role the segment-shared images J[a,b], K[a,c], K[b,c] of the
        # roles are partial sums of the same elements: left in shared memory and reduced by the CTA
        self.defer = split == "d" and os.environ.get("PC_GEN_COOP_DEFER", "1") != "0"
        self.max_acc = max(len(es) for es in self.egroups) * base.nf

    def regroup_by_cap(self, cap):
        """Groups of bra components with at most `cap` accumulators each: the lx1-level chains,
        those above the cap split into their next-level sub-chains, then merged greedily (largest
        first) into the group whose recursion grows least."""
        b = self.b
        lx1 = b.l[0]
        nf = b.nf

        def anc(e, depth):
            while sum(e) > lx1 + depth:
                e = dec(e, first_dir(e))
            return e
        chains = {}
        for e in b.e_list:
            chains.setdefault((anc(e, 0),), []).append(e)
        parts = []
        for key, es in chains.items():
            if len(es) * nf <= cap:
                parts.append(es)
                continue
            subs = {}
            for e in es:
                subs.setdefault(anc(e, 1), []).append(e)      # the lx1-level component is its own key
            root = subs.pop(key[0], [])
            subl = sorted(subs.values(), key=len)
            if subl:
                subl[0] = root + subl[0]
            else:
                subl = [root]
            parts.extend(subl)

        def cost(es):
            return sum(l.count("fma(") + l.count(" * ") for l in self.sub(es).gen_vrr())
        bins = []
        for es in sorted(parts, key=lambda x: -cost(x)):
            fits = [k for k, bn in enumerate(bins) if (len(bn) + len(es)) * nf <= cap]
            if fits:
                k = min(fits, key=lambda k: cost(bins[k] + es) - cost(bins[k]))
                bins[k] = bins[k] + es
            else:
                bins.append(list(es))
        order = {e: k for k, e in enumerate(b.e_list)}
        self.egroups = [sorted(bn, key=order.get) for bn in bins]
        self.G = len(bins)
        self.max_acc = max(len(es) for es in self.egroups) * nf

    def sub(self, es):
        g = ClassGen(*self.b.l, cart_d=self.b.cart_d)
        g.e_list = list(es)
        g.ne = len(es)
        return g

    def ket_tail(self, r, dest="ks_sm[%d + lane]", lane_stride=32):
        """ket HRR + cart->spherical of role r's (e0| rows; results to shared memory (or `dest`)."""
        b = self.b
        lx1, ly1, lx2, ly2 = b.l
        em = Emit("hk")
        f_index = {c: i for i, c in enumerate(b.f_list)}
        e_glob = {c: i for i, c in enumerate(b.e_list)}
        out = []
        for il, e in enumerate(self.egroups[r]):
            memo = {}

            def hk(cx, cy):
                key = (cx, cy)
                if key in memo:
                    return memo[key]
                if sum(cy) == 0:
                    val = "acc[%d]" % (il * b.nf + f_index[cx])
                else:
                    d = first_dir(cy)
                    cy0 = dec(cy, d)
                    val = em.new("fma(CD%d, %s, %s)" % (d, hk(cx, cy0), hk(inc(cx, d), cy0)))
                memo[key] = val
                return val

            cart = {(ix, iy): hk(cx, cy) for ix, cx in enumerate(comps(lx2)) for iy, cy in enumerate(comps(ly2))}
            half = {}
            cd = b.cart_d
            for ix in range(ncart(lx2)):
                for my, row in enumerate(c2s_rows_any(ly2, cd)):
                    half[(ix, my)] = lin_comb(em, [(c, cart[(ix, iy)]) for iy, c in row])
            for mx, row in enumerate(c2s_rows_any(lx2, cd)):
                for my in range(nfun(ly2, cd)):
                    v = lin_comb(em, [(c, half[(ix, my)]) for ix, c in row])
                    q = mx * nfun(ly2, cd) + my
                    out.append("%s = %s;" % (dest % ((e_glob[e] * self.nq + q) * lane_stride), v))
        return em.lines + out

    def bra_tail(self, r):
        """bra HRR + cart->spherical for the ket functions c of role r: g[(a b)][c_local][d]."""
        b = self.b
        lx1, ly1, lx2, ly2 = b.l
        NA, NB, NC, ND = b.nsph
        c0, ncs, d0, nds = self.sub_blocks[r]
        em = Emit("hb")
        e_glob = {c: i for i, c in enumerate(b.e_list)}
        out = []
        nqs = ncs * nds
        for cl in range(ncs):
            for dl in range(nds):
                q = (c0 + cl) * ND + d0 + dl
                ql = cl * nds + dl
                memo = {}

                def hb(cx, cy):
                    key = (cx, cy)
                    if key in memo:
                        return memo[key]
                    if sum(cy) == 0:
                        val = em.new("ks_sm[%d + lane]" % ((e_glob[cx] * self.nq + q) * 32))
                    else:
                        d = first_dir(cy)
                        cy0 = dec(cy, d)
                        val = em.new("fma(AB%d, %s, %s)" % (d, hb(cx, cy0), hb(inc(cx, d), cy0)))
                    memo[key] = val
                    return val

                cart = {(ix, iy): hb(cx, cy) for ix, cx in enumerate(comps(lx1)) for iy, cy in enumerate(comps(ly1))}
                half = {}
                for ix in range(ncart(lx1)):
                    for my, row in enumerate(c2s_rows(ly1)):
                        half[(ix, my)] = lin_comb(em, [(c, cart[(ix, iy)]) for iy, c in row])
                for mx, row in enumerate(c2s_rows(lx1)):
                    for my in range(nsph(ly1)):
                        v = lin_comb(em, [(c, half[(ix, my)]) for ix, c in row])
                        pidx = mx * nsph(ly1) + my
                        out.append("g[%d] = %s;" % (pidx * nqs + ql, v))
        return em.lines + out

    def min_blocks(self):
        override = os.environ.get("PC_GEN_COOP_MINB")
        if override:
            return int(override)
        return 2

    def smem_bytes(self, mode):
        NA, NB, NC, ND = self.b.nsph
        n = self.b.ne * self.nq * 32
        if self.defer and mode != "PC_MODE_NULL":
            gen = mode == "PC_MODE_JK_GEN"
            ks = NA * NC + NB * NC + (NC * NB + NC * NA if gen else 0)
            n += self.G * (NA * NB + (1 if mode == "PC_MODE_JK_RHF" else 2) * ks) * 33
        return n * 8

    def source(self):
        """kernel text (inside the class file's anonymous namespace)"""
        b = self.b
        G = self.G
        NA, NB, NC, ND = b.nsph
        nmax = max(b.La, b.Lc, 1)
        big = max(self.sub_blocks, key=lambda sb: sb[1] * sb[3])
        s = []
        s.append("")
        s.append("// ---- warp-cooperative form: %d roles; bra components per role %s; ket functions per role %s"
                 % (G, [len(e) for e in self.egroups], ["%dx%d" % (sb[1], sb[3]) for sb in self.sub_blocks]))
        s.append("template <int MODE>")
        maxnreg = int(os.environ.get("PC_GEN_COOP_MAXNREG", "0"))
        bounds = "PC_MAXNREG(%d)" % maxnreg if maxnreg else "__launch_bounds__(%d, %d)" % (32 * G, self.min_blocks())
        s.append("__global__ void %s eri_%s_coop_kernel(const __grid_constant__ PcEriArgs A) {" % (bounds, b.name))
        s.append("  typedef PcSegScratch<PcSegNeed<MODE, %d, %d, %d, %d, false>::ROWS> Scratch;" % (NA, NB, big[1], big[3]))
        s.append("  __shared__ Scratch seg_scratch[%d];" % G)
        s.append("  PC_DYN_SMEM(ks_raw);")
        s.append("  double* __restrict__ ks_sm = reinterpret_cast<double*>(ks_raw);      // [e][q][lane]")
        if self.defer:
            s.append("  typedef PcCoopCount<MODE, %d, %d, %d, %d> Coop;" % (NA, NB, NC, G))
            s.append("  double* __restrict__ part = ks_sm + %d;                          // [role][value][33]" % (b.ne * self.nq * 32))
            s.append("  __shared__ PcCoopSeg coop_seg;")
        s.append("  const int role = threadIdx.x >> 5, lane = threadIdx.x & 31;")
        s.append("  Scratch* S = &seg_scratch[role];")
        s.append("  const int gw = blockIdx.x;                                           // the CTA's 32 quartets")
        s.append("  const PcItem& I = A.items[pc_find_item(A, gw)];")
        s.append("  const long long t = (long long)(gw - I.warp0) * 32 + lane;")
        s.append("  int i, j, seg_lo, seg_hi;")
        s.append("  bool valid;")
        s.append("  if (!pc_decode_live(A, I, t, i, j, seg_lo, seg_hi, valid)) return;   // same answer in every role")
        b.bra_record(s, "i")
        s.append("  const int fa = bh2.x;")
        s.append("  const int nk = I.ket.n, KK = __ldg(I.ket.keff + j);")
        if os.environ.get("PC_GEN_COOP_PREFETCH", "1") != "0":
            s.append("  if (role == %d && MODE >= PC_MODE_JK_RHF && MODE <= PC_MODE_JK_GEN)   // one role warms L1 for the CTA's digestion" % (G - 1))
            s.append("    pc_prefetch_density<%d, %d, %d, %d>(A, fa, fb, __ldg(I.ket.fx + j), __ldg(I.ket.fy + j), MODE != PC_MODE_JK_GEN);" % (NA, NB, NC, ND))
        s.append("  const double CD0 = __ldg(I.ket.xy + j), CD1 = __ldg(I.ket.xy + nk + j), CD2 = __ldg(I.ket.xy + 2 * nk + j);")
        s.append("  {")
        s.append("  double acc[%d];" % self.max_acc)
        s.append("#pragma unroll")
        s.append("  for (int k = 0; k < %d; ++k) acc[k] = 0.0;" % self.max_acc)
        b.prim_prologue(s, nmax)
        s.append("      switch (role) {")
        for r in range(G):
            s.append("      case %d: {" % r)
            for line in self.sub(self.egroups[r]).gen_vrr():
                s.append("        " + line)
            s.append("      } break;")
        s.append("      }")
        b.close_prim_loops(s)
        s.append("  (void)CD0; (void)CD1; (void)CD2;")
        s.append("  switch (role) {")
        for r in range(G):
            s.append("  case %d: {" % r)
            for line in self.ket_tail(r):
                s.append("    " + line)
            s.append("  } break;")
        s.append("  }")
        s.append("  }")
        s.append("  __syncthreads();")
        s.append("  (void)AB0; (void)AB1; (void)AB2;")
        s.append("  switch (role) {")
        for r in range(G):
            c0, ncs, d0, nds = self.sub_blocks[r]
            s.append("  case %d: {" % r)
            s.append("    double g[%d];" % (NA * NB * ncs * nds))
            for line in self.bra_tail(r):
                s.append("    " + line)
            if self.defer:
                s.append("    pc_epilogue_sub<MODE, %d, %d, %d, %d>(A, I, valid, fa, fb, pidb, j, %d, %d, seg_lo, seg_hi, g, S, PcReduceDefer{part + %d * Coop::NVT * 33});"
                         % (NA, NB, ncs, nds, c0, d0, r))
            else:
                s.append("    pc_epilogue_sub<MODE, %d, %d, %d, %d>(A, I, valid, fa, fb, pidb, j, %d, %d, seg_lo, seg_hi, g, S);"
                         % (NA, NB, ncs, nds, c0, d0))
            s.append("  } break;")
        s.append("  }")
        if self.defer:
            s.append("  if (MODE != PC_MODE_NULL) {")
            s.append("    const unsigned ends = __ballot_sync(0xffffffffu, valid && lane == seg_hi);")
            s.append("    if (role == 0) coop_seg.idx[lane] = make_int4(fa, fb, __ldg(I.ket.fx + j), seg_lo);")
            s.append("    __syncthreads();")
            s.append("    pc_coop_reduce<MODE, %d, %d, %d, %d>(A, part, &coop_seg, ends);" % (NA, NB, NC, G))
            s.append("  }")
        s.append("}")
        return s

    def launch_lines(self):
        b = self.b
        s = []
        s.append("  static const bool coop = []() { const char* e = getenv(\"PYCHEM_B200_COOP\"); return !(e && e[0] == '0'); }();")
        s.append("  if (coop) {")
        s.append("    switch (mode) {")
        for mode in MODES:
            if mode in self.MODES:
                s.append("      case %s: {" % mode)
                s.append("        static bool attr = false;")
                s.append("        if (!attr) { attr = true; cudaFuncSetAttribute(eri_%s_coop_kernel<%s>, cudaFuncAttributeMaxDynamicSharedMemorySize, %d); }"
                         % (b.name, mode, self.smem_bytes(mode)))
                s.append("        eri_%s_coop_kernel<%s><<<(unsigned)A.nwarps, %d, %d, st>>>(A);" % (b.name, mode, 32 * self.G, self.smem_bytes(mode)))
                s.append("        return cudaGetLastError();")
                s.append("      }")
        s.append("      default: break;")
        s.append("    }")
        s.append("  }")
        return s


# classes built in the multi-pass form (ClassGenPass) and their number of passes: measured on
# (H2O)32 (profiles/r2g_multipass_high_l.txt); (dd|ps), (pp|pp), (dp|ps) and (dd|dd) lose with it
PASS_CLASSES = {"dppp": 2, "dpdp": 3, "dpds": 2, "ddds": 3, "ddpp": 4, "dddp": 6}
if os.environ.get("PC_GEN_PASS") is not None:       # name=passes or name=c<accumulator cap per pass>
    PASS_CLASSES = dict((kv.split("=")[0], kv.split("=")[1] if kv.split("=")[1].startswith("c") else int(kv.split("=")[1]))
                        for kv in os.environ["PC_GEN_PASS"].split(",") if kv and kv != "0")

# classes built in the warp-cooperative form as well, and their number of roles
COOP_CLASSES = {}
if os.environ.get("PC_GEN_COOP"):
    COOP_CLASSES = dict((kv.split("=")[0], int(kv.split("=")[1])) for kv in os.environ["PC_GEN_COOP"].split(",") if kv and kv != "0")


# bra-record prefetch: 0 = none (the record's lines hit L1 after the first touch), 1 = next primitive
# one iteration ahead in registers, 2 = prefetch.global.L1 of the record's lines at task start
PREFETCH = int(os.environ.get("PC_GEN_PREFETCH", "0"))
# experiment: unroll the bra-primitive loop (the loads of the next primitive can be scheduled early)
UNROLL_IB = int(os.environ.get("PC_GEN_UNROLL_IB", "1"))
UNROLL_IB_MAXL = int(os.environ.get("PC_GEN_UNROLL_IB_MAXL", "3"))
FUSE_ACC = os.environ.get("PC_GEN_FUSE_ACC", "1") != "0"
V2_THRESHOLD = int(os.environ.get("PC_GEN_V2_THRESHOLD", "650"))   # classes with more VRR temporaries than this use the rolled form


# classes compiled in the run form and the longest run they take (the plan may use shorter ones,
# PYCHEM_B200_RUN at run time); chosen where digestion is a large share of the class's time and the
# resident images fit the register budget
RUN_CLASSES = {"ssss": 4, "psss": 4, "psps": 4, "dsss": 4, "dsps": 4, "dspp": 4, "dsds": 4}
if os.environ.get("PC_GEN_RUN_CLASSES") is not None:
    RUN_CLASSES = dict((kv.split("=")[0], int(kv.split("=")[1]))
                       for kv in os.environ["PC_GEN_RUN_CLASSES"].split(",") if kv)


# output modes instantiated per class kernel; timing variants (tools/build_variant.py) restrict them
# with PC_GEN_MODES=0,2,5 (BLOCKS for the Schwarz pass, JK_RHF, NULL): a third of the compile time
ALL_MODES = ("PC_MODE_BLOCKS", "PC_MODE_TENSOR", "PC_MODE_JK_RHF", "PC_MODE_JK_UHF", "PC_MODE_JK_GEN", "PC_MODE_NULL",
             "PC_MODE_BLOCKS_SCAT", "PC_MODE_TENSOR_SCAT", "PC_MODE_JK_GEN_BATCH")
MODES = ALL_MODES
if os.environ.get("PC_GEN_MODES"):
    MODES = tuple(ALL_MODES[int(k)] for k in os.environ["PC_GEN_MODES"].split(","))

# A/B variant builds (tools/build_variant.py) skip the Cartesian-d kernels: the dispatch table then
# points at the spherical ones, which is wrong for Cartesian_L = [2] molecules and fine for timing
SKIP_CART = os.environ.get("PC_GEN_SKIP_CART", "0") != "0"


def make_class(cls, cart_d=False):
    g = ClassGen(*cls, cart_d=cart_d)
    if g.name.lower() in PASS_CLASSES:        # the Cartesian-d variant (D in the name) of a class takes its form
        return ClassGenPass(*cls, cart_d=cart_d, npass=PASS_CLASSES[g.name.lower()])
    g.gen_vrr()
    if g.n_vrr > V2_THRESHOLD:
        return ClassGenV2(*cls, cart_d=cart_d)
    g = ClassGen(*cls, cart_d=cart_d)
    if not (cart_d and 2 in cls):
        g.run = RUN_CLASSES.get(g.name, 1)
    return g


def flop_model(g):
    """SURVEY.md section 8(d) flop model evaluated on this generator's recursion DAG:
    per primitive quartet 9(L+1)+12 (Boys cubic + scaling) + 2*vrr_refs + 2*contracted elements;
    per contracted quartet 2*hrr elements + ncart + 2*c2s operations."""
    ncart_tot = 1
    for l in g.l:
        ncart_tot *= ncart(l)
    prim = 9 * (g.L + 1) + 12 + 2 * g.vrr_refs + 2 * g.ne * g.nf
    cont = 2 * g.hrr_el + ncart_tot + 2 * g.c2s_ops
    nsp = 1
    for x in g.nsph:
        nsp *= x
    return {"L": g.L, "vrr_refs": g.vrr_refs, "vrr_elems": g.n_vrr, "contr_elems": g.ne * g.nf,
            "hrr_elems": g.hrr_el, "c2s_ops": g.c2s_ops, "nsph": nsp,
            "flop_prim": prim, "flop_cont": cont}


def all_classes():
    out = []
    for ib, b in enumerate(PAIR_CLASSES):
        for k in PAIR_CLASSES[:ib + 1]:
            out.append(b + k)
    return out


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    names = []
    model = {}
    gens = {}
    for cart in (False, True):
        for cls in all_classes():
            if cart and (2 not in cls or SKIP_CART):
                continue                      # no d shell: the spherical kernel is the kernel
            g = make_class(cls, cart_d=cart)
            path = os.path.join(outdir, "eri_%s.cu" % g.name)
            src = g.source()
            if not cart:
                model[g.name] = flop_model(g)
            if not os.path.exists(path) or open(path).read() != src:
                with open(path, "w") as fh:
                    fh.write(src)
            names.append(g.name)
            gens[(cart, cls)] = g
            print("%s: %d vrr temps, %d tail temps, %d lines" % (g.name, g.n_vrr, g.n_tail, src.count("\n")))
    # dispatch tables
    tab = ['// GENERATED by pychem_b200/codegen/gen_eri.py -- do not edit.', '#include "../pc_common.cuh"']
    for n in names:
        tab.append("cudaError_t pc_launch_%s(int mode, const PcEriArgs& A, cudaStream_t st);" % n)
    tab.append("// index [d shells Cartesian?][bra pair class][ket pair class], pair classes ss ps pp ds dp dd, bra >= ket")
    tab.append("pc_launch_fn pc_launch_table[2][6][6] = {")
    for cart in (False, True):
        tab.append(" {")
        for ib, b in enumerate(PAIR_CLASSES):
            row = []
            for ik, k in enumerate(PAIR_CLASSES):
                if ik <= ib:
                    g = gens.get((cart, b + k)) or gens[(False, b + k)]
                    row.append("pc_launch_%s" % g.name)
                else:
                    row.append("nullptr")
            tab.append("  {" + ", ".join(row) + "},")
        tab.append(" },")
    tab.append("};")
    for key, label in (("flop_prim", "pc_flop_prim_table"), ("flop_cont", "pc_flop_cont_table")):
        tab.append("double %s[6][6] = {" % label)
        for ib, b in enumerate(PAIR_CLASSES):
            row = []
            for ik, k in enumerate(PAIR_CLASSES):
                row.append(str(float(model["".join(LNAME[x] for x in b + k)][key])) if ik <= ib else "0.0")
            tab.append("  {" + ", ".join(row) + "},")
        tab.append("};")
    tab.append("int pc_run_table[2][6][6] = {")
    for cart in (False, True):
        tab.append(" {")
        for ib, b in enumerate(PAIR_CLASSES):
            row = []
            for ik, k in enumerate(PAIR_CLASSES):
                if ik <= ib:
                    g = gens.get((cart, b + k)) or gens[(False, b + k)]
                    row.append(str(g.run))
                else:
                    row.append("0")
            tab.append("  {" + ", ".join(row) + "},")
        tab.append(" },")
    tab.append("};")
    tab.append("int pc_block_table[2][6][6] = {")
    for cart in (False, True):
        tab.append(" {")
        for ib, b in enumerate(PAIR_CLASSES):
            row = []
            for ik, k in enumerate(PAIR_CLASSES):
                if ik <= ib:
                    g = gens.get((cart, b + k)) or gens[(False, b + k)]
                    row.append(str(g.block_size()))
                else:
                    row.append("0")
            tab.append("  {" + ", ".join(row) + "},")
        tab.append(" },")
    tab.append("};")
    import json
    mpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "data", "flop_model.json")
    msrc = json.dumps(model, indent=1, sort_keys=True) + "\n"
    if not os.path.exists(mpath) or open(mpath).read() != msrc:
        with open(mpath, "w") as fh:
            fh.write(msrc)
    path = os.path.join(outdir, "dispatch.cu")
    src = "\n".join(tab) + "\n"
    if not os.path.exists(path) or open(path).read() != src:
        with open(path, "w") as fh:
            fh.write(src)
    # stale files of classes that no longer exist would still be compiled: remove them
    keep = set("eri_%s.cu" % n for n in names) | {"dispatch.cu"}
    for f in os.listdir(outdir):
        if f.endswith(".cu") and f not in keep:
            os.remove(os.path.join(outdir, f))
    return names


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "csrc", "gen"))
