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
    def __init__(self, lx1, ly1, lx2, ly2):
        self.l = (lx1, ly1, lx2, ly2)
        self.La, self.Lc = lx1 + ly1, lx2 + ly2
        self.L = self.La + self.Lc
        self.name = "".join(LNAME[x] for x in self.l)
        # contracted (e0|f0): e = lx1..La, f = lx2..Lc, all components
        self.e_list = [c for le in range(lx1, self.La + 1) for c in comps(le)]
        self.f_list = [c for lf in range(lx2, self.Lc + 1) for c in comps(lf)]
        self.ne, self.nf = len(self.e_list), len(self.f_list)
        self.nsph = [nsph(x) for x in self.l]

    # ------------------------------------------------------------------ VRR
    def gen_vrr(self):
        em = Emit("v")
        memo = {}
        zero = (0, 0, 0)
        self.vrr_refs = 0

        def get(a, c, m):
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
                expr = "fma(QX%d, %s, Re%d * %s)" % (d, b0, d, b1)
                self.vrr_refs += 2 + (2 if n > 0 else 0)
                if n > 0:
                    c1 = dec(c0, d)
                    b2, b3 = get(zero, c1, m), get(zero, c1, m + 1)
                    expr = "fma(ne%d, fma(-eta, %s, %s), %s)" % (n, b3, b2, expr)
                val = em.new(expr)
            else:
                d = first_dir(a)
                a0 = dec(a, d)
                n = a0[d]
                b0, b1 = get(a0, c, m), get(a0, c, m + 1)
                expr = "fma(PX%d, %s, Rz%d * %s)" % (d, b0, d, b1)
                self.vrr_refs += 2 + (2 if n > 0 else 0) + (1 if c[d] > 0 else 0)
                if n > 0:
                    a1 = dec(a0, d)
                    b2, b3 = get(a1, c, m), get(a1, c, m + 1)
                    expr = "fma(nz%d, fma(-zeta, %s, %s), %s)" % (n, b3, b2, expr)
                if c[d] > 0:
                    b4 = get(a0, dec(c, d), m + 1)
                    expr = "fma(nze%d, %s, %s)" % (c[d], b4, expr)
                val = em.new(expr)
            memo[key] = val
            return val

        for ie, e in enumerate(self.e_list):
            for jf, f in enumerate(self.f_list):
                v = get(e, f, 0)
                em.raw("acc[%d] += %s;" % (ie * self.nf + jf, v))
        self.n_vrr = em.n
        return em.lines

    # ------------------------------------------------------------------ HRR + c2s
    def gen_tail(self):
        lx1, ly1, lx2, ly2 = self.l
        em = Emit("h")
        self.hrr_el = 0
        self.c2s_ops = 0
        f_index = {c: i for i, c in enumerate(self.f_list)}
        e_index = {c: i for i, c in enumerate(self.e_list)}
        nqs = nsph(lx2) * nsph(ly2)

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
                for my, row in enumerate(c2s_rows(ly2)):
                    half[(ix, my)] = lin_comb(em, [(c, cart[(ix, iy)]) for iy, c in row])
            for mx, row in enumerate(c2s_rows(lx2)):
                for my in range(nsph(ly2)):
                    ks[(ie, mx * nsph(ly2) + my)] = lin_comb(em, [(c, half[(ix, my)]) for ix, c in row])

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
                for my, row in enumerate(c2s_rows(ly1)):
                    half[(ix, my)] = lin_comb(em, [(c, cart[(ix, iy)]) for iy, c in row])
            for mx, row in enumerate(c2s_rows(lx1)):
                for my in range(nsph(ly1)):
                    v = lin_comb(em, [(c, half[(ix, my)]) for ix, c in row])
                    p = mx * nsph(ly1) + my
                    out.append("g[%d] = %s;" % (p * nqs + q, v))
        self.n_tail = em.n
        self.c2s_ops = em.ops
        return em.lines + out

    # ------------------------------------------------------------------ file
    def source(self):
        lx1, ly1, lx2, ly2 = self.l
        vrr = self.gen_vrr()
        tail = self.gen_tail()
        NA, NB, NC, ND = self.nsph
        nsp = NA * NB * NC * ND
        nmax = max(self.La, self.Lc, 1)
        block = 128 if self.L <= 4 else 64
        s = []
        s.append("// GENERATED by pychem_b200/codegen/gen_eri.py -- do not edit.")
        s.append("// class (%s%s|%s%s): L=%d, %d x %d contracted (e0|f0), %d VRR temporaries, %d tail temporaries"
                 % (LNAME[lx1], LNAME[ly1], LNAME[lx2], LNAME[ly2], self.L, self.ne, self.nf, self.n_vrr, self.n_tail))
        s.append('#include "../pc_common.cuh"')
        s.append("")
        s.append("namespace {")
        s.append("constexpr int L = %d, NE = %d, NF = %d, NSPH = %d;" % (self.L, self.ne, self.nf, nsp))
        s.append("")
        s.append("template <int MODE>")
        s.append("__global__ void __launch_bounds__(%d) eri_%s_kernel(const PcEriArgs A) {" % (block, self.name))
        s.append("  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;")
        s.append("  int i, j;")
        s.append("  if (!pc_decode_task(A, t, i, j)) return;")
        s.append("  const int nb = A.bra.n, nk = A.ket.n, KB = A.bra.K, KK = A.ket.K;")
        s.append("  const double AB0 = A.bra.xy[i], AB1 = A.bra.xy[nb + i], AB2 = A.bra.xy[2 * nb + i];")
        s.append("  const double CD0 = A.ket.xy[j], CD1 = A.ket.xy[nk + j], CD2 = A.ket.xy[2 * nk + j];")
        s.append("  double acc[NE * NF];")
        s.append("#pragma unroll")
        s.append("  for (int k = 0; k < NE * NF; ++k) acc[k] = 0.0;")
        s.append("  const double* __restrict__ bp = A.bra.prim + i;")
        s.append("  const double* __restrict__ kp = A.ket.prim + j;")
        s.append("  const size_t sb = (size_t)KB * nb, sk = (size_t)KK * nk;")
        s.append("  for (int ib = 0; ib < KB; ++ib) {")
        s.append("    const double sP = bp[(size_t)ib * nb], UP = bp[sb + (size_t)ib * nb];")
        s.append("    const double Px = bp[2 * sb + (size_t)ib * nb], Py = bp[3 * sb + (size_t)ib * nb], Pz = bp[4 * sb + (size_t)ib * nb];")
        s.append("    const double kzP = bp[5 * sb + (size_t)ib * nb];")
        s.append("    const double zeta = 0.5 * sP;")
        s.append("    const double PX0 = -AB0 * kzP, PX1 = -AB1 * kzP, PX2 = -AB2 * kzP;")
        for n in range(1, nmax + 1):
            s.append("    const double nz%d = %d.0 * zeta;" % (n, n))
        s.append("    for (int ik = 0; ik < KK; ++ik) {")
        s.append("      const double sQ = kp[(size_t)ik * nk], UQ = kp[sk + (size_t)ik * nk];")
        s.append("      const double Qx = kp[2 * sk + (size_t)ik * nk], Qy = kp[3 * sk + (size_t)ik * nk], Qz = kp[4 * sk + (size_t)ik * nk];")
        s.append("      const double kzQ = kp[5 * sk + (size_t)ik * nk];")
        s.append("      const double eta = 0.5 * sQ;")
        s.append("      const double QX0 = -CD0 * kzQ, QX1 = -CD1 * kzQ, QX2 = -CD2 * kzQ;")
        s.append("      const double R0 = Px - Qx, R1 = Py - Qy, R2_ = Pz - Qz;")
        s.append("      const double Rsq = R0 * R0 + R1 * R1 + R2_ * R2_;")
        s.append("      double F[L + 1];")
        s.append("      pc_fundamentals<L>(sP, UP, sQ, UQ, Rsq, A.boys, F);")
        s.append("      const double Rz0 = -R0 * zeta, Rz1 = -R1 * zeta, Rz2 = -R2_ * zeta;")
        s.append("      const double Re0 = R0 * eta, Re1 = R1 * eta, Re2 = R2_ * eta;")
        s.append("      const double ze = zeta * eta;")
        for n in range(1, nmax + 1):
            s.append("      const double ne%d = %d.0 * eta; const double nze%d = %d.0 * ze;" % (n, n, n, n))
        s.append("      (void)QX0; (void)QX1; (void)QX2; (void)PX0; (void)PX1; (void)PX2; (void)ze;")
        s.append("      (void)Rz0; (void)Rz1; (void)Rz2; (void)Re0; (void)Re1; (void)Re2;")
        for line in vrr:
            s.append("      " + line)
        s.append("    }")
        s.append("  }")
        s.append("  (void)AB0; (void)AB1; (void)AB2; (void)CD0; (void)CD1; (void)CD2;")
        s.append("  double g[NSPH];")
        for line in tail:
            s.append("  " + line)
        s.append("  pc_epilogue<MODE, %d, %d, %d, %d>(A, t, i, j, g);" % (NA, NB, NC, ND))
        s.append("}")
        s.append("}  // namespace")
        s.append("")
        s.append("cudaError_t pc_launch_%s(int mode, const PcEriArgs& A, cudaStream_t st) {" % self.name)
        s.append("  if (A.t_count <= 0) return cudaSuccess;")
        s.append("  const int block = %d;" % block)
        s.append("  const unsigned grid = (unsigned)((A.t_count + block - 1) / block);")
        s.append("  switch (mode) {")
        for mode in ("PC_MODE_BLOCKS", "PC_MODE_TENSOR", "PC_MODE_JK_RHF", "PC_MODE_JK_UHF", "PC_MODE_JK_GEN"):
            s.append("    case %s: eri_%s_kernel<%s><<<grid, block, 0, st>>>(A); break;" % (mode, self.name, mode))
        s.append("    default: return cudaErrorInvalidValue;")
        s.append("  }")
        s.append("  return cudaGetLastError();")
        s.append("}")
        return "\n".join(s) + "\n"


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
    for cls in all_classes():
        g = ClassGen(*cls)
        path = os.path.join(outdir, "eri_%s.cu" % g.name)
        src = g.source()
        model[g.name] = flop_model(g)
        if not os.path.exists(path) or open(path).read() != src:
            with open(path, "w") as fh:
                fh.write(src)
        names.append(g.name)
        print("%s: %d vrr temps, %d tail temps, %d lines" % (g.name, g.n_vrr, g.n_tail, src.count("\n")))
    # dispatch table
    tab = ['// GENERATED by pychem_b200/codegen/gen_eri.py -- do not edit.', '#include "../pc_common.cuh"']
    for n in names:
        tab.append("cudaError_t pc_launch_%s(int mode, const PcEriArgs& A, cudaStream_t st);" % n)
    tab.append("// index [bra pair class][ket pair class], pair classes ss ps pp ds dp dd, bra >= ket")
    tab.append("pc_launch_fn pc_launch_table[6][6] = {")
    for ib, b in enumerate(PAIR_CLASSES):
        row = []
        for ik, k in enumerate(PAIR_CLASSES):
            row.append("pc_launch_%s" % "".join(LNAME[x] for x in b + k) if ik <= ib else "nullptr")
        tab.append("  {" + ", ".join(row) + "},")
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
    return names


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "csrc", "gen"))
