/* c_ints_shim.c -- host-side `_c_ints` module with the reference's 11 legacy entry points.
 *
 * pychem imports `_c_ints` in two places: Methods/integrals.py:8 (the per-class micro-steps of its
 * integral code) and Util/structures.py:18,931 (ShellPair construction calls
 * `_c_ints.shellpair_quantities` eagerly for every shell pair).  Upstream's own extension
 * (Methods/_c_ints.c:68-81, built by Methods/setup.py:4-15) cannot link: its Boys table
 * (c_ints/interpolation_table.c) is a missing blob and setup.py needs distutils / numpy.distutils.
 * This module is what makes the drop-in true outside this repository: same module name, same 11
 * function names, same PyArg_ParseTuple formats (cited per function), arrays borrowed from the
 * caller and written IN PLACE, None returned -- so the reference's Python (structures.py,
 * integrals.py) runs unchanged on top of it, while the hot callers (evaluate_2e_ints,
 * make_coulomb_exchange_matrices) are rebound to the CUDA path by pychem_b200.install().
 * Built by pychem_b200/setup_c_ints.py (setuptools Extension) into pychem_b200/compat/.
 *
 * Written from the call sites and array layouts, not from upstream's C: arrays are read through
 * numpy's contiguous views with explicit index arithmetic (no row-pointer tables, so nothing
 * leaks per call as upstream's PyArray_AsCArray does), shapes are checked, errors are Python
 * exceptions.  The arithmetic per element follows the same formulas in the same order, so results
 * agree with upstream's kernels to the last bits (tests/test_c_ints_shim.py compares both).
 *
 * Layouts (upstream conventions): Cartesian components of shell l in the order lx = l..0,
 * ly = l-lx..0; component index (ly+lz)(ly+lz+1)/2 + lz (c_ints/angmom_index.c:3-15); integral
 * arrays are [component * nprim + prim] on both sides, primitives (ia, ib) flattened ia-major.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#define NPY_NO_DEPRECATED_API NPY_1_7_API_VERSION
#include <numpy/arrayobject.h>

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "pc_boys_table.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ---------------------------------------------------------------------------------------------
 * Boys table (rows m = 0..PCS_NM-1), generated on first use
 * ------------------------------------------------------------------------------------------- */
#define PCS_NM 29                      /* interpolation_table.h:11: orders up to 28 */
static double* g_tab = NULL;

static const double* boys_row(int m, int j) {
  if (!g_tab) {
    g_tab = (double*)malloc(sizeof(double) * (size_t)PCS_NM * PCB_NPOINTS * 4);
    pcb_make_table(PCS_NM, g_tab);
  }
  return g_tab + ((size_t)m * PCB_NPOINTS + j) * 4;
}

/* F_m(T) with upstream's three branches (two_electron_fundamentals.c:51-83) */
static void boys_values(int l_max, double T, double R2, double* F) {
  if (R2 < 1.e-14) {
    for (int m = 0; m <= l_max; ++m) F[m] = 1 / (2 * (double)m + 1);
    return;
  }
  const double sT = T / (2 * PCB_D);
  const int j = (int)sT;
  if (j < PCB_NPOINTS) {
    for (int m = 0; m <= l_max; ++m) {
      const double* c = boys_row(m, j);
      F[m] = c[0] + sT * (c[1] + sT * (c[2] + sT * c[3]));
    }
  } else {
    for (int m = 0; m <= l_max; ++m) F[m] = tgamma(m + 0.5) / (2 * pow(T, m + 0.5));
  }
}

/* ---------------------------------------------------------------------------------------------
 * array access
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  PyArrayObject* arr;     /* owned reference (a contiguous view or copy for inputs) */
  double* p;
  npy_intp n;             /* total elements */
} View;

static void view_release(View* v, int count) {
  for (int k = 0; k < count; ++k) Py_XDECREF(v[k].arr);
}

/* input: any array-like, read through a C-contiguous double view (copied when it has to be:
 * the reference passes transposed and negated views, Methods/integrals.py:585-586) */
static int view_in(PyObject* obj, View* v, npy_intp min_elems, const char* what) {
  v->arr = (PyArrayObject*)PyArray_FROM_OTF(obj, NPY_DOUBLE, NPY_ARRAY_IN_ARRAY);
  if (!v->arr) return -1;
  v->p = (double*)PyArray_DATA(v->arr);
  v->n = PyArray_SIZE(v->arr);
  if (v->n < min_elems) {
    PyErr_Format(PyExc_ValueError, "_c_ints: %s holds %zd values, %zd needed", what, (Py_ssize_t)v->n, (Py_ssize_t)min_elems);
    Py_CLEAR(v->arr);
    return -1;
  }
  return 0;
}

/* output: must be the caller's own C-contiguous, writable double array (results are written in
 * place; upstream allocates them with numpy.zeros, Methods/integrals.py:469-470,583,607,630) */
static int view_out(PyObject* obj, View* v, npy_intp min_elems, const char* what) {
  if (!PyArray_Check(obj) || PyArray_TYPE((PyArrayObject*)obj) != NPY_DOUBLE ||
      !PyArray_ISCARRAY((PyArrayObject*)obj)) {
    PyErr_Format(PyExc_TypeError, "_c_ints: %s must be a C-contiguous writable float64 array (it is filled in place)", what);
    return -1;
  }
  Py_INCREF(obj);
  v->arr = (PyArrayObject*)obj;
  v->p = (double*)PyArray_DATA(v->arr);
  v->n = PyArray_SIZE(v->arr);
  if (v->n < min_elems) {
    PyErr_Format(PyExc_ValueError, "_c_ints: %s holds %zd values, %zd needed", what, (Py_ssize_t)v->n, (Py_ssize_t)min_elems);
    Py_CLEAR(v->arr);
    return -1;
  }
  return 0;
}

static PyObject* parse_error(void) {
  PyErr_SetString(PyExc_TypeError, "Error parsing objects passed to C");     /* _c_ints.c:135 */
  return NULL;
}

/* ---------------------------------------------------------------------------------------------
 * Cartesian bookkeeping
 * ------------------------------------------------------------------------------------------- */
#define NCART(l) ((((l) + 1) * ((l) + 2)) / 2)
#define LMAX_SHIM 14

static int comp_index(int lx, int ly, int lz) { (void)lx; return (ly + lz) * (ly + lz + 1) / 2 + lz; }

/* components of shell l in loop order; returns their number */
static int comp_list(int l, int (*c)[3]) {
  int n = 0;
  for (int x = l; x >= 0; --x)
    for (int y = l - x; y >= 0; --y, ++n) { c[n][0] = x; c[n][1] = y; c[n][2] = l - x - y; }
  return n;
}

/* first non-zero direction of a component (two_electron_vrr.c:33-48), -1 for s */
static int first_dir(const int* c) { return c[0] ? 0 : (c[1] ? 1 : (c[2] ? 2 : -1)); }

static int check_l(int l, const char* what) {
  if (l < 0 || l > LMAX_SHIM) { PyErr_Format(PyExc_ValueError, "_c_ints: %s = %d out of range", what, l); return -1; }
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * shellpair_quantities(sigmas, overlaps, centres, alpha[], A[3], n_alpha, beta[], B[3], n_beta)
 * "OOOOOiOOi" (_c_ints.c:132); shellpair_quantities.c:23-36; caller Util/structures.py:931
 * ------------------------------------------------------------------------------------------- */
static PyObject* shim_shellpair_quantities(PyObject* self, PyObject* args) {
  PyObject *o_sig, *o_ov, *o_cen, *o_al, *o_A, *o_be, *o_B;
  int na, nb;
  if (!PyArg_ParseTuple(args, "OOOOOiOOi", &o_sig, &o_ov, &o_cen, &o_al, &o_A, &na, &o_be, &o_B, &nb)) return parse_error();
  if (na < 0 || nb < 0) { PyErr_SetString(PyExc_ValueError, "_c_ints: negative primitive count"); return NULL; }
  View v[7];
  memset(v, 0, sizeof(v));
  const npy_intp nn = (npy_intp)na * nb;
  if (view_out(o_sig, &v[0], nn, "sigmas") || view_out(o_ov, &v[1], nn, "overlaps") || view_out(o_cen, &v[2], 3 * nn, "centres") ||
      view_in(o_al, &v[3], na, "alpha_exponents") || view_in(o_A, &v[4], 3, "A") || view_in(o_be, &v[5], nb, "beta_exponents") ||
      view_in(o_B, &v[6], 3, "B")) { view_release(v, 7); return NULL; }
  const double *al = v[3].p, *A = v[4].p, *be = v[5].p, *B = v[6].p;
  double r2 = 0;
  for (int i = 0; i < 3; ++i) { const double r = A[i] - B[i]; r2 += r * r; }
  for (int ia = 0; ia < na; ++ia)
    for (int ib = 0; ib < nb; ++ib) {
      const double a = al[ia], b = be[ib], s = 1.0 / (a + b);
      const size_t k = (size_t)ia * nb + ib;
      v[0].p[k] = s;
      v[1].p[k] = pow(M_PI * s, 1.5) * exp(-a * b * s * r2);
      for (int i = 0; i < 3; ++i) v[2].p[3 * k + i] = (a * A[i] + b * B[i]) * s;
    }
  view_release(v, 7);
  Py_RETURN_NONE;
}

/* ---------------------------------------------------------------------------------------------
 * two_electron_bound(bounds, C_P, C_Q, nla, nlb, nlc, nld)   "OOOiiii" (_c_ints.c:174)
 * bounds[((a nlb + b) nlc + c) nld + d] = sqrt(C_P[a][b] C_Q[c][d])   (two_electron_bound.c)
 * ------------------------------------------------------------------------------------------- */
static PyObject* shim_two_electron_bound(PyObject* self, PyObject* args) {
  PyObject *o_b, *o_P, *o_Q;
  int nla, nlb, nlc, nld;
  if (!PyArg_ParseTuple(args, "OOOiiii", &o_b, &o_P, &o_Q, &nla, &nlb, &nlc, &nld)) return parse_error();
  if (nla < 0 || nlb < 0 || nlc < 0 || nld < 0) { PyErr_SetString(PyExc_ValueError, "_c_ints: negative size"); return NULL; }
  View v[3];
  memset(v, 0, sizeof(v));
  if (view_out(o_b, &v[0], (npy_intp)nla * nlb * nlc * nld, "bounds") || view_in(o_P, &v[1], (npy_intp)nla * nlb, "C_P") ||
      view_in(o_Q, &v[2], (npy_intp)nlc * nld, "C_Q")) { view_release(v, 3); return NULL; }
  size_t k = 0;
  for (int a = 0; a < nla * nlb; ++a)
    for (int c = 0; c < nlc * nld; ++c) v[0].p[k++] = sqrt(v[1].p[a] * v[2].p[c]);
  view_release(v, 3);
  Py_RETURN_NONE;
}

/* ---------------------------------------------------------------------------------------------
 * spherical Bessel functions z^-m j_m(z), m = 0..l_max, three regimes (spherical_bessel_j.c:5-81)
 * ------------------------------------------------------------------------------------------- */
static void bessel_scaled(double* j, double z, int l_max) {
  if (z < 1.e1) {                         /* series about z = 0 */
    const int kmax = z < 1.e-3 ? 1 : (z < 1.e-1 ? 3 : (z < 1.e0 ? 6 : 20));
    const double z2 = z * z;
    double df = 1.0;
    for (int m = 0; m <= l_max; ++m) {
      const double mm1 = (double)(2 * m + 1);
      df *= mm1;
      double sum = 1.0, mmdf = 1.0, mm = mm1, zz = 1.0, sign = 1.0, denom = 1.0;
      for (int k = 1; k <= kmax; ++k) {
        mm += 2; mmdf *= mm; zz *= z2; sign *= -1; denom *= 2 * k;
        sum += sign * zz / (denom * mmdf);
      }
      j[m] = sum / df;
    }
  } else if (z > 1.e2) {                  /* asymptotic expansion */
    const double zi1 = 1 / z, zi2 = zi1 * zi1;
    double zi = 1.0, zoff = z;
    j[0] = sin(z) * zi1;
    for (int m = 1; m <= l_max; ++m) {
      const double mm1 = (double)(m * (m + 1) / 2);
      zi *= zi1;
      zoff -= 0.5 * M_PI;
      j[m] = zi * (zi1 * sin(zoff) + mm1 * zi2 * cos(zoff));
    }
  } else {                                /* upward recursion up to order 16, zero above */
    double t[17];
    const double zi1 = 1 / z;
    t[0] = sin(z) * zi1;
    t[1] = (sin(z) - z * cos(z)) * (zi1 * zi1);
    for (int m = 2; m <= 16; ++m) t[m] = ((2 * m - 1) * t[m - 1] * zi1 - t[m - 2]);
    double scale = 1.0;
    for (int m = 0; m <= 16; ++m) { t[m] *= scale; scale *= zi1; }
    for (int m = 0; m <= l_max; ++m) j[m] = m <= 16 ? t[m] : 0.0;
  }
  for (int m = 0; m <= l_max; ++m)
    if (fabs(j[m]) < 1.e-16) j[m] = 0.0;
}

/* ---------------------------------------------------------------------------------------------
 * two_electron_fundamentals(fund, sP, UP, P, sQ, UQ, Q, R, na, nb, nc, nd, l_max, ints_type, grid)
 * "OOOOOOOOiiiiiid" (_c_ints.c:216).  fund[m][bra prim][ket prim], R[bra prim][ket prim][3].
 * ints_type 0: Gill-scaled Boys fundamentals (two_electron_fundamentals.c:41-89);
 * ints_type 1: scattering kernel (two_electron_scattering.c:21-80)
 * ------------------------------------------------------------------------------------------- */
static PyObject* shim_two_electron_fundamentals(PyObject* self, PyObject* args) {
  PyObject *o_f, *o_sP, *o_UP, *o_P, *o_sQ, *o_UQ, *o_Q, *o_R;
  int na, nb, nc, nd, l_max, ints_type;
  double grid;
  if (!PyArg_ParseTuple(args, "OOOOOOOOiiiiiid", &o_f, &o_sP, &o_UP, &o_P, &o_sQ, &o_UQ, &o_Q, &o_R, &na, &nb, &nc, &nd, &l_max,
                        &ints_type, &grid)) return parse_error();
  if (na < 0 || nb < 0 || nc < 0 || nd < 0 || l_max < 0 || l_max >= PCS_NM) { PyErr_SetString(PyExc_ValueError, "_c_ints: bad sizes"); return NULL; }
  const npy_intp nbra = (npy_intp)na * nb, nket = (npy_intp)nc * nd;
  View v[8];
  memset(v, 0, sizeof(v));
  if (view_out(o_f, &v[0], (l_max + 1) * nbra * nket, "fundamentals") || view_in(o_sP, &v[1], nbra, "sigma_P") ||
      view_in(o_UP, &v[2], nbra, "U_P") || view_in(o_P, &v[3], 3 * nbra, "P") || view_in(o_sQ, &v[4], nket, "sigma_Q") ||
      view_in(o_UQ, &v[5], nket, "U_Q") || view_in(o_Q, &v[6], 3 * nket, "Q") || view_out(o_R, &v[7], 3 * nbra * nket, "R")) {
    view_release(v, 8);
    return NULL;
  }
  const double pf = pow(2 / M_PI, 0.5);
  double F[PCS_NM], jz[PCS_NM], S2[PCS_NM], fdf[PCS_NM];
  if (ints_type == 1) {
    double Spow = 1.0, df = 1.0;
    S2[0] = 1.0; fdf[0] = 1.0;
    for (int m = 1; m <= l_max; ++m) { Spow *= grid * grid; df *= (double)(2 * m + 1); S2[m] = Spow; fdf[m] = 1.0 / df; }
  }
  for (npy_intp ib = 0; ib < nbra; ++ib)
    for (npy_intp ik = 0; ik < nket; ++ik) {
      const double U = v[2].p[ib] * v[5].p[ik];
      double R2 = 0;
      for (int i = 0; i < 3; ++i) {
        const double r = v[3].p[3 * ib + i] - v[6].p[3 * ik + i];
        v[7].p[(ib * nket + ik) * 3 + i] = r;
        R2 += r * r;
      }
      double* out = v[0].p + ib * nket + ik;              /* stride nbra*nket per order */
      const npy_intp st = nbra * nket;
      if (ints_type == 1) {
        const double eS = exp(-0.25 * grid * grid * (v[1].p[ib] + v[4].p[ik]));
        if (grid < 1.e-14) {
          out[0] = U;
          for (int m = 1; m <= l_max; ++m) out[m * st] = 0;
        } else if (R2 < 1.e-14) {
          for (int m = 0; m <= l_max; ++m) out[m * st] = U * eS * S2[m] * fdf[m];
        } else {
          bessel_scaled(jz, grid * sqrt(R2), l_max);
          for (int m = 0; m <= l_max; ++m) out[m * st] = U * eS * S2[m] * jz[m];
        }
      } else {
        const double theta_sq = 1 / (v[1].p[ib] + v[4].p[ik]);
        boys_values(l_max, theta_sq * R2, R2, F);
        for (int m = 0; m <= l_max; ++m) out[m * st] = pf * U * pow(2 * theta_sq, m + 0.5) * F[m];
      }
    }
  view_release(v, 8);
  Py_RETURN_NONE;
}

/* ---------------------------------------------------------------------------------------------
 * two_electron_vrr(target, b0, b1, b2, b3, b4, zeta, eta, kappa, Rx, R, na, nb, nc, nd, lbra, lket,
 * kappa_index)   "OOOOOOOOOOOiiiiiii" (_c_ints.c:270); one vertical step in Gill-scaled form
 * (two_electron_vrr.c:92-108):
 *   [a|c] = Rx_i kappa zeta [a-1_i|c]^(m) + R_i zeta [a-1_i|c]^(m+1)
 *         + (a_i - 1) zeta ([a-2_i|c]^(m) - zeta [a-2_i|c]^(m+1)) + c_i zeta eta [a-1_i|c-1_i]^(m+1)
 * arrays [bra component * na*nb + bra prim][ket component * nc*nd + ket prim]; the signs of Rx and
 * R are the caller's (integrals.py:561-569,586)
 * ------------------------------------------------------------------------------------------- */
static PyObject* shim_two_electron_vrr(PyObject* self, PyObject* args) {
  PyObject *o_t, *o_b[5], *o_ze, *o_et, *o_ka, *o_Rx, *o_R;
  int na, nb, nc, nd, lbra, lket, kidx;
  if (!PyArg_ParseTuple(args, "OOOOOOOOOOOiiiiiii", &o_t, &o_b[0], &o_b[1], &o_b[2], &o_b[3], &o_b[4], &o_ze, &o_et, &o_ka, &o_Rx,
                        &o_R, &na, &nb, &nc, &nd, &lbra, &lket, &kidx)) return parse_error();
  if (check_l(lbra, "lbra") || check_l(lket, "lket")) return NULL;
  if (lbra < 1 || na < 1 || nb < 1 || nc < 1 || nd < 1) { PyErr_SetString(PyExc_ValueError, "_c_ints: two_electron_vrr needs lbra >= 1 and primitives"); return NULL; }
  const npy_intp nbp = (npy_intp)na * nb, nkp = (npy_intp)nc * nd;
  const npy_intp ncol = NCART(lket) * nkp;                  /* columns of target, base0..3 */
  const npy_intp ncol1 = lket > 0 ? NCART(lket - 1) * nkp : 0;   /* columns of base4 */
  View v[11];
  memset(v, 0, sizeof(v));
  int bad = view_out(o_t, &v[0], NCART(lbra) * nbp * ncol, "target_ints") ||
            view_in(o_b[0], &v[1], NCART(lbra - 1) * nbp * ncol, "base0") || view_in(o_b[1], &v[2], NCART(lbra - 1) * nbp * ncol, "base1") ||
            view_in(o_b[2], &v[3], lbra > 1 ? NCART(lbra - 2) * nbp * ncol : 0, "base2") ||
            view_in(o_b[3], &v[4], lbra > 1 ? NCART(lbra - 2) * nbp * ncol : 0, "base3") ||
            view_in(o_b[4], &v[5], lket > 0 ? NCART(lbra - 1) * nbp * ncol1 : 0, "base4") || view_in(o_ze, &v[6], nbp, "zeta") ||
            view_in(o_et, &v[7], nkp, "eta") || view_in(o_ka, &v[8], kidx == 0 ? na : nb, "kappa") || view_in(o_Rx, &v[9], 3, "Rx") ||
            view_in(o_R, &v[10], 3 * nbp * nkp, "R");
  if (bad) { view_release(v, 11); return NULL; }
  int bc[NCART(LMAX_SHIM)][3], kc[NCART(LMAX_SHIM)][3];
  const int nbc = comp_list(lbra, bc), nkc = comp_list(lket, kc);
  for (int ia_c = 0; ia_c < nbc; ++ia_c) {
    const int* a = bc[ia_c];
    const int dir = first_dir(a);
    int a0[3] = {a[0], a[1], a[2]};
    a0[dir] -= 1;
    const int na_prev = a0[dir];                          /* a_i - 1 */
    int a1[3] = {a0[0], a0[1], a0[2]};
    a1[dir] -= 1;
    const npy_intp row = (npy_intp)comp_index(a[0], a[1], a[2]) * nbp;
    const npy_intp row0 = (npy_intp)comp_index(a0[0], a0[1], a0[2]) * nbp;
    const npy_intp row1 = na_prev > 0 ? (npy_intp)comp_index(a1[0], a1[1], a1[2]) * nbp : 0;
    for (int ia = 0; ia < na; ++ia)
      for (int ib = 0; ib < nb; ++ib) {
        const npy_intp bp = (npy_intp)ia * nb + ib;
        const double z = v[6].p[bp];
        const double kap = v[8].p[kidx == 0 ? ia : ib];
        for (int ic_c = 0; ic_c < nkc; ++ic_c) {
          const int* c = kc[ic_c];
          const npy_intp col = (npy_intp)comp_index(c[0], c[1], c[2]) * nkp;
          const int cn = c[dir];                          /* c_i */
          npy_intp col1 = 0;
          if (cn > 0) {
            int c1[3] = {c[0], c[1], c[2]};
            c1[dir] -= 1;
            col1 = (npy_intp)comp_index(c1[0], c1[1], c1[2]) * nkp;
          }
          for (npy_intp kp = 0; kp < nkp; ++kp) {
            const double Ri = v[10].p[(bp * nkp + kp) * 3 + dir];
            double val = v[9].p[dir] * kap * z * v[1].p[(row0 + bp) * ncol + col + kp] + Ri * z * v[2].p[(row0 + bp) * ncol + col + kp];
            if (na_prev > 0) val += na_prev * z * (v[3].p[(row1 + bp) * ncol + col + kp] - z * v[4].p[(row1 + bp) * ncol + col + kp]);
            if (cn > 0) val += cn * z * v[7].p[kp] * v[5].p[(row0 + bp) * ncol1 + col1 + kp];
            v[0].p[(row + bp) * ncol + col + kp] = val;
          }
        }
      }
  }
  view_release(v, 11);
  Py_RETURN_NONE;
}

/* ---------------------------------------------------------------------------------------------
 * two_electron_contract(out, prim, cc_bra, cc_ket, na, nb, nc, nd, lbra, lket)   "OOOOiiiiii"
 * (_c_ints.c:323): out[bra comp][ket comp] += sum_prims cc_bra cc_ket prim  (two_electron_contract.c:45)
 * ------------------------------------------------------------------------------------------- */
static PyObject* shim_two_electron_contract(PyObject* self, PyObject* args) {
  PyObject *o_c, *o_p, *o_cb, *o_ck;
  int na, nb, nc, nd, lbra, lket;
  if (!PyArg_ParseTuple(args, "OOOOiiiiii", &o_c, &o_p, &o_cb, &o_ck, &na, &nb, &nc, &nd, &lbra, &lket)) return parse_error();
  if (check_l(lbra, "lbra") || check_l(lket, "lket") || na < 0 || nb < 0 || nc < 0 || nd < 0) return NULL;
  const npy_intp nbp = (npy_intp)na * nb, nkp = (npy_intp)nc * nd, nbc = NCART(lbra), nkc = NCART(lket);
  View v[4];
  memset(v, 0, sizeof(v));
  if (view_out(o_c, &v[0], nbc * nkc, "contracted_ints") || view_in(o_p, &v[1], nbc * nbp * nkc * nkp, "primitive_ints") ||
      view_in(o_cb, &v[2], nbp, "cc_bra") || view_in(o_ck, &v[3], nkp, "cc_ket")) { view_release(v, 4); return NULL; }
  for (npy_intp b = 0; b < nbc; ++b)
    for (npy_intp bp = 0; bp < nbp; ++bp)
      for (npy_intp k = 0; k < nkc; ++k)
        for (npy_intp kp = 0; kp < nkp; ++kp)
          v[0].p[b * nkc + k] += v[2].p[bp] * v[3].p[kp] * v[1].p[(b * nbp + bp) * (nkc * nkp) + k * nkp + kp];
  view_release(v, 4);
  Py_RETURN_NONE;
}

/* ---------------------------------------------------------------------------------------------
 * two_electron_hrr(target, b0, b1, Rx, la, lb, lc, ld, goofy)   "OOOOiiiii" (_c_ints.c:364)
 *   (a, b+1_i | cd) = (a+1_i, b | cd) + Rx_i (a b | cd)            (two_electron_hrr.c:82)
 * goofy = 0 builds b from a; goofy = 1 builds a from b (the caller passes Rx with its sign,
 * integrals.py:617-623).  Arrays [bra pair component][ket pair component], pair component =
 * ia * ncart(second) + ib.  `la`, `lb` are the angular momenta BEFORE the step.
 * Upstream indexes base0 of the goofy case with the row length angmom_index(0,0,lb+1) -- one
 * short of ncart(lb+1), harmless while la = 0 at the step (lower shell s or p: every s, p, d basis) and wrong for (d f) pairs
 * (oracle/make_golden_f.py).  The shim uses the true row length ncart(lb) of the array it is given
 * and therefore agrees with upstream wherever upstream is right.
 * ------------------------------------------------------------------------------------------- */
static PyObject* shim_two_electron_hrr(PyObject* self, PyObject* args) {
  PyObject *o_t, *o_b0, *o_b1, *o_Rx;
  int la, lb, lc, ld, goofy;
  if (!PyArg_ParseTuple(args, "OOOOiiiii", &o_t, &o_b0, &o_b1, &o_Rx, &la, &lb, &lc, &ld, &goofy)) return parse_error();
  if (check_l(la + 1, "la") || check_l(lb + 1, "lb") || check_l(lc, "lc") || check_l(ld, "ld")) return NULL;
  const npy_intp nket = (npy_intp)NCART(lc) * NCART(ld);
  /* target (ta, tb); base0 = (one side raised, other unchanged); base1 = (la, lb) */
  const int ta = goofy ? la + 1 : la, tb = goofy ? lb : lb + 1;
  const int b0a = goofy ? la : la + 1, b0b = goofy ? lb + 1 : lb;
  View v[4];
  memset(v, 0, sizeof(v));
  if (view_out(o_t, &v[0], (npy_intp)NCART(ta) * NCART(tb) * nket, "target_ints") ||
      view_in(o_b0, &v[1], (npy_intp)NCART(b0a) * NCART(b0b) * nket, "base0") ||
      view_in(o_b1, &v[2], (npy_intp)NCART(la) * NCART(lb) * nket, "base1") || view_in(o_Rx, &v[3], 3, "Rx")) {
    view_release(v, 4);
    return NULL;
  }
  int ca[NCART(LMAX_SHIM)][3], cb[NCART(LMAX_SHIM)][3];
  const int nca = comp_list(ta, ca), ncb = comp_list(tb, cb);
  npy_intp row = 0;
  for (int i = 0; i < nca; ++i)
    for (int j = 0; j < ncb; ++j, ++row) {
      /* the component that was raised in this step gives the direction */
      const int* grown = goofy ? ca[i] : cb[j];
      const int dir = first_dir(grown);
      int a[3] = {ca[i][0], ca[i][1], ca[i][2]}, b[3] = {cb[j][0], cb[j][1], cb[j][2]};
      npy_intp r0, r1;
      if (!goofy) {           /* (a, b) <- (a+1_i, b-1_i) + Rx_i (a, b-1_i) */
        b[dir] -= 1;
        r1 = (npy_intp)comp_index(a[0], a[1], a[2]) * NCART(lb) + comp_index(b[0], b[1], b[2]);
        a[dir] += 1;
        r0 = (npy_intp)comp_index(a[0], a[1], a[2]) * NCART(b0b) + comp_index(b[0], b[1], b[2]);
      } else {                /* (a, b) <- (a-1_i, b+1_i) + Rx_i (a-1_i, b) */
        a[dir] -= 1;
        r1 = (npy_intp)comp_index(a[0], a[1], a[2]) * NCART(lb) + comp_index(b[0], b[1], b[2]);
        b[dir] += 1;
        r0 = (npy_intp)comp_index(a[0], a[1], a[2]) * NCART(b0b) + comp_index(b[0], b[1], b[2]);
      }
      const double rx = v[3].p[dir];
      for (npy_intp k = 0; k < nket; ++k) v[0].p[row * nket + k] = v[1].p[r0 * nket + k] + rx * v[2].p[r1 * nket + k];
    }
  view_release(v, 4);
  Py_RETURN_NONE;
}

/* ---------------------------------------------------------------------------------------------
 * one_electron_fundamentals(f, sigma_P, U_P, P, Rc, Z, na, nb, l_max)   "OOOOOdiii" (_c_ints.c:409)
 *   f[m][ia][ib] = -Z sqrt(2/pi) U sqrt(2 zeta) F_m(zeta |P - Rc|^2)    (one_electron_fundamentals.c:47-91)
 * ------------------------------------------------------------------------------------------- */
static PyObject* shim_one_electron_fundamentals(PyObject* self, PyObject* args) {
  PyObject *o_f, *o_s, *o_U, *o_P, *o_Rc;
  double Z;
  int na, nb, l_max;
  if (!PyArg_ParseTuple(args, "OOOOOdiii", &o_f, &o_s, &o_U, &o_P, &o_Rc, &Z, &na, &nb, &l_max)) return parse_error();
  if (na < 0 || nb < 0 || l_max < 0 || l_max >= PCS_NM) { PyErr_SetString(PyExc_ValueError, "_c_ints: bad sizes"); return NULL; }
  const npy_intp np_ = (npy_intp)na * nb;
  View v[5];
  memset(v, 0, sizeof(v));
  if (view_out(o_f, &v[0], (l_max + 1) * np_, "nuclear_fundamentals") || view_in(o_s, &v[1], np_, "sigma_P") ||
      view_in(o_U, &v[2], np_, "U_P") || view_in(o_P, &v[3], 3 * np_, "P") || view_in(o_Rc, &v[4], 3, "Rc")) {
    view_release(v, 5);
    return NULL;
  }
  const double pf = -Z * pow(2 / M_PI, 0.5);
  double F[PCS_NM];
  for (npy_intp k = 0; k < np_; ++k) {
    double R2 = 0;
    for (int i = 0; i < 3; ++i) { const double r = v[3].p[3 * k + i] - v[4].p[i]; R2 += r * r; }
    const double zeta = 1.0 / v[1].p[k];
    const double spf = pf * v[2].p[k] * pow(2.0 * zeta, 0.5);
    boys_values(l_max, zeta * R2, R2, F);
    for (int m = 0; m <= l_max; ++m) v[0].p[m * np_ + k] = spf * F[m];
  }
  view_release(v, 5);
  Py_RETURN_NONE;
}

/* base list of the one-electron steps: a Python list of 2-D arrays; missing entries alias the first */
static int list_views(PyObject* list, int nwant, int nhave, View* out, npy_intp min_elems, const char* what) {
  if (!PyList_Check(list) || PyList_Size(list) < (nhave < 1 ? 1 : nhave)) {
    PyErr_Format(PyExc_TypeError, "_c_ints: %s must be a list of at least %d arrays", what, nhave < 1 ? 1 : nhave);
    return -1;
  }
  for (int k = 0; k < nwant; ++k) {
    PyObject* item = PyList_GetItem(list, k < nhave ? k : 0);
    if (view_in(item, &out[k], k < nhave ? min_elems : 0, what)) { view_release(out, k); return -1; }
  }
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * one_electron_vrr(target, [bases], sigma_P, P, Ra, Rc, atom_index, nbase, na, nb, la)
 * "OOOOOOiiiii" (_c_ints.c:457).  atom_index == -1: overlap-type step
 *   [a|0] = (P - A)_i [a-1_i] + (a_i - 1) sigma/2 [a-2_i]                    (bases: [a-1], [a-2])
 * otherwise nuclear attraction
 *   [a|0]^(m) = (P-A)_i [a-1_i]^(m) - (P-C)_i [a-1_i]^(m+1) + (a_i-1) sigma/2 ([a-2_i]^(m) - [a-2_i]^(m+1))
 * (one_electron_vrr.c).  Arrays [component * na + ia][ib].
 * ------------------------------------------------------------------------------------------- */
static PyObject* shim_one_electron_vrr(PyObject* self, PyObject* args) {
  PyObject *o_t, *o_bases, *o_s, *o_P, *o_Ra, *o_Rc;
  int atom_index, nbase, na, nb, la;
  if (!PyArg_ParseTuple(args, "OOOOOOiiiii", &o_t, &o_bases, &o_s, &o_P, &o_Ra, &o_Rc, &atom_index, &nbase, &na, &nb, &la)) return parse_error();
  if (check_l(la, "la") || la < 1 || na < 1 || nb < 1) { if (!PyErr_Occurred()) PyErr_SetString(PyExc_ValueError, "_c_ints: one_electron_vrr needs la >= 1"); return NULL; }
  const int aux = atom_index == -1 ? 0 : 1;
  View v[5], b[4];
  memset(v, 0, sizeof(v));
  memset(b, 0, sizeof(b));
  const npy_intp np_ = (npy_intp)na * nb;
  if (view_out(o_t, &v[0], NCART(la) * np_, "target_ints") || view_in(o_s, &v[1], np_, "sigma_P") || view_in(o_P, &v[2], 3 * np_, "P") ||
      view_in(o_Ra, &v[3], 3, "Ra") || view_in(o_Rc, &v[4], aux ? 3 : 0, "Rc")) { view_release(v, 5); return NULL; }   /* overlap: Rc = [] (integrals.py:389) */
  /* upstream's wrapper: base1 = list[1] if nbase > 1, base2/3 = list[2], list[3] if nbase == 4 */
  const int have = nbase == 4 ? 4 : (nbase > 1 ? 2 : 1);
  if (list_views(o_bases, 4, have, b, 0, "base_ints")) { view_release(v, 5); return NULL; }
  int ac[NCART(LMAX_SHIM)][3];
  const int nac = comp_list(la, ac);
  int ok = 1;
  for (int i = 0; i < nac && ok; ++i) {
    const int* a = ac[i];
    const int dir = first_dir(a);
    int a0[3] = {a[0], a[1], a[2]};
    a0[dir] -= 1;
    const int nprev = a0[dir];
    int a1[3] = {a0[0], a0[1], a0[2]};
    a1[dir] -= 1;
    const npy_intp r0 = (npy_intp)comp_index(a0[0], a0[1], a0[2]) * na;
    const npy_intp r1 = nprev > 0 ? (npy_intp)comp_index(a1[0], a1[1], a1[2]) * na : 0;
    /* which list entries hold [a-2]: overlap: bases[1]; nuclear: bases[2], bases[3] */
    const View* lo0 = aux ? &b[2] : &b[1];
    const View* lo1 = &b[3];
    if ((r0 + na) * nb > b[0].n || (aux && (r0 + na) * nb > b[1].n) ||
        (nprev > 0 && ((r1 + na) * nb > lo0->n || (aux && (r1 + na) * nb > lo1->n)))) { ok = 0; break; }
    for (int ia = 0; ia < na; ++ia)
      for (int ib = 0; ib < nb; ++ib) {
        const npy_intp k = (npy_intp)ia * nb + ib;
        const double PA = v[2].p[3 * k + dir] - v[3].p[dir];
        double val;
        if (!aux) {
          val = PA * b[0].p[(r0 + ia) * nb + ib];
          if (nprev > 0) val += nprev * 0.5 * v[1].p[k] * lo0->p[(r1 + ia) * nb + ib];
        } else {
          const double PC = v[2].p[3 * k + dir] - v[4].p[dir];
          val = PA * b[0].p[(r0 + ia) * nb + ib] - PC * b[1].p[(r0 + ia) * nb + ib];
          if (nprev > 0) val += nprev * 0.5 * v[1].p[k] * (lo0->p[(r1 + ia) * nb + ib] - lo1->p[(r1 + ia) * nb + ib]);
        }
        v[0].p[((npy_intp)i * na + ia) * nb + ib] = val;
      }
  }
  view_release(v, 5);
  view_release(b, 4);
  if (!ok) { PyErr_SetString(PyExc_ValueError, "_c_ints: one_electron_vrr base arrays are too small"); return NULL; }
  Py_RETURN_NONE;
}

/* ---------------------------------------------------------------------------------------------
 * one_electron_hrr(target, [bases], Rab, nbase, na, nb, la, lb)   "OOOiiiii" (_c_ints.c:524)
 *   [a | b] = [a+1_i | b-1_i] + Rab_i [a | b-1_i]                           (one_electron_hrr.c)
 * arrays [a component * na + ia][b component * nb + ib]; target and base1 have ncart(la) bra
 * components, base0 ncart(la+1); `lb` is the ket momentum AFTER the step
 * ------------------------------------------------------------------------------------------- */
static PyObject* shim_one_electron_hrr(PyObject* self, PyObject* args) {
  PyObject *o_t, *o_bases, *o_R;
  int nbase, na, nb, la, lb;
  if (!PyArg_ParseTuple(args, "OOOiiiii", &o_t, &o_bases, &o_R, &nbase, &na, &nb, &la, &lb)) return parse_error();
  if (check_l(la + 1, "la") || check_l(lb, "lb") || lb < 1 || na < 1 || nb < 1) { if (!PyErr_Occurred()) PyErr_SetString(PyExc_ValueError, "_c_ints: one_electron_hrr needs lb >= 1"); return NULL; }
  View v[2], b[2];
  memset(v, 0, sizeof(v));
  memset(b, 0, sizeof(b));
  const npy_intp wt = (npy_intp)NCART(lb) * nb, w0 = (npy_intp)NCART(lb - 1) * nb;
  if (view_out(o_t, &v[0], (npy_intp)NCART(la) * na * wt, "target_ints") || view_in(o_R, &v[1], 3, "Rab")) { view_release(v, 2); return NULL; }
  if (list_views(o_bases, 2, nbase == 2 ? 2 : 1, b, 0, "base_ints")) { view_release(v, 2); return NULL; }
  if (b[0].n < (npy_intp)NCART(la + 1) * na * w0 || b[1].n < (npy_intp)NCART(la) * na * w0) {
    view_release(v, 2); view_release(b, 2);
    PyErr_SetString(PyExc_ValueError, "_c_ints: one_electron_hrr base arrays are too small");
    return NULL;
  }
  int bc[NCART(LMAX_SHIM)][3], ac[NCART(LMAX_SHIM)][3];
  const int nbc = comp_list(lb, bc), nac = comp_list(la, ac);
  for (int j = 0; j < nbc; ++j) {
    const int dir = first_dir(bc[j]);
    int b0[3] = {bc[j][0], bc[j][1], bc[j][2]};
    b0[dir] -= 1;
    const npy_intp cb0 = (npy_intp)comp_index(b0[0], b0[1], b0[2]) * nb, cb1 = (npy_intp)comp_index(bc[j][0], bc[j][1], bc[j][2]) * nb;
    for (int i = 0; i < nac; ++i) {
      int a1[3] = {ac[i][0], ac[i][1], ac[i][2]};
      a1[dir] += 1;
      const npy_intp ra0 = (npy_intp)comp_index(ac[i][0], ac[i][1], ac[i][2]) * na, ra1 = (npy_intp)comp_index(a1[0], a1[1], a1[2]) * na;
      for (int ia = 0; ia < na; ++ia)
        for (int ib = 0; ib < nb; ++ib)
          v[0].p[(ra0 + ia) * wt + cb1 + ib] = b[0].p[(ra1 + ia) * w0 + cb0 + ib] + v[1].p[dir] * b[1].p[(ra0 + ia) * w0 + cb0 + ib];
    }
  }
  view_release(v, 2);
  view_release(b, 2);
  Py_RETURN_NONE;
}

/* ---------------------------------------------------------------------------------------------
 * one_electron_kinetic(target, [bases], exB, nbase, na, nb, la, lb)   "OOOiiiii" (_c_ints.c:575)
 *   T[a|b] = beta (2 l_b + 3) S[a|b] - 2 beta^2 (S[a|b+2x] + S[a|b+2y] + S[a|b+2z])
 *            - 1/2 sum_i b_i (b_i - 1) S[a|b-2_i]                          (one_electron_kinetic.c:62-66)
 * bases = [S(la, lb+2), S(la, lb), S(la, lb-2)]; arrays [a comp * na + ia][b comp * nb + ib]
 * ------------------------------------------------------------------------------------------- */
static PyObject* shim_one_electron_kinetic(PyObject* self, PyObject* args) {
  PyObject *o_t, *o_bases, *o_ex;
  int nbase, na, nb, la, lb;
  if (!PyArg_ParseTuple(args, "OOOiiiii", &o_t, &o_bases, &o_ex, &nbase, &na, &nb, &la, &lb)) return parse_error();
  if (check_l(la, "la") || check_l(lb + 2, "lb") || na < 1 || nb < 1) { if (!PyErr_Occurred()) PyErr_SetString(PyExc_ValueError, "_c_ints: bad sizes"); return NULL; }
  View v[2], b[3];
  memset(v, 0, sizeof(v));
  memset(b, 0, sizeof(b));
  const npy_intp rows = (npy_intp)NCART(la) * na;
  const npy_intp wp = (npy_intp)NCART(lb + 2) * nb, w = (npy_intp)NCART(lb) * nb, wm = lb >= 2 ? (npy_intp)NCART(lb - 2) * nb : 0;
  if (view_out(o_t, &v[0], rows * w, "target_ints") || view_in(o_ex, &v[1], nb, "exB")) { view_release(v, 2); return NULL; }
  if (list_views(o_bases, 3, nbase == 3 ? 3 : 2, b, 0, "base_ints")) { view_release(v, 2); return NULL; }
  if (b[0].n < rows * wp || b[1].n < rows * w || (lb >= 2 && nbase == 3 && b[2].n < rows * wm)) {
    view_release(v, 2); view_release(b, 3);
    PyErr_SetString(PyExc_ValueError, "_c_ints: one_electron_kinetic base arrays are too small");
    return NULL;
  }
  int bc[NCART(LMAX_SHIM)][3];
  const int nbc = comp_list(lb, bc);
  for (int j = 0; j < nbc; ++j) {
    const int* c = bc[j];
    const npy_intp cb = (npy_intp)comp_index(c[0], c[1], c[2]) * nb;
    npy_intp up[3], dn[3];
    for (int i = 0; i < 3; ++i) {
      int t[3] = {c[0], c[1], c[2]};
      t[i] += 2;
      up[i] = (npy_intp)comp_index(t[0], t[1], t[2]) * nb;
      t[i] -= 4;
      dn[i] = c[i] > 1 ? (npy_intp)comp_index(t[0], t[1], t[2]) * nb : -1;
    }
    for (npy_intp r = 0; r < rows; ++r)
      for (int ib = 0; ib < nb; ++ib) {
        const double beta = v[1].p[ib];
        double val = beta * (2 * c[0] + 2 * c[1] + 2 * c[2] + 3) * b[1].p[r * w + cb + ib];
        val -= 2 * beta * beta * (b[0].p[r * wp + up[0] + ib] + b[0].p[r * wp + up[1] + ib] + b[0].p[r * wp + up[2] + ib]);
        for (int i = 0; i < 3; ++i)
          if (dn[i] >= 0) val -= 0.5 * c[i] * (c[i] - 1) * b[2].p[r * wm + dn[i] + ib];
        v[0].p[r * w + cb + ib] = val;
      }
  }
  view_release(v, 2);
  view_release(b, 3);
  Py_RETURN_NONE;
}

/* ---------------------------------------------------------------------------------------------
 * one_electron_contract(target, base, cc, na, nb, la, lb)   "OOOiiii" (_c_ints.c:625)
 *   target[a comp][b comp] += sum_{ia,ib} cc[ia][ib] base[a comp * na + ia][b comp * nb + ib]
 * ------------------------------------------------------------------------------------------- */
static PyObject* shim_one_electron_contract(PyObject* self, PyObject* args) {
  PyObject *o_t, *o_b, *o_cc;
  int na, nb, la, lb;
  if (!PyArg_ParseTuple(args, "OOOiiii", &o_t, &o_b, &o_cc, &na, &nb, &la, &lb)) return parse_error();
  if (check_l(la, "la") || check_l(lb, "lb") || na < 0 || nb < 0) return NULL;
  const npy_intp nac = NCART(la), nbc = NCART(lb);
  View v[3];
  memset(v, 0, sizeof(v));
  if (view_out(o_t, &v[0], nac * nbc, "target_ints") || view_in(o_b, &v[1], nac * na * nbc * nb, "base_ints") ||
      view_in(o_cc, &v[2], (npy_intp)na * nb, "cc")) { view_release(v, 3); return NULL; }
  for (npy_intp a = 0; a < nac; ++a)
    for (npy_intp ia = 0; ia < na; ++ia)
      for (npy_intp b = 0; b < nbc; ++b)
        for (npy_intp ib = 0; ib < nb; ++ib)
          v[0].p[a * nbc + b] += v[2].p[ia * nb + ib] * v[1].p[(a * na + ia) * (nbc * nb) + b * nb + ib];
  view_release(v, 3);
  Py_RETURN_NONE;
}

/* --------------------------------------------------------------------------------------------- */
static PyMethodDef shim_methods[] = {
    {"shellpair_quantities", shim_shellpair_quantities, METH_VARARGS, "sigma, U, P of every primitive pair of a shell pair (in place)"},
    {"two_electron_bound", shim_two_electron_bound, METH_VARARGS, "sqrt(C_P[a][b] C_Q[c][d]) (in place)"},
    {"two_electron_fundamentals", shim_two_electron_fundamentals, METH_VARARGS, "[00|00]^(m) of every primitive quartet (in place)"},
    {"two_electron_vrr", shim_two_electron_vrr, METH_VARARGS, "one vertical recursion step (in place)"},
    {"two_electron_contract", shim_two_electron_contract, METH_VARARGS, "contraction of primitive integrals (accumulates in place)"},
    {"two_electron_hrr", shim_two_electron_hrr, METH_VARARGS, "one horizontal recursion step (in place)"},
    {"one_electron_fundamentals", shim_one_electron_fundamentals, METH_VARARGS, "nuclear-attraction fundamentals (in place)"},
    {"one_electron_vrr", shim_one_electron_vrr, METH_VARARGS, "one-electron vertical step (in place)"},
    {"one_electron_hrr", shim_one_electron_hrr, METH_VARARGS, "one-electron horizontal step (in place)"},
    {"one_electron_kinetic", shim_one_electron_kinetic, METH_VARARGS, "kinetic-energy integrals from overlaps (in place)"},
    {"one_electron_contract", shim_one_electron_contract, METH_VARARGS, "one-electron contraction (accumulates in place)"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef shim_module = {PyModuleDef_HEAD_INIT, "_c_ints",
                                         "pychem's legacy `_c_ints` entry points (host side), provided by pychem_b200", -1,
                                         shim_methods, NULL, NULL, NULL, NULL};

PyMODINIT_FUNC PyInit__c_ints(void) {
  PyObject* m = PyModule_Create(&shim_module);
  if (!m) return NULL;
  import_array();
  PyModule_AddStringConstant(m, "__provider__", "pychem_b200");
  return m;
}
