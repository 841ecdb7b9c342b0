/* CPU port of the reference's time-dominant loops (TEST / BASELINE INFRASTRUCTURE ONLY -- never linked into the product).
 *
 * Same loop nests, same OpenMP decomposition and the same compiler flags (-O3 -ffast-math -fopenmp,
 * CMakeLists.txt:10) as the reference, restated in C because no Fortran compiler exists in this image:
 *   afesp_ref_ring     the 6-deep o^3 v^3 loop of update_amplitudes_restricted   (src/ccsd.f90:1680-1695)
 *   afesp_ref_triples  the (i,j,k) loop of do_ccsd_t_spatial, 12 dot products per (a,b,c)   (src/ccsd.f90:2152-2233)
 * Arrays are column-major with the reference's index order.  bench.py times these (plus the ladder dgemm through the
 * image's OpenBLAS) on the GPU box's host cores as the reported cpu_baseline / --impl reference arm. */
#include <omp.h>
#include <stdlib.h>
#include <string.h>

#define T2(i, j, a, b) t2[(i) + o * ((j) + (size_t)o * ((a) + (size_t)v * (b)))]
#define ASYM(i, j, a, b) asym[(i) + o * ((j) + (size_t)o * ((a) + (size_t)v * (b)))]
#define IOVOV(i, e, m, b) I_ovov[(i) + o * ((e) + (size_t)v * ((m) + (size_t)o * (b)))]
#define IVOOV(e, j, m, b) I_voov[(e) + v * ((j) + (size_t)o * ((m) + (size_t)o * (b)))]
#define TMP(i, j, a, b) tmp_t2[(i) + o * ((j) + (size_t)o * ((a) + (size_t)v * (b)))]

/* bmax < v restricts the outermost loop to b < bmax: a bounded, extrapolatable sample of the same loop nest. */
void afesp_ref_ring(int o, int v, const double* t2, const double* I_ovov, const double* asym, const double* I_voov,
                    double* tmp_t2, int bmax) {
#pragma omp parallel for schedule(static, 10) collapse(3)
  for (int b = 0; b < bmax; ++b)
    for (int a = 0; a < v; ++a)
      for (int j = 0; j < o; ++j)
        for (int i = 0; i < o; ++i) {
          double tmp = 0.0;
          for (int e = 0; e < v; ++e)
            for (int m = 0; m < o; ++m)
              tmp += -T2(m, j, a, e) * IOVOV(i, e, m, b) - IOVOV(i, e, m, a) * T2(m, j, e, b) +
                     ASYM(m, i, e, a) * IVOOV(e, j, m, b);
          TMP(i, j, a, b) += tmp;
        }
}

/* t2r(b,a,j,i) = t2(i,j,a,b); vvovv(d,k,b,c) = v_vvov(c,b,k,d); vovoo(l,c,j,k) = v_oovo(k,j,c,l)   (:2056-2066) */
#define T2R(f, a, j, i) t2r[(f) + v * ((a) + (size_t)v * ((j) + (size_t)o * (i)))]
#define VVOVV(d, k, b, c) vvovv[(d) + v * ((k) + (size_t)o * ((b) + (size_t)v * (c)))]
#define VOVOO(l, c, j, k) vovoo[(l) + o * ((c) + (size_t)v * ((j) + (size_t)o * (k)))]
#define VOOVV(i, j, a, b) voovv[(i) + o * ((j) + (size_t)o * ((a) + (size_t)v * (b)))]
#define T1(i, a) t1[(i) + o * (a)]
#define X3(p, a, b, c) p[(a) + v * ((b) + (size_t)v * (c))]

static inline double dotv(const double* x, const double* y, int n) {
  double s = 0.0;
  for (int q = 0; q < n; ++q) s += x[q] * y[q];
  return s;
}

static void make_x_bar(int v, const double* x, double* xb) { /* src/ccsd.f90:2314-2318 */
  for (int c = 0; c < v; ++c)
    for (int b = 0; b < v; ++b)
      for (int a = 0; a < v; ++a)
        X3(xb, a, b, c) = 4.0 * X3(x, a, b, c) / 3.0 - 2.0 * X3(x, a, c, b) + 2.0 * X3(x, c, a, b) / 3.0;
}

/* out[0..3] = e_T, e_TT, D_T, D_TT over the listed ordered triples (no CR part). */
/* amax < v restricts the outer virtual loop to a < amax (bounded timing sample of the same loop nest; the sums are then
 * partial and only the time is meaningful). */
void afesp_ref_triples_bounded(int o, int v, const double* t1, const double* t2, const double* t2r, const double* vvovv,
                               const double* vovoo, const double* voovv, const double* eps, int ntri, const int* ijk,
                               int doing_T, int doing_R, int amax, double* out) {
  double e_T = 0.0, e_TT = 0.0, D_T = 0.0, D_TT = 0.0;
  const size_t v3 = (size_t)v * v * v;
#pragma omp parallel reduction(+ : e_T, e_TT, D_T, D_TT)
  {
    double* w = (double*)calloc(v3, sizeof(double));
    double* t3 = (double*)calloc(v3, sizeof(double));
    double* tb = (double*)malloc(v3 * sizeof(double));
    double* z3 = (double*)calloc(v3, sizeof(double));
    double* zb = (double*)calloc(v3, sizeof(double));
    double* y = (double*)calloc(v3, sizeof(double));
#pragma omp for schedule(static, 1)
    for (int t = 0; t < ntri; ++t) {
      const int i = ijk[3 * t], j = ijk[3 * t + 1], k = ijk[3 * t + 2];
      for (int a = 0; a < amax; ++a)
        for (int b = 0; b < v; ++b)
          for (int c = 0; c < v; ++c) {
            double x = dotv(&T2R(0, a, j, i), &VVOVV(0, k, b, c), v) - dotv(&T2(0, i, b, a), &VOVOO(0, c, j, k), o) +
                       dotv(&T2R(0, b, i, j), &VVOVV(0, k, a, c), v) - dotv(&T2(0, j, a, b), &VOVOO(0, c, i, k), o) +
                       dotv(&T2R(0, c, j, k), &VVOVV(0, i, b, a), v) - dotv(&T2(0, k, b, c), &VOVOO(0, a, j, i), o) +
                       dotv(&T2R(0, a, k, i), &VVOVV(0, j, c, b), v) - dotv(&T2(0, i, c, a), &VOVOO(0, b, k, j), o) +
                       dotv(&T2R(0, b, k, j), &VVOVV(0, i, c, a), v) - dotv(&T2(0, j, c, b), &VOVOO(0, a, k, i), o) +
                       dotv(&T2R(0, c, i, k), &VVOVV(0, j, a, b), v) - dotv(&T2(0, k, a, c), &VOVOO(0, b, i, j), o);
            const double d = eps[i] + eps[j] + eps[k] - eps[a + o] - eps[b + o] - eps[c + o];
            X3(w, a, b, c) = x;
            X3(t3, a, b, c) = x / d;
            if (doing_T)
              X3(z3, a, b, c) = (T1(i, a) * VOOVV(j, k, b, c) + T1(j, b) * VOOVV(i, k, a, c) + T1(k, c) * VOOVV(i, j, a, b)) / d;
            if (doing_R)
              X3(y, a, b, c) = T1(i, a) * T1(j, b) * T1(k, c) + T1(i, a) * T2(j, k, b, c) + T1(j, b) * T2(i, k, a, c) +
                               T1(k, c) * T2(i, j, a, b);
          }
      make_x_bar(v, t3, tb);
      if (doing_T && doing_R) make_x_bar(v, z3, zb); /* Q2: only for (T) and (R or CR), :2211-2215 */
      double tmp = dotv(tb, w, (int)v3);
      e_T += tmp;
      if (doing_T) e_TT += tmp + dotv(zb, w, (int)v3);
      if (doing_R) {
        tmp = dotv(tb, y, (int)v3);
        D_T += tmp;
        if (doing_T) D_TT += tmp + dotv(zb, y, (int)v3);
      }
    }
    free(w); free(t3); free(tb); free(z3); free(zb); free(y);
  }
  out[0] = e_T; out[1] = e_TT; out[2] = D_T; out[3] = D_TT;
}

void afesp_ref_triples(int o, int v, const double* t1, const double* t2, const double* t2r, const double* vvovv,
                       const double* vovoo, const double* voovv, const double* eps, int ntri, const int* ijk,
                       int doing_T, int doing_R, double* out) {
  afesp_ref_triples_bounded(o, v, t1, t2, t2r, vvovv, vovoo, voovv, eps, ntri, ijk, doing_T, doing_R, v, out);
}


/* Completely renormalised variant (doing_CR, src/ccsd.f90:2188-2193, 2222-2226): the M3 array from the CR intermediates
 * I_vovv_pp(v,o,v,v), I_ooov_pp(o,o,o,v), with the reference's strided sums.  out[0..5] = e_T, e_TT, D_T, D_TT, e_CR, e_CRT. */
#define IVOVV(d, k, b, c) I_vovv_pp[(d) + v * ((k) + (size_t)o * ((b) + (size_t)v * (c)))]
#define IOOOV(j, k, l, c) I_ooov_pp[(j) + o * ((k) + (size_t)o * ((l) + (size_t)o * (c)))]
static inline double dots(const double* x, size_t sx, const double* y, size_t sy, int n) {
  double s = 0.0;
  for (int q = 0; q < n; ++q) s += x[q * sx] * y[q * sy];
  return s;
}
void afesp_ref_triples_cr(int o, int v, const double* t1, const double* t2, const double* t2r, const double* vvovv,
                          const double* vovoo, const double* voovv, const double* I_vovv_pp, const double* I_ooov_pp,
                          const double* eps, int ntri, const int* ijk, int doing_T, double* out) {
  double e_T = 0.0, e_TT = 0.0, D_T = 0.0, D_TT = 0.0, e_CR = 0.0, e_CRT = 0.0;
  const size_t v3 = (size_t)v * v * v, se = (size_t)o * o * v /* stride of the last index of t2 */, so2 = (size_t)o * o;
#pragma omp parallel reduction(+ : e_T, e_TT, D_T, D_TT, e_CR, e_CRT)
  {
    double* w = (double*)calloc(v3, sizeof(double));
    double* t3 = (double*)calloc(v3, sizeof(double));
    double* tb = (double*)malloc(v3 * sizeof(double));
    double* z3 = (double*)calloc(v3, sizeof(double));
    double* zb = (double*)calloc(v3, sizeof(double));
    double* y = (double*)calloc(v3, sizeof(double));
    double* m3 = (double*)calloc(v3, sizeof(double));
#pragma omp for schedule(static, 10)
    for (int t = 0; t < ntri; ++t) {
      const int i = ijk[3 * t], j = ijk[3 * t + 1], k = ijk[3 * t + 2];
      for (int a = 0; a < v; ++a)
        for (int b = 0; b < v; ++b)
          for (int c = 0; c < v; ++c) {
            double x = dotv(&T2R(0, a, j, i), &VVOVV(0, k, b, c), v) - dotv(&T2(0, i, b, a), &VOVOO(0, c, j, k), o) +
                       dotv(&T2R(0, b, i, j), &VVOVV(0, k, a, c), v) - dotv(&T2(0, j, a, b), &VOVOO(0, c, i, k), o) +
                       dotv(&T2R(0, c, j, k), &VVOVV(0, i, b, a), v) - dotv(&T2(0, k, b, c), &VOVOO(0, a, j, i), o) +
                       dotv(&T2R(0, a, k, i), &VVOVV(0, j, c, b), v) - dotv(&T2(0, i, c, a), &VOVOO(0, b, k, j), o) +
                       dotv(&T2R(0, b, k, j), &VVOVV(0, i, c, a), v) - dotv(&T2(0, j, c, b), &VOVOO(0, a, k, i), o) +
                       dotv(&T2R(0, c, i, k), &VVOVV(0, j, a, b), v) - dotv(&T2(0, k, a, c), &VOVOO(0, b, i, j), o);
            const double d = eps[i] + eps[j] + eps[k] - eps[a + o] - eps[b + o] - eps[c + o];
            X3(w, a, b, c) = x;
            X3(t3, a, b, c) = x / d;
            if (doing_T)
              X3(z3, a, b, c) = (T1(i, a) * VOOVV(j, k, b, c) + T1(j, b) * VOOVV(i, k, a, c) + T1(k, c) * VOOVV(i, j, a, b)) / d;
            X3(y, a, b, c) = T1(i, a) * T1(j, b) * T1(k, c) + T1(i, a) * T2(j, k, b, c) + T1(j, b) * T2(i, k, a, c) +
                             T1(k, c) * T2(i, j, a, b);
            X3(m3, a, b, c) =
                dots(&T2(i, j, a, 0), se, &IVOVV(0, k, b, c), 1, v) - dots(&T2(0, i, b, a), 1, &IOOOV(j, k, 0, c), so2, o) +
                dots(&T2(j, i, b, 0), se, &IVOVV(0, k, a, c), 1, v) - dots(&T2(0, j, a, b), 1, &IOOOV(i, k, 0, c), so2, o) +
                dots(&T2(k, j, c, 0), se, &IVOVV(0, i, b, a), 1, v) - dots(&T2(0, k, b, c), 1, &IOOOV(j, i, 0, a), so2, o) +
                dots(&T2(i, k, a, 0), se, &IVOVV(0, j, c, b), 1, v) - dots(&T2(0, i, c, a), 1, &IOOOV(k, j, 0, b), so2, o) +
                dots(&T2(j, k, b, 0), se, &IVOVV(0, i, c, a), 1, v) - dots(&T2(0, j, c, b), 1, &IOOOV(k, i, 0, a), so2, o) +
                dots(&T2(k, i, c, 0), se, &IVOVV(0, j, a, b), 1, v) - dots(&T2(0, k, a, c), 1, &IOOOV(i, j, 0, b), so2, o);
          }
      make_x_bar(v, t3, tb);
      if (doing_T) make_x_bar(v, z3, zb);
      double tmp = dotv(tb, w, (int)v3);
      e_T += tmp;
      if (doing_T) e_TT += tmp + dotv(zb, w, (int)v3);
      tmp = dotv(tb, m3, (int)v3);
      e_CR += tmp;
      if (doing_T) e_CRT += tmp + dotv(zb, m3, (int)v3);
      tmp = dotv(tb, y, (int)v3);
      D_T += tmp;
      if (doing_T) D_TT += tmp + dotv(zb, y, (int)v3);
    }
    free(w); free(t3); free(tb); free(z3); free(zb); free(y); free(m3);
  }
  out[0] = e_T; out[1] = e_TT; out[2] = D_T; out[3] = D_TT; out[4] = e_CR; out[5] = e_CRT;
}

int afesp_ref_threads(void) { return omp_get_max_threads(); }


/* Epilogue of the BLAS orbit form of the [T] accumulator (oracle/afesp_oracle.py: triples_bracket_T_orbit_form) for large
 * shapes (tests/golden/make_bench_pins.py cpu_T_fast): the six permuted products of one (i <= j <= k) orbit have been
 * accumulated by dgemm (beta = 1) into three column-major buffers according to which virtual label indexes their rows,
 *     W(a,b,c) = Y0(a; b,c) + Y1(b; a,c) + Y2(c; a,b),
 * and the orbit's contribution is  sum_abc W(abc) / D3(abc) * [ 8 W(abc) - 4 (W(acb) + W(cba) + W(bac)) + 2 (W(cab) + W(bca)) ]
 * (times orderings/6, applied by the caller).  w is a v^3 scratch array.  eo3 = eps_i + eps_j + eps_k. */
double afesp_orbit_T_epilogue(int v, const double* y0, const double* y1, const double* y2, const double* ev, double eo3,
                              double* w) {
  const size_t n = (size_t)v, n2 = n * n;
#pragma omp parallel for schedule(static)
  for (int c = 0; c < v; ++c)
    for (int b = 0; b < v; ++b)
      for (int a = 0; a < v; ++a)
        w[a + n * b + n2 * c] = y0[a + n * b + n2 * c] + y1[b + n * a + n2 * c] + y2[c + n * a + n2 * b];
  double e = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : e)
  for (int c = 0; c < v; ++c)
    for (int b = 0; b < v; ++b)
      for (int a = 0; a < v; ++a) {
        const double wabc = w[a + n * b + n2 * c];
        const double s = 8.0 * wabc - 4.0 * (w[a + n * c + n2 * b] + w[c + n * b + n2 * a] + w[b + n * a + n2 * c]) +
                         2.0 * (w[c + n * a + n2 * b] + w[b + n * c + n2 * a]);
        e += wabc * s / (eo3 - ev[a] - ev[b] - ev[c]);
      }
  return e;
}
