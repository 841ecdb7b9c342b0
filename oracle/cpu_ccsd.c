/* CPU port of the reference's spin-free CCSD iteration and AO->MO transform (TEST / BASELINE INFRASTRUCTURE ONLY --
 * never linked into the product; see oracle/cpu_port.py).
 *
 * One call of afesp_ref_ccsd_iter() issues the same dgemm calls (through the image's OpenBLAS, handed in as a function
 * pointer), the same omp_reshape / antisymmetrise passes and the same naive OpenMP loop nests, in the same order, as
 *     update_restricted_intermediates   src/ccsd.f90:1040-1312
 *     update_amplitudes_restricted      src/ccsd.f90:1538-1732
 * with dgemm_wrapper / omp_reshape / antisymmetrise / deantisymmetrise as in src/linalg.fpp:58-340.  It is restated in C
 * because no Fortran compiler exists in this image; compiler flags are the reference's (-O3 -ffast-math -fopenmp,
 * CMakeLists.txt:10).  Arrays are column-major with the reference's index order; indices below are 0-based.
 *
 * afesp_ref_ao2mo() restates the four O(n^5) quarter transforms and the repack of do_mp2_spatial (src/mp2.f90:321-410).
 * The `*max` arguments restrict the OUTERMOST loop of a nest to a slab (bounded timing samples at shapes where the
 * full nest takes minutes); with the full extents the results are exact and are tested against the NumPy oracle. */
#include <omp.h>
#include <stdlib.h>
#include <string.h>

typedef long bint;
/* Fortran-interface dgemm of the image's OpenBLAS: numpy bundles the ILP64 build (scipy_dgemm_64_, 64-bit integers),
 * scipy the LP64 one (scipy_dgemm_, 32-bit integers); cpu_port.py hands in whichever it finds. */
typedef void (*dgemm64_fn)(const char*, const char*, const bint*, const bint*, const bint*, const double*, const double*,
                           const bint*, const double*, const bint*, const double*, double*, const bint*);
typedef void (*dgemm32_fn)(const char*, const char*, const int*, const int*, const int*, const double*, const double*,
                           const int*, const double*, const int*, const double*, double*, const int*);
static void* g_dgemm = 0;
static int g_ilp64 = 1;
void afesp_ref_set_dgemm(void* fn, int ilp64) { g_dgemm = fn; g_ilp64 = ilp64; }
void afesp_ref_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }

static double now(void) { return omp_get_wtime(); }

/* dgemm_wrapper (src/linalg.fpp:58-89) */
static void gemm(char ta, char tb, bint M, bint N, bint K, const double* A, const double* B, double* C, double alpha,
                 double beta) {
  bint lda = (ta == 'T') ? K : M, ldb = (tb == 'T') ? N : K;
  if (g_ilp64) {
    ((dgemm64_fn)g_dgemm)(&ta, &tb, &M, &N, &K, &alpha, A, &lda, B, &ldb, &beta, C, &M);
  } else {
    const int m = (int)M, n = (int)N, k = (int)K, la = (int)lda, lb = (int)ldb;
    ((dgemm32_fn)g_dgemm)(&ta, &tb, &m, &n, &k, &alpha, A, &la, B, &lb, &beta, C, &m);
  }
}
/* plain dgemm for the harness (ladder column blocks, tests) */
void afesp_ref_dgemm(char ta, char tb, long M, long N, long K, const double* A, const double* B, double* C, double alpha,
                     double beta) { gemm(ta, tb, M, N, K, A, B, C, alpha, beta); }

#define IX(a, b, c, d, n1, n2, n3) ((size_t)(a) + (size_t)(n1) * ((size_t)(b) + (size_t)(n2) * ((size_t)(c) + (size_t)(n3) * (size_t)(d))))

/* omp_reshape (src/linalg.fpp:99-156): out(idx[ord0], idx[ord1], idx[ord2], idx[ord3]) = beta*out + in(i,j,k,l) */
void afesp_ref_reshape(double* out, const double* in, const int* d, const char* order, int has_beta, double beta) {
  const int p0 = order[0] - '1', p1 = order[1] - '1', p2 = order[2] - '1', p3 = order[3] - '1';
  const int e0 = d[p0], e1 = d[p1], e2 = d[p2];
  const size_t total = (size_t)d[0] * d[1] * d[2] * d[3];
  if (!has_beta) { memset(out, 0, total * sizeof(double)); beta = 0.0; }
#pragma omp parallel for
  for (int l = 0; l < d[3]; ++l)
    for (int k = 0; k < d[2]; ++k)
      for (int j = 0; j < d[1]; ++j)
        for (int i = 0; i < d[0]; ++i) {
          const int idx[4] = {i, j, k, l};
          const size_t po = IX(idx[p0], idx[p1], idx[p2], idx[p3], e0, e1, e2);
          out[po] = beta * out[po] + in[IX(i, j, k, l, d[0], d[1], d[2])];
        }
}

/* antisymmetrise(..., inplace=.true.) (src/linalg.fpp:176-221): A = 2A - A^(pair swapped) */
static void antisym(double* a, const int* d, const char* order) {
  const int iu = d[0], ju = d[1], ku = d[2], lu = d[3];
  if (!strncmp(order, "1243", 4)) {
#pragma omp parallel for
    for (int l = 0; l < lu; ++l)
      for (int k = 0; k <= l; ++k)
        for (int j = 0; j < ju; ++j)
          for (int i = 0; i < iu; ++i) {
            double* x = &a[IX(i, j, k, l, iu, ju, ku)]; double* y = &a[IX(i, j, l, k, iu, ju, ku)];
            const double u = *x, t = *y; *x = 2 * u - t; *y = 2 * t - u;
          }
  } else if (!strncmp(order, "2134", 4)) {
#pragma omp parallel for
    for (int l = 0; l < lu; ++l)
      for (int k = 0; k < ku; ++k)
        for (int j = 0; j < ju; ++j)
          for (int i = 0; i <= j; ++i) {
            double* x = &a[IX(i, j, k, l, iu, ju, ku)]; double* y = &a[IX(j, i, k, l, iu, ju, ku)];
            const double u = *x, t = *y; *x = 2 * u - t; *y = 2 * t - u;
          }
  } else { /* 4231 */
#pragma omp parallel for
    for (int l = 0; l < lu; ++l)
      for (int k = 0; k < ku; ++k)
        for (int j = 0; j < ju; ++j)
          for (int i = 0; i <= l; ++i) {
            double* x = &a[IX(i, j, k, l, iu, ju, ku)]; double* y = &a[IX(l, j, k, i, iu, ju, ku)];
            const double u = *x, t = *y; *x = 2 * u - t; *y = 2 * t - u;
          }
  }
}

/* deantisymmetrise (src/linalg.fpp:276-340) */
static void deantisym(double* a, const int* d, const char* order) {
  const int iu = d[0], ju = d[1], ku = d[2], lu = d[3];
#define DEANTI(X, Y) { double* x = (X); double* y = (Y); const double s = *x + *y, df = (*x - *y) / 3; *x = (s + df) / 2; *y = (s - df) / 2; }
  if (!strncmp(order, "1243", 4)) {
#pragma omp parallel for
    for (int l = 0; l < lu; ++l)
      for (int k = 0; k <= l; ++k)
        for (int j = 0; j < ju; ++j)
          for (int i = 0; i < iu; ++i) DEANTI(&a[IX(i, j, k, l, iu, ju, ku)], &a[IX(i, j, l, k, iu, ju, ku)])
  } else if (!strncmp(order, "2134", 4)) {
#pragma omp parallel for
    for (int l = 0; l < lu; ++l)
      for (int k = 0; k < ku; ++k)
        for (int j = 0; j < ju; ++j)
          for (int i = 0; i <= j; ++i) DEANTI(&a[IX(i, j, k, l, iu, ju, ku)], &a[IX(j, i, k, l, iu, ju, ku)])
  } else {
#pragma omp parallel for
    for (int l = 0; l < lu; ++l)
      for (int k = 0; k < ku; ++k)
        for (int j = 0; j < ju; ++j)
          for (int i = 0; i <= l; ++i) DEANTI(&a[IX(i, j, k, l, iu, ju, ku)], &a[IX(l, j, k, i, iu, ju, ku)])
  }
#undef DEANTI
}

static double* dalloc(size_t n) { return (double*)malloc((n ? n : 1) * sizeof(double)); }

/* Sampling controls (all <= 0 or >= extent: the full loop).  ring_bmax: outer b of :1680-1695; iovov_amax: outer a of
 * :1170-1182; ladder_ncol: columns (a,b) of v_vvvv handed in (the caller passes a column block of the dense slice and
 * its width; v*v = everything).  times[0..7]: seconds of intermediates-dgemm part, I_ovov loop, other intermediate
 * loops, T1 part, ladder dgemm, ring loop, remaining T2 terms, final combine/divide.
 *
 * v_oovv(o,o,v,v) v_ovov(o,v,o,v) v_vvov(v,v,o,v) v_oovo(o,o,v,o) v_oooo(o,o,o,o) v_vvvv(v,v,v*v or ncol block)
 * t1(o,v), t2(o,o,v,v) are updated in place (t1 = tmp_t1/D_ia, t2 = tmp_t2/D_ijab, :1727-1728). */
void afesp_ref_ccsd_iter(int o, int v, double* v_oovv, const double* v_ovov, double* v_vvov, double* v_oovo,
                         const double* v_oooo, const double* v_vvvv, int ladder_ncol, const double* eps, double* t1,
                         double* t2, int ring_bmax, int iovov_amax, double* times) {
  const size_t oovv = (size_t)o * o * v * v, ov = (size_t)o * v;
  const int d_oovv[4] = {o, o, v, v}, d_vvov[4] = {v, v, o, v}, d_oovo[4] = {o, o, v, o}, d_ovov[4] = {o, v, o, v};
  if (ring_bmax <= 0 || ring_bmax > v) ring_bmax = v;
  if (iovov_amax <= 0 || iovov_amax > v) iovov_amax = v;
  if (ladder_ncol <= 0 || ladder_ncol > v * v) ladder_ncol = v * v;
  double *asym = dalloc(oovv), *c = dalloc(oovv), *I_vo = dalloc(ov), *I_vv = dalloc((size_t)v * v),
         *I_oo_p = dalloc((size_t)o * o), *I_oo = dalloc((size_t)o * o), *I_oooo = dalloc((size_t)o * o * o * o),
         *I_ovov = dalloc(oovv), *I_voov = dalloc(oovv), *I_vovv_p = dalloc((size_t)o * v * v * v),
         *x_voov = dalloc(oovv), *I_ooov_p = dalloc((size_t)o * o * o * v), *tmp_t1 = dalloc(ov), *tmp_t2 = dalloc(oovv);
  double* rt;  /* reshape_tmp */
  double t0 = now(), tt;
  for (int q = 0; q < 8; ++q) times[q] = 0.0;
#define LAP(slot) { tt = now(); times[slot] += tt - t0; t0 = tt; }

  /* ---- update_restricted_intermediates ---- */
  /* asym_t2 = 2 t2 - t2(j,i,a,b)   :1063-1064 */
  afesp_ref_reshape(asym, t2, d_oovv, "2134", 0, 0.0);
#pragma omp parallel for
  for (size_t q = 0; q < oovv; ++q) asym[q] = -asym[q] + 2 * t2[q];
  /* c = t2 + t1 t1   :1070-1079 */
#pragma omp parallel for collapse(2) schedule(static, 10)
  for (int b = 0; b < v; ++b)
    for (int a = 0; a < v; ++a)
      for (int j = 0; j < o; ++j)
        for (int i = 0; i < o; ++i) c[IX(i, j, a, b, o, o, v)] = t2[IX(i, j, a, b, o, o, v)] + t1[i + o * a] * t1[j + o * b];
  /* I_vo   :1089-1094 */
  antisym(v_oovv, d_oovv, "1243");
  rt = dalloc(oovv);
  afesp_ref_reshape(rt, v_oovv, d_oovv, "3124", 0, 0.0);
  gemm('N', 'N', ov, 1, ov, rt, t1, I_vo, 1.0, 0.0);
  free(rt);
  /* I_vv   :1101-1113 */
  antisym(v_vvov, d_vvov, "2134");
  rt = dalloc((size_t)v * v * o * v);
  afesp_ref_reshape(rt, v_vvov, d_vvov, "2431", 0, 0.0);
  gemm('N', 'N', (bint)v * v, 1, ov, rt, t1, I_vv, 1.0, 0.0);
  deantisym(v_vvov, d_vvov, "2134");
  free(rt);
  rt = dalloc(oovv);
  afesp_ref_reshape(rt, v_oovv, d_oovv, "4123", 0, 0.0);
  gemm('N', 'N', v, v, (bint)o * o * v, rt, c, I_vv, -1.0, 1.0);
  free(rt);
  deantisym(v_oovv, d_oovv, "1243");
  /* I_oo_p   :1121-1132 */
  antisym(v_oovo, d_oovo, "2134");
  rt = dalloc((size_t)o * o * o * v);
  afesp_ref_reshape(rt, v_oovo, d_oovo, "4213", 0, 0.0);
  gemm('N', 'N', (bint)o * o, 1, ov, rt, t1, I_oo_p, 1.0, 0.0);
  free(rt);
  deantisym(v_oovo, d_oovo, "2134");
  rt = dalloc(oovv);
  afesp_ref_reshape(rt, v_oovv, d_oovv, "1432", 0, 0.0);
  gemm('N', 'N', o, o, (bint)o * v * v, asym, rt, I_oo_p, 1.0, 1.0);
  free(rt);
  /* I_oo   :1136-1137 */
  gemm('N', 'N', o, o, v, t1, I_vo, I_oo, 1.0, 0.0);
  for (int q = 0; q < o * o; ++q) I_oo[q] += I_oo_p[q];
  /* I_oooo   :1143-1156 */
  memcpy(I_oooo, v_oooo, (size_t)o * o * o * o * sizeof(double));
  rt = dalloc(oovv);
  afesp_ref_reshape(rt, v_oovv, d_oovv, "3412", 0, 0.0);
  gemm('N', 'N', (bint)o * o, (bint)o * o, (bint)v * v, c, rt, I_oooo, 1.0, 1.0);
  free(rt);
  {
    rt = dalloc((size_t)v * o * o * o);
    afesp_ref_reshape(rt, v_oovo, d_oovo, "3214", 0, 0.0);
    double* scratch = dalloc((size_t)o * o * o * o);
    gemm('N', 'N', o, (bint)o * o * o, v, t1, rt, scratch, 1.0, 0.0);
    free(rt);
    rt = dalloc((size_t)o * o * o * o);
    const int d4[4] = {o, o, o, o};
    afesp_ref_reshape(rt, scratch, d4, "2143", 0, 0.0);
    for (size_t q = 0; q < (size_t)o * o * o * o; ++q) I_oooo[q] += scratch[q] + rt[q];
    free(rt); free(scratch);
  }
  LAP(0)
  /* I_ovov   :1165-1191 */
  memcpy(I_ovov, v_ovov, oovv * sizeof(double));
#pragma omp parallel for collapse(2) schedule(static, 10)
  for (int a = 0; a < iovov_amax; ++a)
    for (int i = 0; i < o; ++i)
      for (int b = 0; b < v; ++b)
        for (int j = 0; j < o; ++j)
          for (int e = 0; e < v; ++e)
            for (int m = 0; m < o; ++m)
              I_ovov[IX(j, b, i, a, o, v, o)] -= 0.5 * v_oovv[IX(m, i, b, e, o, o, v)] * c[IX(m, j, a, e, o, o, v)];
  LAP(1)
  {
    rt = dalloc((size_t)o * v * o * o);
    /* reshape(v_oovo, shape(o,v,o,o), order=(/4,3,2,1/)): out(j,b,i,m) = v_oovo(m,i,b,j)   :1185 */
#pragma omp parallel for collapse(2)
    for (int m = 0; m < o; ++m)
      for (int i = 0; i < o; ++i)
        for (int b = 0; b < v; ++b)
          for (int j = 0; j < o; ++j) rt[IX(j, b, i, m, o, v, o)] = v_oovo[IX(m, i, b, j, o, o, v)];
    gemm('N', 'N', (bint)o * o * v, v, o, rt, t1, I_ovov, -1.0, 1.0);
    free(rt);
    gemm('N', 'N', o, (bint)o * v * v, v, t1, v_vvov, I_ovov, 1.0, 1.0);
  }
  /* I_voov   :1205-1254 */
  {
    const int d_voov[4] = {v, o, o, v};
    rt = dalloc(oovv);
    afesp_ref_reshape(rt, v_oovv, d_oovv, "3124", 0, 0.0);
    antisym(rt, d_voov, "4231");
    double* scratch2 = dalloc(oovv);
    double* scratch = dalloc(oovv);
    afesp_ref_reshape(scratch, t2, d_oovv, "1342", 0, 0.0);
    gemm('N', 'N', ov, ov, ov, rt, scratch, scratch2, 0.5, 0.0);
    free(scratch);
    deantisym(rt, d_voov, "4231");
    scratch = dalloc(oovv);
    afesp_ref_reshape(scratch, c, d_oovv, "1432", 0, 0.0);
    gemm('N', 'N', ov, ov, ov, rt, scratch, scratch2, -0.5, 1.0);
    const int d_vovo[4] = {v, o, v, o};
    afesp_ref_reshape(I_voov, scratch2, d_vovo, "1423", 0, 0.0);
    free(scratch2); free(scratch); free(rt);
    rt = dalloc((size_t)v * o * o * o);
    afesp_ref_reshape(rt, v_oovo, d_oovo, "3412", 0, 0.0);
    gemm('N', 'N', (bint)o * o * v, v, o, rt, t1, I_voov, -1.0, 1.0);
    free(rt);
  }
  LAP(0)
#pragma omp parallel
  {
#pragma omp for collapse(2) schedule(static, 10)
    for (int a = 0; a < v; ++a)
      for (int i = 0; i < o; ++i)
        for (int b = 0; b < v; ++b)
          for (int j = 0; j < o; ++j) {
            double* x = &I_voov[IX(b, j, i, a, v, o, o)];
            *x += v_oovv[IX(j, i, a, b, o, o, v)];
            for (int e = 0; e < v; ++e) *x += v_vvov[IX(b, e, i, a, v, v, o)] * t1[j + o * e];
          }
    /* I_vovv_p   :1261-1275 */
#pragma omp for collapse(2) schedule(static, 10)
    for (int b = 0; b < v; ++b)
      for (int a = 0; a < v; ++a)
        for (int i = 0; i < o; ++i)
          for (int cc = 0; cc < v; ++cc) {
            double* x = &I_vovv_p[IX(cc, i, a, b, v, o, v)];
            *x = v_vvov[IX(b, a, i, cc, v, v, o)];
            for (int m = 0; m < o; ++m) *x -= v_oovv[IX(m, i, cc, b, o, o, v)] * t1[m + o * a];
          }
    /* x_voov   :1281-1292 */
#pragma omp for collapse(2) schedule(static, 10)
    for (int a = 0; a < v; ++a)
      for (int i = 0; i < o; ++i)
        for (int j = 0; j < o; ++j)
          for (int b = 0; b < v; ++b) {
            double s = 0.0;
            for (int e = 0; e < v; ++e) s += v_vvov[IX(b, e, i, a, v, v, o)] * t1[j + o * e];
            x_voov[IX(b, j, i, a, v, o, o)] = s;
          }
  }
  LAP(2)
  rt = dalloc(oovv);
  afesp_ref_reshape(rt, v_ovov, d_ovov, "4321", 0, 0.0);
  gemm('N', 'N', (bint)o * v * v, v, o, rt, t1, I_vovv_p, -1.0, 1.0);
  free(rt);
  /* I_ooov_p = reshape(v_oovo, order=(/2,1,4,3/)): out(j,k,i,a) = v_oovo(k,j,a,i)   :1306-1308 */
#pragma omp parallel for collapse(2)
  for (int a = 0; a < v; ++a)
    for (int i = 0; i < o; ++i)
      for (int k = 0; k < o; ++k)
        for (int j = 0; j < o; ++j) I_ooov_p[IX(j, k, i, a, o, o, o)] = v_oovo[IX(k, j, a, i, o, o, v)];
  gemm('N', 'N', (bint)o * o, ov, (bint)v * v, t2, v_vvov, I_ooov_p, 1.0, 1.0);
  gemm('N', 'N', o, (bint)o * o * v, v, t1, x_voov, I_ooov_p, 1.0, 1.0);
  LAP(0)

  /* ---- update_amplitudes_restricted ---- */
  gemm('N', 'N', o, v, v, t1, I_vv, tmp_t1, 1.0, 0.0);                       /* :1571 */
  gemm('N', 'N', o, v, o, I_oo_p, t1, tmp_t1, -1.0, 1.0);                    /* :1572 */
#pragma omp parallel for schedule(static, 10) collapse(2)
  for (int a = 0; a < v; ++a)                                                 /* :1577-1589 */
    for (int i = 0; i < o; ++i)
      for (int e = 0; e < v; ++e)
        for (int m = 0; m < o; ++m)
          tmp_t1[i + o * a] += I_vo[e + v * m] * asym[IX(m, i, e, a, o, o, v)] +
                               t1[m + o * e] * (2 * v_oovv[IX(m, i, e, a, o, o, v)] - v_ovov[IX(m, a, i, e, o, v, o)]);
  rt = dalloc((size_t)o * o * o * v);
  afesp_ref_reshape(rt, v_oovo, d_oovo, "2143", 0, 0.0);                      /* :1605-1608 */
  gemm('N', 'N', o, v, (bint)o * o * v, rt, asym, tmp_t1, -1.0, 1.0);
  free(rt);
#pragma omp parallel for schedule(static, 10) collapse(2)
  for (int a = 0; a < v; ++a)                                                 /* :1619-1633 */
    for (int i = 0; i < o; ++i) {
      double s = 0.0;
      for (int e = 0; e < v; ++e)
        for (int f = 0; f < v; ++f)
          for (int m = 0; m < o; ++m) s += v_vvov[IX(e, f, m, a, v, v, o)] * asym[IX(m, i, e, f, o, o, v)];
      tmp_t1[i + o * a] += s;
    }
  LAP(3)
  gemm('N', 'N', (bint)o * o * v, v, v, t2, I_vv, tmp_t2, 1.0, 0.0);         /* :1647 */
#pragma omp parallel for schedule(static, 10) collapse(2)
  for (int b = 0; b < v; ++b)                                                 /* :1654-1664 */
    for (int a = 0; a < v; ++a)
      for (int j = 0; j < o; ++j)
        for (int i = 0; i < o; ++i)
          for (int m = 0; m < o; ++m) tmp_t2[IX(i, j, a, b, o, o, v)] -= t2[IX(m, i, b, a, o, o, v)] * I_oo[j + o * m];
  LAP(6)
  gemm('N', 'N', (bint)o * o, ladder_ncol, (bint)v * v, c, v_vvvv, tmp_t2, 0.5, 1.0);   /* :1669 */
  LAP(4)
  gemm('N', 'N', (bint)o * o, (bint)v * v, (bint)o * o, I_oooo, c, tmp_t2, 0.5, 1.0);  /* :1673 */
  LAP(6)
#pragma omp parallel for schedule(static, 10) collapse(3)
  for (int b = 0; b < ring_bmax; ++b)                                         /* :1680-1695 */
    for (int a = 0; a < v; ++a)
      for (int j = 0; j < o; ++j)
        for (int i = 0; i < o; ++i) {
          double s = 0.0;
          for (int e = 0; e < v; ++e)
            for (int m = 0; m < o; ++m)
              s += -t2[IX(m, j, a, e, o, o, v)] * I_ovov[IX(i, e, m, b, o, v, o)] -
                   I_ovov[IX(i, e, m, a, o, v, o)] * t2[IX(m, j, e, b, o, o, v)] +
                   asym[IX(m, i, e, a, o, o, v)] * I_voov[IX(e, j, m, b, v, o, o)];
          tmp_t2[IX(i, j, a, b, o, o, v)] += s;
        }
  LAP(5)
  gemm('N', 'N', o, (bint)o * v * v, v, t1, I_vovv_p, tmp_t2, 1.0, 1.0);     /* :1700 */
#pragma omp parallel for schedule(static, 10) collapse(2)
  for (int b = 0; b < v; ++b)                                                 /* :1702-1713 */
    for (int a = 0; a < v; ++a)
      for (int j = 0; j < o; ++j)
        for (int i = 0; i < o; ++i)
          for (int m = 0; m < o; ++m) tmp_t2[IX(i, j, a, b, o, o, v)] -= t1[m + o * a] * I_ooov_p[IX(i, j, m, b, o, o, o)];
  LAP(6)
  rt = dalloc(oovv);
  afesp_ref_reshape(rt, tmp_t2, d_oovv, "2143", 0, 0.0);                      /* :1719-1722 */
  for (size_t q = 0; q < oovv; ++q) tmp_t2[q] = tmp_t2[q] + rt[q] + v_oovv[q];
  free(rt);
  for (int a = 0; a < v; ++a)                                                 /* :1727-1728 (serial in the reference) */
    for (int i = 0; i < o; ++i) t1[i + o * a] = tmp_t1[i + o * a] / (eps[i] - eps[o + a]);
  for (int b = 0; b < v; ++b)
    for (int a = 0; a < v; ++a)
      for (int j = 0; j < o; ++j)
        for (int i = 0; i < o; ++i)
          t2[IX(i, j, a, b, o, o, v)] = tmp_t2[IX(i, j, a, b, o, o, v)] / (eps[i] + eps[j] - eps[o + a] - eps[o + b]);
  LAP(7)
#undef LAP
  free(asym); free(c); free(I_vo); free(I_vv); free(I_oo_p); free(I_oo); free(I_oooo); free(I_ovov); free(I_voov);
  free(I_vovv_p); free(x_voov); free(I_ooov_p); free(tmp_t1); free(tmp_t2);
}

/* ------------------------------------------------------------------------------------------------------------------
 * AO -> MO (src/mp2.f90:321-410).  eri_ind (src/integrals.f90:196-210), 0-based here. */
static inline size_t eri_ind(size_t i, size_t j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }

/* C is sys%canon_coeff(mo, ao), column-major n x n: C(p,i) = Cm[p + n*i].
 * lmax == n and smax == n: the complete transform; eri_mo (packed, may be NULL) is then filled by the serial repack of
 * :388-410.  lmax < n: quarter transforms 1-3 on the slab l < lmax only (tmp arrays n^3*lmax); the 4th, whose l loop
 * is a contraction, then runs for s < smax over the available slab cyclically (l mod lmax) -- a timing sample of the
 * same plane-axpy loop structure on fewer distinct planes (cache-friendlier than the real thing: a lower bound).
 * times[0..4]: seconds of the four quarter transforms and the repack. */
void afesp_ref_ao2mo(int n, const double* eri, const double* Cm, int lmax, int smax, double* eri_mo, double* times) {
  if (lmax <= 0 || lmax > n) lmax = n;
  if (smax <= 0 || smax > n) smax = n;
  const size_t n3 = (size_t)n * n * n;
  const int L = lmax;
  const int S = (lmax == n) ? n : smax;   /* extent of the last axis of tmp_b in the 4th transform */
  const size_t nb = n3 * (size_t)(L > S ? L : S);
  double* tmp_a = (double*)calloc(n3 * L, sizeof(double));
  double* tmp_b = (double*)calloc(nb, sizeof(double));
  double t0 = now(), tt;
#define A4(p, q, r, s) tmp_a[IX(p, q, r, s, n, n, n)]
#define B4(p, q, r, s) tmp_b[IX(p, q, r, s, n, n, n)]
#define CM(p, i) Cm[(p) + (size_t)n * (i)]
#pragma omp parallel
  {
#pragma omp for schedule(static, 10) collapse(2)
    for (int l = 0; l < lmax; ++l)                       /* (ij|kl) -> (pj|kl)   :321-334 */
      for (int k = 0; k < n; ++k) {
        const size_t kl = eri_ind(k, l);
        for (int i = 0; i < n; ++i)
          for (int j = 0; j < n; ++j) {
            const size_t ij = eri_ind(i, j);
            const double x = eri[eri_ind(ij, kl)];
            for (int p = 0; p < n; ++p) A4(p, j, k, l) += x * CM(p, i);
          }
      }
#pragma omp single
    { tt = now(); times[0] = tt - t0; t0 = tt; }
#pragma omp for schedule(static, 10) collapse(2)
    for (int l = 0; l < lmax; ++l)                       /* (pj|kl) -> (pq|kl)   :338-350 */
      for (int k = 0; k < n; ++k)
        for (int j = 0; j < n; ++j)
          for (int q = 0; q < n; ++q)
            for (int p = 0; p < n; ++p) B4(p, q, k, l) += A4(p, j, k, l) * CM(q, j);
#pragma omp single
    { memset(tmp_a, 0, n3 * L * sizeof(double)); tt = now(); times[1] = tt - t0; t0 = tt; }
#pragma omp for schedule(static, 10) collapse(2)
    for (int l = 0; l < lmax; ++l)                       /* (pq|kl) -> (pq|rl)   :357-369 */
      for (int r = 0; r < n; ++r)
        for (int k = 0; k < n; ++k)
          for (int q = 0; q < n; ++q)
            for (int p = 0; p < n; ++p) A4(p, q, r, l) += B4(p, q, k, l) * CM(r, k);
#pragma omp single
    { memset(tmp_b, 0, nb * sizeof(double)); tt = now(); times[2] = tt - t0; t0 = tt; }
#pragma omp for schedule(static, 10) collapse(2)
    for (int s = 0; s < smax; ++s)                       /* (pq|rl) -> (pq|rs)   :375-387 */
      for (int r = 0; r < n; ++r)
        for (int l = 0; l < n; ++l)
          for (int q = 0; q < n; ++q)
            for (int p = 0; p < n; ++p) B4(p, q, r, s) += A4(p, q, r, l % L) * CM(s, l);
  }
  tt = now(); times[3] = tt - t0; t0 = tt;
  if (eri_mo && lmax == n && smax == n) {                /* serial repack   :388-410 */
    size_t pqrs = 0;
    for (int p = 0; p < n; ++p)
      for (int q = 0; q <= p; ++q)
        for (int r = 0; r <= p; ++r) {
          const int s_up = (p == r) ? q : r;
          for (int s = 0; s <= s_up; ++s) eri_mo[pqrs++] = B4(s, r, q, p);
        }
  }
  times[4] = now() - t0;
#undef A4
#undef B4
#undef CM
  free(tmp_a); free(tmp_b);
}

/* init_cc slice gather (src/ccsd.f90:496-512): out(p,q,r,s) = eri_mo(eri_ind(eri_ind(p,r), eri_ind(q,s))) over the
 * orbital ranges [lo_x, lo_x + n_x); s restricted to [s0, s0 + ns) of its range (column blocks of v_vvvv). */
void afesp_ref_slice(const double* eri_mo, int lo_p, int n_p, int lo_q, int n_q, int lo_r, int n_r, int lo_s, int ns,
                     double* out) {
#pragma omp parallel for collapse(2)
  for (int s = 0; s < ns; ++s)
    for (int r = 0; r < n_r; ++r)
      for (int q = 0; q < n_q; ++q)
        for (int p = 0; p < n_p; ++p)
          out[IX(p, q, r, s, n_p, n_q, n_r)] = eri_mo[eri_ind(eri_ind(lo_p + p, lo_r + r), eri_ind(lo_q + q, lo_s + s))];
}
