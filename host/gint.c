/* gint.c -- Gaussian one- and two-electron integrals over contracted, real solid-harmonic shells (s, p, d, f).
 *
 * Host-side tool of the drop-in: the reference reads s.dat / t.dat / v.dat / eri.dat produced by
 * Psi4 (utils/psi4_integrals_nosym.py of the reference tree), and its checkout ships no eri.dat for
 * sample_data/h2o-cc-pvtz (listed in .MISSING_LARGE_BLOBS).  This file regenerates all four from geom.dat and the public
 * basis-set parameters, in Psi4's conventions, so that the cc-pVTZ sample can be run: contracted functions normalised to
 * unit self-overlap, pure functions ordered m = 0, +1, -1, +2, -2, ... within a shell (p: z, x, y), shells in basis-file
 * order atom by atom.  Validated against the shipped s/t/v.dat (and against the shipped eri.dat of the cc-pVDZ samples)
 * in tests/test_gint.py.
 *
 * Method: McMurchie-Davidson.  Products of Cartesian Gaussians are expanded in Hermite Gaussians (coefficients E_t^{ij},
 * two-term recursions), Coulomb-type integrals reduce to Hermite integrals R_tuv built from the Boys function F_n
 * (series + downward recursion for T < 30, erf + upward recursion above), and the Cartesian blocks of a contracted shell
 * quartet are transformed to solid harmonics at the end.  Plain C + OpenMP; everything in double precision.
 *
 *   pair index(i,j) = i(i+1)/2 + j (i >= j, 0-based); eri[index(index(i,j), index(k,l))] = (ij|kl)   (src/integrals.f90:196-210)
 */
#include <math.h>
#include <omp.h>
#include <stdlib.h>
#include <string.h>

#define LMAX 3
#define NCART(l) (((l) + 1) * ((l) + 2) / 2)
#define MAXCART 10
#define MAXPRIM 16
#define MAXL4 (4 * LMAX)

typedef struct {
  double c[3];
  int l, nprim, first; /* first = index of the shell's first (pure) basis function */
  double a[MAXPRIM], d[MAXPRIM]; /* exponents, contraction coefficients including the primitive norm */
} shell_t;

/* ---- Cartesian component tables and Cartesian -> real solid harmonic coefficients (unnormalised; every contracted pure
 *      function is normalised numerically afterwards).  Order of the pure functions: m = 0, +1, -1, +2, -2, +3, -3. */
static int cart_l[LMAX + 1][MAXCART][3];
static double sph[LMAX + 1][2 * LMAX + 1][MAXCART];
static int tables_ready = 0;

static int cart_index(int l, int lx, int ly, int lz) {
  for (int q = 0; q < NCART(l); ++q)
    if (cart_l[l][q][0] == lx && cart_l[l][q][1] == ly && cart_l[l][q][2] == lz) return q;
  return -1;
}

static void build_tables(void) {
  if (tables_ready) return;
  for (int l = 0; l <= LMAX; ++l) {
    int q = 0;
    for (int lx = l; lx >= 0; --lx)
      for (int ly = l - lx; ly >= 0; --ly) {
        cart_l[l][q][0] = lx; cart_l[l][q][1] = ly; cart_l[l][q][2] = l - lx - ly;
        ++q;
      }
  }
  memset(sph, 0, sizeof(sph));
#define S(l, m, lx, ly, lz, v) sph[l][m][cart_index(l, lx, ly, lz)] = (v)
  S(0, 0, 0, 0, 0, 1.0);
  S(1, 0, 0, 0, 1, 1.0); S(1, 1, 1, 0, 0, 1.0); S(1, 2, 0, 1, 0, 1.0);                       /* z, x, y */
  S(2, 0, 0, 0, 2, 1.0); S(2, 0, 2, 0, 0, -0.5); S(2, 0, 0, 2, 0, -0.5);                       /* zz - (xx+yy)/2 */
  S(2, 1, 1, 0, 1, 1.0);                                                                       /* xz */
  S(2, 2, 0, 1, 1, 1.0);                                                                       /* yz */
  S(2, 3, 2, 0, 0, 1.0); S(2, 3, 0, 2, 0, -1.0);                                               /* xx - yy */
  S(2, 4, 1, 1, 0, 1.0);                                                                       /* xy */
  S(3, 0, 0, 0, 3, 2.0); S(3, 0, 2, 0, 1, -3.0); S(3, 0, 0, 2, 1, -3.0);                       /* z(2zz - 3xx - 3yy) */
  S(3, 1, 1, 0, 2, 4.0); S(3, 1, 3, 0, 0, -1.0); S(3, 1, 1, 2, 0, -1.0);                       /* x(4zz - xx - yy) */
  S(3, 2, 0, 1, 2, 4.0); S(3, 2, 2, 1, 0, -1.0); S(3, 2, 0, 3, 0, -1.0);                       /* y(4zz - xx - yy) */
  S(3, 3, 2, 0, 1, 1.0); S(3, 3, 0, 2, 1, -1.0);                                               /* z(xx - yy) */
  S(3, 4, 1, 1, 1, 1.0);                                                                       /* xyz */
  S(3, 5, 3, 0, 0, 1.0); S(3, 5, 1, 2, 0, -3.0);                                               /* x(xx - 3yy) */
  S(3, 6, 2, 1, 0, 3.0); S(3, 6, 0, 3, 0, -1.0);                                               /* y(3xx - yy) */
#undef S
  tables_ready = 1;
}

/* ---- Boys function F_n(T), n = 0..nmax */
static void boys(int nmax, double T, double* F) {
  if (T < 30.0) {
    /* F_nmax by its (all-positive) series, then downward recursion F_{n-1} = (2T F_n + e^-T)/(2n-1) */
    const double eT = exp(-T);
    double term = 1.0 / (2 * nmax + 1), sum = term;
    for (int k = 1; k < 400; ++k) {
      term *= 2.0 * T / (2 * nmax + 2 * k + 1);
      sum += term;
      if (term < 1e-18 * sum) break;
    }
    F[nmax] = eT * sum;
    for (int n = nmax; n > 0; --n) F[n - 1] = (2.0 * T * F[n] + eT) / (2 * n - 1);
  } else {
    const double eT = exp(-T);
    F[0] = 0.5 * sqrt(M_PI / T) * erf(sqrt(T));
    for (int n = 0; n < nmax; ++n) F[n + 1] = ((2 * n + 1) * F[n] - eT) / (2.0 * T);
  }
}

/* ---- Hermite expansion coefficients in one dimension: E[i][j][t], 0 <= i <= la, 0 <= j <= lb, 0 <= t <= i+j */
typedef double ecoef_t[LMAX + 3][LMAX + 3][2 * LMAX + 5];
static void hermite_E(int la, int lb, double a, double b, double A, double B, ecoef_t E) {
  const double p = a + b, mu = a * b / p, P = (a * A + b * B) / p, XPA = P - A, XPB = P - B, XAB = A - B;
  for (int i = 0; i <= la; ++i)
    for (int j = 0; j <= lb; ++j)
      for (int t = 0; t <= la + lb + 1; ++t) E[i][j][t] = 0.0;
  E[0][0][0] = exp(-mu * XAB * XAB);
  for (int i = 0; i <= la; ++i) {
    if (i > 0)
      for (int t = 0; t <= i; ++t)
        E[i][0][t] = (t > 0 ? E[i - 1][0][t - 1] / (2 * p) : 0.0) + XPA * E[i - 1][0][t] + (t + 1) * E[i - 1][0][t + 1];
    for (int j = 1; j <= lb; ++j)
      for (int t = 0; t <= i + j; ++t)
        E[i][j][t] = (t > 0 ? E[i][j - 1][t - 1] / (2 * p) : 0.0) + XPB * E[i][j - 1][t] + (t + 1) * E[i][j - 1][t + 1];
  }
}

/* ---- Hermite Coulomb integrals R[t][u][v] = R^0_tuv(alpha, PC), t+u+v <= L */
#define RDIM (MAXL4 + 1)
static void hermite_R(int L, double alpha, const double* PC, double R[RDIM][RDIM][RDIM]) {
  double F[MAXL4 + 2];
  static _Thread_local double Rn[MAXL4 + 1][RDIM][RDIM][RDIM];
  const double T = alpha * (PC[0] * PC[0] + PC[1] * PC[1] + PC[2] * PC[2]);
  boys(L, T, F);
  double f = 1.0;
  for (int n = 0; n <= L; ++n) { Rn[n][0][0][0] = f * F[n]; f *= -2.0 * alpha; }
  /* build orders downwards: R^n_{tuv} for t+u+v <= L-n */
  for (int n = L - 1; n >= 0; --n) {
    const int m = L - n;
    for (int t = 0; t <= m; ++t)
      for (int u = 0; u + t <= m; ++u)
        for (int v = 0; v + u + t <= m; ++v) {
          if (t + u + v == 0) continue;
          double val;
          if (t > 0) val = (t > 1 ? (t - 1) * Rn[n + 1][t - 2][u][v] : 0.0) + PC[0] * Rn[n + 1][t - 1][u][v];
          else if (u > 0) val = (u > 1 ? (u - 1) * Rn[n + 1][t][u - 2][v] : 0.0) + PC[1] * Rn[n + 1][t][u - 1][v];
          else val = (v > 1 ? (v - 1) * Rn[n + 1][t][u][v - 2] : 0.0) + PC[2] * Rn[n + 1][t][u][v - 1];
          Rn[n][t][u][v] = val;
        }
  }
  for (int t = 0; t <= L; ++t)
    for (int u = 0; u + t <= L; ++u)
      for (int v = 0; v + u + t <= L; ++v) R[t][u][v] = Rn[0][t][u][v];
}

static double prim_norm(double a, int l) { /* a-dependence of the normalisation of x^l exp(-a r^2); constants drop out */
  return pow(2.0 * a / M_PI, 0.75) * pow(4.0 * a, 0.5 * l);
}

/* Pure-function block from a Cartesian block: out[ma][mb] = sum sph_a[ma][ca] sph_b[mb][cb] in[ca][cb] */
static void to_pure2(int la, int lb, const double* in, double* out) {
  const int na = NCART(la), nb = NCART(lb), pa = 2 * la + 1, pb = 2 * lb + 1;
  for (int ma = 0; ma < pa; ++ma)
    for (int mb = 0; mb < pb; ++mb) {
      double s = 0.0;
      for (int ca = 0; ca < na; ++ca) {
        if (sph[la][ma][ca] == 0.0) continue;
        for (int cb = 0; cb < nb; ++cb) s += sph[la][ma][ca] * sph[lb][mb][cb] * in[ca * nb + cb];
      }
      out[ma * pb + mb] = s;
    }
}

/* One-electron Cartesian blocks of a shell pair: overlap S, kinetic T, nuclear attraction V (all nuclei). */
static void shell_pair_1e(const shell_t* A, const shell_t* B, int natom, const double* Z, const double* xyz, double* Sb,
                          double* Tb, double* Vb) {
  const int la = A->l, lb = B->l, na = NCART(la), nb = NCART(lb);
  memset(Sb, 0, sizeof(double) * na * nb); memset(Tb, 0, sizeof(double) * na * nb); memset(Vb, 0, sizeof(double) * na * nb);
  static _Thread_local double R[RDIM][RDIM][RDIM];
  for (int ia = 0; ia < A->nprim; ++ia)
    for (int ib = 0; ib < B->nprim; ++ib) {
      const double a = A->a[ia], b = B->a[ib], p = a + b, cc = A->d[ia] * B->d[ib];
      ecoef_t E[3];
      for (int x = 0; x < 3; ++x) hermite_E(la, lb + 2, a, b, A->c[x], B->c[x], E[x]);
      const double pref = pow(M_PI / p, 1.5);
      double P[3];
      for (int x = 0; x < 3; ++x) P[x] = (a * A->c[x] + b * B->c[x]) / p;
      for (int ca = 0; ca < na; ++ca)
        for (int cb = 0; cb < nb; ++cb) {
          const int* u = cart_l[la][ca];
          const int* w = cart_l[lb][cb];
          double s1[3], d2[3];
          for (int x = 0; x < 3; ++x) {
            const int j = w[x];
            s1[x] = E[x][u[x]][j][0];
            d2[x] = (j > 1 ? j * (j - 1) * E[x][u[x]][j - 2][0] : 0.0) - 2.0 * b * (2 * j + 1) * E[x][u[x]][j][0] +
                    4.0 * b * b * E[x][u[x]][j + 2][0];
          }
          Sb[ca * nb + cb] += cc * pref * s1[0] * s1[1] * s1[2];
          Tb[ca * nb + cb] += cc * pref * (-0.5) * (d2[0] * s1[1] * s1[2] + s1[0] * d2[1] * s1[2] + s1[0] * s1[1] * d2[2]);
        }
      for (int c = 0; c < natom; ++c) {
        double PC[3] = {P[0] - xyz[3 * c], P[1] - xyz[3 * c + 1], P[2] - xyz[3 * c + 2]};
        hermite_R(la + lb, p, PC, R);
        const double f = -Z[c] * 2.0 * M_PI / p * cc;
        for (int ca = 0; ca < na; ++ca)
          for (int cb = 0; cb < nb; ++cb) {
            const int* u = cart_l[la][ca];
            const int* w = cart_l[lb][cb];
            double s = 0.0;
            for (int t = 0; t <= u[0] + w[0]; ++t)
              for (int uu = 0; uu <= u[1] + w[1]; ++uu)
                for (int v = 0; v <= u[2] + w[2]; ++v)
                  s += E[0][u[0]][w[0]][t] * E[1][u[1]][w[1]][uu] * E[2][u[2]][w[2]][v] * R[t][uu][v];
            Vb[ca * nb + cb] += f * s;
          }
      }
    }
}

/* Cartesian ERI block (ab|cd) of a contracted shell quartet: out[ca][cb][cc][cd] */
static void shell_quartet(const shell_t* A, const shell_t* B, const shell_t* C, const shell_t* D, double* out) {
  const int la = A->l, lb = B->l, lc = C->l, ld = D->l;
  const int na = NCART(la), nb = NCART(lb), nc = NCART(lc), nd = NCART(ld);
  const int Lab = la + lb, Lcd = lc + ld, L = Lab + Lcd;
  memset(out, 0, sizeof(double) * na * nb * nc * nd);
  static _Thread_local double R[RDIM][RDIM][RDIM];
  static _Thread_local double G[(2 * LMAX + 1) * (2 * LMAX + 1) * (2 * LMAX + 1)][MAXCART * MAXCART];
  const int tdim = Lab + 1;
  for (int ia = 0; ia < A->nprim; ++ia)
    for (int ib = 0; ib < B->nprim; ++ib) {
      const double a = A->a[ia], b = B->a[ib], p = a + b;
      ecoef_t Eab[3];
      double P[3];
      for (int x = 0; x < 3; ++x) { hermite_E(la, lb, a, b, A->c[x], B->c[x], Eab[x]); P[x] = (a * A->c[x] + b * B->c[x]) / p; }
      const double cab = A->d[ia] * B->d[ib];
      for (int ic = 0; ic < C->nprim; ++ic)
        for (int id = 0; id < D->nprim; ++id) {
          const double c = C->a[ic], d = D->a[id], q = c + d;
          ecoef_t Ecd[3];
          double Q[3], PQ[3];
          for (int x = 0; x < 3; ++x) {
            hermite_E(lc, ld, c, d, C->c[x], D->c[x], Ecd[x]);
            Q[x] = (c * C->c[x] + d * D->c[x]) / q; PQ[x] = P[x] - Q[x];
          }
          const double alpha = p * q / (p + q);
          hermite_R(L, alpha, PQ, R);
          const double pref = 2.0 * pow(M_PI, 2.5) / (p * q * sqrt(p + q)) * cab * C->d[ic] * D->d[id];
          /* G[tuv][cd] = sum_{tau nu phi} (-1)^(tau+nu+phi) Ecd R[t+tau][u+nu][v+phi] */
          for (int t = 0; t <= Lab; ++t)
            for (int u = 0; u + t <= Lab; ++u)
              for (int v = 0; v + u + t <= Lab; ++v) {
                double* g = G[(t * tdim + u) * tdim + v];
                for (int cc = 0; cc < nc; ++cc)
                  for (int cd = 0; cd < nd; ++cd) {
                    const int* m = cart_l[lc][cc];
                    const int* n = cart_l[ld][cd];
                    double s = 0.0;
                    for (int tau = 0; tau <= m[0] + n[0]; ++tau) {
                      const double ex = Ecd[0][m[0]][n[0]][tau];
                      for (int nu = 0; nu <= m[1] + n[1]; ++nu) {
                        const double exy = ex * Ecd[1][m[1]][n[1]][nu];
                        for (int phi = 0; phi <= m[2] + n[2]; ++phi) {
                          const double sg = ((tau + nu + phi) & 1) ? -1.0 : 1.0;
                          s += sg * exy * Ecd[2][m[2]][n[2]][phi] * R[t + tau][u + nu][v + phi];
                        }
                      }
                    }
                    g[cc * nd + cd] = s;
                  }
              }
          for (int ca = 0; ca < na; ++ca)
            for (int cb = 0; cb < nb; ++cb) {
              const int* i = cart_l[la][ca];
              const int* j = cart_l[lb][cb];
              double* o = out + ((size_t)(ca * nb + cb)) * nc * nd;
              for (int t = 0; t <= i[0] + j[0]; ++t) {
                const double ex = Eab[0][i[0]][j[0]][t];
                for (int u = 0; u <= i[1] + j[1]; ++u) {
                  const double exy = ex * Eab[1][i[1]][j[1]][u];
                  for (int v = 0; v <= i[2] + j[2]; ++v) {
                    const double e = pref * exy * Eab[2][i[2]][j[2]][v];
                    const double* g = G[(t * tdim + u) * tdim + v];
                    for (int x = 0; x < nc * nd; ++x) o[x] += e * g[x];
                  }
                }
              }
            }
        }
    }
}

static size_t tri(size_t i, size_t j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }

/* nshell shells described by flat arrays (per shell: atom index, l, nprim, offset into exps/coefs); natom nuclei.
 * Outputs: S, T, V as dense nbf x nbf (row-major, symmetric) and, when eri != NULL, the packed two-electron integrals
 * (npair(npair+1)/2 doubles).  Returns nbf (or -1 on a bad argument).  Basis functions follow shell order. */
int afesp_gint_compute(int natom, const double* Z, const double* xyz, int nshell, const int* sh_atom, const int* sh_l,
                       const int* sh_nprim, const int* sh_off, const double* exps, const double* coefs, double* Smat,
                       double* Tmat, double* Vmat, double* eri) {
  build_tables();
  shell_t* sh = (shell_t*)calloc(nshell, sizeof(shell_t));
  int nbf = 0;
  for (int s = 0; s < nshell; ++s) {
    if (sh_l[s] < 0 || sh_l[s] > LMAX || sh_nprim[s] < 1 || sh_nprim[s] > MAXPRIM || sh_atom[s] < 0 || sh_atom[s] >= natom) {
      free(sh);
      return -1;
    }
    for (int x = 0; x < 3; ++x) sh[s].c[x] = xyz[3 * sh_atom[s] + x];
    sh[s].l = sh_l[s]; sh[s].nprim = sh_nprim[s]; sh[s].first = nbf;
    for (int k = 0; k < sh_nprim[s]; ++k) {
      sh[s].a[k] = exps[sh_off[s] + k];
      sh[s].d[k] = coefs[sh_off[s] + k] * prim_norm(sh[s].a[k], sh[s].l);
    }
    nbf += 2 * sh[s].l + 1;
  }
  /* norms of the contracted pure functions from their own self-overlap */
  double* nrm = (double*)malloc(sizeof(double) * nbf);
  for (int s = 0; s < nshell; ++s) {
    double Sb[MAXCART * MAXCART], Tb[MAXCART * MAXCART], Vb[MAXCART * MAXCART], P[49];
    shell_pair_1e(&sh[s], &sh[s], 0, Z, xyz, Sb, Tb, Vb);
    to_pure2(sh[s].l, sh[s].l, Sb, P);
    const int np = 2 * sh[s].l + 1;
    for (int m = 0; m < np; ++m) nrm[sh[s].first + m] = 1.0 / sqrt(P[m * np + m]);
  }
#pragma omp parallel for schedule(dynamic, 1)
  for (int sa = 0; sa < nshell; ++sa)
    for (int sb = 0; sb <= sa; ++sb) {
      double Sb[MAXCART * MAXCART], Tb[MAXCART * MAXCART], Vb[MAXCART * MAXCART], Ps[49], Pt[49], Pv[49];
      shell_pair_1e(&sh[sa], &sh[sb], natom, Z, xyz, Sb, Tb, Vb);
      to_pure2(sh[sa].l, sh[sb].l, Sb, Ps); to_pure2(sh[sa].l, sh[sb].l, Tb, Pt); to_pure2(sh[sa].l, sh[sb].l, Vb, Pv);
      const int pa = 2 * sh[sa].l + 1, pb = 2 * sh[sb].l + 1;
      for (int ma = 0; ma < pa; ++ma)
        for (int mb = 0; mb < pb; ++mb) {
          const int i = sh[sa].first + ma, j = sh[sb].first + mb;
          const double f = nrm[i] * nrm[j];
          Smat[(size_t)i * nbf + j] = Smat[(size_t)j * nbf + i] = f * Ps[ma * pb + mb];
          Tmat[(size_t)i * nbf + j] = Tmat[(size_t)j * nbf + i] = f * Pt[ma * pb + mb];
          Vmat[(size_t)i * nbf + j] = Vmat[(size_t)j * nbf + i] = f * Pv[ma * pb + mb];
        }
    }
  if (eri) {
    /* unique shell quartets (sa >= sb, sc >= sd, (sa,sb) >= (sc,sd)); every function quartet of the block is written to
     * its canonical packed slot (duplicates within a block write the same value) */
    const int npairs = nshell * (nshell + 1) / 2;
#pragma omp parallel
    {
      double* cart = (double*)malloc(sizeof(double) * MAXCART * MAXCART * MAXCART * MAXCART);
      double* t1 = (double*)malloc(sizeof(double) * MAXCART * MAXCART * MAXCART * MAXCART);
      double* t2 = (double*)malloc(sizeof(double) * MAXCART * MAXCART * MAXCART * MAXCART);
#pragma omp for schedule(dynamic, 1)
      for (int ab = npairs - 1; ab >= 0; --ab) {
        int sa = (int)((sqrt(8.0 * ab + 1.0) - 1.0) / 2.0);
        while ((sa + 1) * (sa + 2) / 2 <= ab) ++sa;
        while (sa * (sa + 1) / 2 > ab) --sa;
        const int sb = ab - sa * (sa + 1) / 2;
        for (int cd = 0; cd <= ab; ++cd) {
          int sc = (int)((sqrt(8.0 * cd + 1.0) - 1.0) / 2.0);
          while ((sc + 1) * (sc + 2) / 2 <= cd) ++sc;
          while (sc * (sc + 1) / 2 > cd) --sc;
          const int sd = cd - sc * (sc + 1) / 2;
          const shell_t *A = &sh[sa], *B = &sh[sb], *C = &sh[sc], *D = &sh[sd];
          shell_quartet(A, B, C, D, cart);
          const int n[4] = {NCART(A->l), NCART(B->l), NCART(C->l), NCART(D->l)};
          const int pz[4] = {2 * A->l + 1, 2 * B->l + 1, 2 * C->l + 1, 2 * D->l + 1};
          const int ls[4] = {A->l, B->l, C->l, D->l};
          /* transform the four indices one after the other: in[x0][x1][x2][x3] -> out[x1][x2][x3][m0] (cyclic) */
          int dims[4] = {n[0], n[1], n[2], n[3]};
          double *src = cart, *dst = t1;
          for (int pass = 0; pass < 4; ++pass) {
            const int l = ls[pass], nc0 = dims[0], rest = dims[1] * dims[2] * dims[3], np0 = pz[pass];
            for (int r = 0; r < rest; ++r)
              for (int m = 0; m < np0; ++m) {
                double s = 0.0;
                for (int c0 = 0; c0 < nc0; ++c0) s += sph[l][m][c0] * src[(size_t)c0 * rest + r];
                dst[(size_t)r * np0 + m] = s;
              }
            dims[0] = dims[1]; dims[1] = dims[2]; dims[2] = dims[3]; dims[3] = np0;
            src = dst; dst = (dst == t1) ? t2 : t1;
          }
          /* src now holds [ma][mb][mc][md] */
          for (int ma = 0; ma < pz[0]; ++ma)
            for (int mb = 0; mb < pz[1]; ++mb)
              for (int mc = 0; mc < pz[2]; ++mc)
                for (int md = 0; md < pz[3]; ++md) {
                  const int i = A->first + ma, j = B->first + mb, k = C->first + mc, l2 = D->first + md;
                  const double val = nrm[i] * nrm[j] * nrm[k] * nrm[l2] * src[((ma * pz[1] + mb) * pz[2] + mc) * pz[3] + md];
                  eri[tri(tri(i, j), tri(k, l2))] = val;
                }
        }
      }
      free(cart); free(t1); free(t2);
    }
  }
  free(nrm); free(sh);
  return nbf;
}

/* ---- built-in basis sets (host/gint_basis.inc, generated from afesp_b200/gint.py) ------------------------------------
 * afesp_gint_named: the same computation with the shells looked up by basis-set name ("cc-pvdz", "cc-pvtz", "def2-svp";
 * H and O; cc-pvdz also N and F), atom by atom in basis-file order -- what the C++ host program uses when a run directory
 * has geom.dat but no eri.dat.  Smat/Tmat/Vmat/eri may each be NULL (nothing written); returns nbf, -1 for a bad argument, -2 for an
 * unknown basis / element.  Call with all outputs NULL to learn nbf. */
#include "gint_basis.inc"
#include <ctype.h>

static int same_name(const char* a, const char* b) {
  for (; *a && *b; ++a, ++b)
    if (tolower((unsigned char)*a) != tolower((unsigned char)*b)) return 0;
  return *a == *b;
}

int afesp_gint_named(int natom, const double* Z, const double* xyz, const char* basis, double* Smat, double* Tmat,
                     double* Vmat, double* eri) {
  if (natom <= 0 || !Z || !xyz || !basis) return -1;
  int nshell = 0, nprim = 0, nbf = 0;
  for (int at = 0; at < natom; ++at) {
    int found = 0;
    for (int q = 0; q < basis_table_len; ++q)
      if (basis_table[q].Z == (int)(Z[at] + 0.5) && same_name(basis_table[q].basis, basis)) {
        ++nshell; nprim += basis_table[q].nprim; nbf += 2 * basis_table[q].l + 1; found = 1;
      }
    if (!found) return -2;
  }
  if (!Smat && !Tmat && !Vmat && !eri) return nbf;
  int* sh_atom = (int*)malloc(sizeof(int) * nshell * 4);
  int *sh_l = sh_atom + nshell, *sh_np = sh_l + nshell, *sh_off = sh_np + nshell;
  double* ex = (double*)malloc(sizeof(double) * nprim * 2);
  double* co = ex + nprim;
  int s = 0, off = 0;
  for (int at = 0; at < natom; ++at)
    for (int q = 0; q < basis_table_len; ++q)
      if (basis_table[q].Z == (int)(Z[at] + 0.5) && same_name(basis_table[q].basis, basis)) {
        sh_atom[s] = at; sh_l[s] = basis_table[q].l; sh_np[s] = basis_table[q].nprim; sh_off[s] = off;
        for (int k = 0; k < basis_table[q].nprim; ++k) { ex[off + k] = basis_table[q].a[k]; co[off + k] = basis_table[q].c[k]; }
        off += basis_table[q].nprim; ++s;
      }
  double* scratch = (double*)calloc((size_t)3 * nbf * nbf, sizeof(double));
  const int rc = afesp_gint_compute(natom, Z, xyz, nshell, sh_atom, sh_l, sh_np, sh_off, ex, co,
                                    Smat ? Smat : scratch, Tmat ? Tmat : scratch + (size_t)nbf * nbf,
                                    Vmat ? Vmat : scratch + (size_t)2 * nbf * nbf, eri);
  free(scratch); free(ex); free(sh_atom);
  return rc;
}
