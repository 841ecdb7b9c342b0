// Energy / convergence quantities and CC-DIIS shared by the spin-free and spin-orbital drivers.
#include <algorithm>
#include <cmath>
#include <vector>

#include "ccsd.cuh"

namespace afesp {

void CCDiis::init(int nerr_, int o, int v) {
  AFESP_REQUIRE(nerr_ >= 0 && nerr_ <= 64, "CC-DIIS: ccsd_diis_n_errmat must lie in 0..64");   // (2 x nerr amplitude-sized arrays)
  nerr = nerr_;
  use = nerr >= 2;  // src/ccsd.f90:593-595
  slot = 0; n_active = 0;
  t1.clear(); t2.clear(); e1.clear(); e2.clear();
  if (!use) return;
  for (int k = 0; k < nerr; ++k) {
    t1.emplace_back(std::vector<int>{o, v});
    e1.emplace_back(std::vector<int>{o, v});
    t2.emplace_back(std::vector<int>{o, o, v, v});
    e2.emplace_back(std::vector<int>{o, o, v, v});
  }
  t1_s.init({o, v});
  t2_s.init({o, o, v, v});
  B.assign((size_t)nerr * nerr, 0.0);
}

// src/ccsd.f90:1734-1810: E_CC and sum (t2 - t2_old)^2 in one pass; t2_old <- t2.
void cc_update_energy(CCState& s) {
  if (s.red_out.n < 16) s.red_out.alloc(16);
  const double* vint = s.restricted ? s.get("v_oovv").p() : s.get("oovv").p();
  if (s.restricted) cc_energy_restricted(s.eng, vint, s.t2.p(), s.t1.p(), s.t2_old.p(), s.o, s.v, s.red_out.p);
  else cc_energy_spinorb(s.eng, vint, s.t2.p(), s.t1.p(), s.t2_old.p(), s.o, s.v, s.red_out.p);
  double h[2];
  AFESP_CUDA_CHECK(cudaMemcpyAsync(h, s.red_out.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, s.eng.stream));
  AFESP_CUDA_CHECK(cudaMemcpyAsync(s.t2_old.p(), s.t2.p(), s.t2.size() * sizeof(double), cudaMemcpyDeviceToDevice,
                                   s.eng.stream));
  AFESP_CUDA_CHECK(cudaStreamSynchronize(s.eng.stream));
  s.energy_old = s.energy;
  s.energy = h[0];
  s.rms = h[1];  // Q4: the squared Frobenius norm, printed un-rooted (src/ccsd.f90:1806)
}

void cc_diis_stash(CCState& s) {
  if (!s.diis.use) return;
  AFESP_CUDA_CHECK(cudaMemcpyAsync(s.diis.t1_s.p(), s.t1.p(), s.t1.size() * 8, cudaMemcpyDeviceToDevice, s.eng.stream));
  AFESP_CUDA_CHECK(cudaMemcpyAsync(s.diis.t2_s.p(), s.t2.p(), s.t2.size() * 8, cudaMemcpyDeviceToDevice, s.eng.stream));
}

namespace {
// Dense solve by LU with partial pivoting (the reference calls LAPACK dsysv on the lower triangle,
// src/linalg.fpp:38-56; the system is at most 9x9).  Returns false on a singular pivot.
bool solve_dense(std::vector<double>& A, std::vector<double>& b, int n) {
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int r = c + 1; r < n; ++r) if (std::fabs(A[r * n + c]) > std::fabs(A[piv * n + c])) piv = r;
    if (A[piv * n + c] == 0.0) return false;
    if (piv != c) {
      for (int k = 0; k < n; ++k) std::swap(A[c * n + k], A[piv * n + k]);
      std::swap(b[c], b[piv]);
    }
    for (int r = c + 1; r < n; ++r) {
      double f = A[r * n + c] / A[c * n + c];
      if (f == 0.0) continue;
      for (int k = c; k < n; ++k) A[r * n + k] -= f * A[c * n + k];
      b[r] -= f * b[c];
    }
  }
  for (int r = n - 1; r >= 0; --r) {
    double x = b[r];
    for (int k = r + 1; k < n; ++k) x -= A[r * n + k] * b[k];
    b[r] = x / A[r * n + r];
  }
  return true;
}
}  // namespace

// src/ccsd.f90:617-676.  Only the row of B belonging to the new error vector is recomputed: the other entries are
// dot products of unchanged vectors and the reduction kernel is deterministic, so the matrix equals the reference's
// full recomputation.
void cc_diis_update(CCState& s) {
  CCDiis& d = s.diis;
  if (!d.use) return;
  d.slot += 1;
  if (d.slot > d.nerr) d.slot -= d.nerr;
  if (d.n_active < d.nerr) d.n_active += 1;
  const int k = d.slot - 1, n = d.n_active;
  cudaStream_t st = s.eng.stream;
  const long long n1 = s.t1.size(), n2 = s.t2.size();
  AFESP_CUDA_CHECK(cudaMemcpyAsync(d.t1[k].p(), s.t1.p(), n1 * 8, cudaMemcpyDeviceToDevice, st));
  AFESP_CUDA_CHECK(cudaMemcpyAsync(d.t2[k].p(), s.t2.p(), n2 * 8, cudaMemcpyDeviceToDevice, st));
  // e = T - T_s
  AFESP_CUDA_CHECK(cudaMemcpyAsync(d.e1[k].p(), s.t1.p(), n1 * 8, cudaMemcpyDeviceToDevice, st));
  AFESP_CUDA_CHECK(cudaMemcpyAsync(d.e2[k].p(), s.t2.p(), n2 * 8, cudaMemcpyDeviceToDevice, st));
  axpby(st, n1, -1.0, d.t1_s.p(), 1.0, d.e1[k].p());
  axpby(st, n2, -1.0, d.t2_s.p(), 1.0, d.e2[k].p());
  // new row of B: e_k . e_j for all active j (the reduction kernel takes up to 8 vectors per pass; the reference's usual
  // depth of 8 is one pass, deeper histories take ceil(n/8))
  const int npad = ((std::max(n, 1) + 7) / 8) * 8;
  std::vector<const double*> p1(npad, nullptr), p2(npad, nullptr);
  for (int j = 0; j < n; ++j) { p1[j] = d.e1[j].p(); p2[j] = d.e2[j].p(); }
  if (s.red_out.n < (size_t)std::max(16, 2 * npad)) s.red_out.alloc((size_t)std::max(16, 2 * npad));
  for (int j0 = 0; j0 < n; j0 += 8) {
    const int nx = std::min(8, n - j0);
    dotn(s.eng, n1, nx, p1.data() + j0, d.e1[k].p(), s.red_out.p + j0);
    dotn(s.eng, n2, nx, p2.data() + j0, d.e2[k].p(), s.red_out.p + npad + j0);
  }
  std::vector<double> h((size_t)2 * npad, 0.0);
  AFESP_CUDA_CHECK(cudaMemcpyAsync(h.data(), s.red_out.p, (size_t)2 * npad * 8, cudaMemcpyDeviceToHost, st));
  AFESP_CUDA_CHECK(cudaStreamSynchronize(st));
  for (int j = 0; j < n; ++j) {
    double val = h[j] + h[npad + j];
    d.B[(size_t)k * d.nerr + j] = val;
    d.B[(size_t)j * d.nerr + k] = val;
  }
  const int m = n + 1;
  std::vector<double> A((size_t)m * m, 0.0), rhs(m, 0.0);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) A[i * m + j] = d.B[(size_t)i * d.nerr + j];
  for (int i = 0; i < n; ++i) { A[i * m + n] = -1.0; A[n * m + i] = -1.0; }
  rhs[n] = -1.0;
  if (!solve_dense(A, rhs, m)) throw Error(3, "ccsd::update_diis_cc: Linear solve failed!");
  // T <- sum_i c_i T_i
  std::vector<const double*> q1(npad, nullptr), q2(npad, nullptr);
  for (int j = 0; j < n; ++j) { q1[j] = d.t1[j].p(); q2[j] = d.t2[j].p(); }
  const int n_first = std::min(8, n);
  lincomb(st, n1, n_first, q1.data(), rhs.data(), s.t1.p());
  lincomb(st, n2, n_first, q2.data(), rhs.data(), s.t2.p());
  for (int j = 8; j < n; ++j) {   // histories deeper than 8: the remaining terms are accumulated one by one
    axpby(st, n1, rhs[j], q1[j], 1.0, s.t1.p());
    axpby(st, n2, rhs[j], q2[j], 1.0, s.t2.p());
  }
}

double cc_t1_norm2(CCState& s) {
  if (s.red_out.n < 16) s.red_out.alloc(16);
  const double* p[1] = {s.t1.p()};
  dotn(s.eng, s.t1.size(), 1, p, s.t1.p(), s.red_out.p);
  double h = 0.0;
  AFESP_CUDA_CHECK(cudaMemcpyAsync(&h, s.red_out.p, 8, cudaMemcpyDeviceToHost, s.eng.stream));
  AFESP_CUDA_CHECK(cudaStreamSynchronize(s.eng.stream));
  return h;
}

}  // namespace afesp
