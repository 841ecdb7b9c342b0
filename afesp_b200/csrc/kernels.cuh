// Bandwidth-bound helper kernels of the CC path (element-wise builds, denominators, reductions, DIIS algebra).
#pragma once
#include "tensor.cuh"

namespace afesp {

// out(i,j,a,b) = x(i,j,a,b) / (eo_i + eo_j - ev_a - ev_b); out(i,a) = x(i,a)/(eo_i - ev_a).
// Denominators come from the orbital energies on the fly (the reference materialises D_ijab, src/ccsd.f90:431-457).
void divide_d2(cudaStream_t st, double* out, const double* x, const double* eo, const double* ev, int o, int v);
void divide_d2_probe(cudaStream_t st, double* out, const double* x, const double* eo, const double* ev, int o, int v);  // diagnostic
void divide_d1(cudaStream_t st, double* out, const double* x, const double* eo, const double* ev, int o, int v);

// out(i,j,a,b) = t2(i,j,a,b) + ca * t1(i,a) t1(j,b) + cb * t1(i,b) t1(j,a)
//   spin-free c_oovv: ca=1, cb=0 (src/ccsd.f90:1071-1079); tau: ca=1, cb=-1; tau_tilde: ca=.5, cb=-.5 (:701-711)
void t2_plus_t1t1(cudaStream_t st, double* out, const double* t2, const double* t1, int o, int v, double ca, double cb);

// out(b,a) = alpha * sum_m Z(m,b,m,a) for Z(o,v,o,v)   (diagonal of a two-step contraction, see ccsd_spatial.cu)
void diag_sum_nbma(cudaStream_t st, double* out, const double* Z, int o, int v, double alpha);

// y = a*x + b*y ; y = a*x (b == 0 never reads y)
void axpby(cudaStream_t st, long long n, double a, const double* x, double b, double* y);
void fill(cudaStream_t st, long long n, double val, double* y);

// Deterministic reductions.  `out` is a device array; results are complete when the stream reaches this point.
// dotn: out[k] = sum_i x_k[i] * y[i] for k < nx (nx <= 8) -- one pass over y for a whole DIIS row.
void dotn(Engine& e, long long n, int nx, const double* const* x_host_ptrs, const double* y, double* out);
// out[0] = sum (2 v(i,j,a,b) - v(i,j,b,a)) (t2 + t1 t1)(i,j,a,b), out[1] = sum (t2 - t2_old)^2   (src/ccsd.f90:1767-1786)
void cc_energy_restricted(Engine& e, const double* v_oovv, const double* t2, const double* t1, const double* t2_old,
                          int o, int v, double* out);
// out[0] = 1/4 sum oovv (t2 + 2 t1 t1), out[1] = sum (t2 - t2_old)^2                                (src/ccsd.f90:1787-1801)
void cc_energy_spinorb(Engine& e, const double* oovv, const double* t2, const double* t1, const double* t2_old, int o,
                       int v, double* out);
// y = sum_k c[k] * x_k  (k < nx <= 8), DIIS extrapolation (src/ccsd.f90:668-673)
void lincomb(cudaStream_t st, long long n, int nx, const double* const* x_host_ptrs, const double* c_host, double* y);

// Block-level helper shared by reduction kernels in other translation units.
double* reduce_scratch(Engine& e, size_t n);           // grow-only device scratch for per-block partials
void finish_partials(Engine& e, const double* partials, int nblocks, int nvals, double* out);  // out[v] = sum_b partials[b*nvals+v]

}  // namespace afesp
