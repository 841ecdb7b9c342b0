// Bandwidth-bound helper kernels of the CC path.  All are grid-stride, coalesced along the fastest (first) index;
// reductions are two-stage and deterministic (warp shuffle -> block partials -> fixed-order final sum).
#include "kernels.cuh"

namespace afesp {
namespace {

constexpr int RB = 256;  // reduction block size

inline int grid_for(long long n, int block = 256) {
  long long b = (n + block - 1) / block;
  return (int)std::max<long long>(1, std::min<long long>(b, 148LL * 16));
}

// ---- division-free walk over an (o,o,v,v) array --------------------------------------------------------------------
// These kernels move 16-24 B per element, so the per-element index arithmetic decides whether they run at HBM speed:
// a 64-bit idx -> (i,j,a,b) decode (four integer divisions) costs several times the memory time.  Instead each block
// takes chunks of OOVV_CHUNK consecutive (a,b) pairs -- a contiguous run of CHUNK*o^2 doubles -- and keeps in shared
// memory (i, j) for every ij of the o^2 plane and (a, b) for every pair of the chunk, filled with a handful of divisions
// per block.  A thread then steps through the run with e += blockDim.x and updates (ij, pair) by compare-and-subtract.
constexpr int OOVV_T = 256, OOVV_CHUNK = 32;
constexpr int OOVV_MAX_OO = 3600;   // o <= 60 (tables stay under the 48 KB default dynamic shared memory); larger planes take the generic kernels below

struct OovvTables {
  unsigned short* i;   // [oo]
  unsigned short* j;   // [oo]
  unsigned short* a;   // [CHUNK]
  unsigned short* b;   // [CHUNK]
  double* eoo;         // [oo]     eo_i + eo_j   (optional)
  double* evv;         // [CHUNK]  ev_a + ev_b   (optional)
};

__host__ __device__ inline size_t oovv_smem_bytes(int oo, bool energies) {
  return (size_t)(energies ? (oo + OOVV_CHUNK) * sizeof(double) : 0) + (size_t)(2 * oo + 2 * OOVV_CHUNK) * sizeof(unsigned short);
}

template <bool ENERGIES, int UNROLL, class F>
__device__ __forceinline__ void oovv_walk(int o, int v, const double* __restrict__ eo, const double* __restrict__ ev,
                                          F&& f) {
  extern __shared__ __align__(16) unsigned char oovv_sm[];
  const int oo = o * o;
  const long long vv = (long long)v * v;
  OovvTables t;
  unsigned char* q = oovv_sm;
  t.eoo = reinterpret_cast<double*>(q); q += ENERGIES ? oo * sizeof(double) : 0;
  t.evv = reinterpret_cast<double*>(q); q += ENERGIES ? OOVV_CHUNK * sizeof(double) : 0;
  t.i = reinterpret_cast<unsigned short*>(q); q += oo * sizeof(unsigned short);
  t.j = reinterpret_cast<unsigned short*>(q); q += oo * sizeof(unsigned short);
  t.a = reinterpret_cast<unsigned short*>(q); q += OOVV_CHUNK * sizeof(unsigned short);
  t.b = reinterpret_cast<unsigned short*>(q);
  const int tid = threadIdx.x;
  for (int ij = tid; ij < oo; ij += OOVV_T) {
    const int jj = ij / o, ii = ij - jj * o;
    t.i[ij] = (unsigned short)ii; t.j[ij] = (unsigned short)jj;
    if (ENERGIES) t.eoo[ij] = eo[ii] + eo[jj];
  }
  const int ij0 = tid % oo, c0 = tid / oo;   // position of this thread's first element inside a chunk
  for (long long ab0 = (long long)blockIdx.x * OOVV_CHUNK; ab0 < vv; ab0 += (long long)gridDim.x * OOVV_CHUNK) {
    const int nab = (int)min((long long)OOVV_CHUNK, vv - ab0);
    __syncthreads();   // previous chunk fully consumed (and the ij tables written, first time round)
    if (tid < nab) {
      const long long ab = ab0 + tid;
      const int bb = (int)(ab / v), aa = (int)(ab - (long long)bb * v);
      t.a[tid] = (unsigned short)aa; t.b[tid] = (unsigned short)bb;
      if (ENERGIES) t.evv[tid] = ev[aa] + ev[bb];
    }
    __syncthreads();
    const long long base = ab0 * oo;
    const int n = nab * oo;
    int ij = ij0, c = c0;
    int e = tid;
    // four elements per trip: the positions come first so that the four independent loads can be in flight together
    if (UNROLL > 1) {
      for (; e + (UNROLL - 1) * OOVV_T < n; e += UNROLL * OOVV_T) {
        int ijs[UNROLL], cs[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          ijs[u] = ij; cs[u] = c;
          ij += OOVV_T;
          while (ij >= oo) { ij -= oo; ++c; }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) f(base + e + u * OOVV_T, ijs[u], cs[u], t);
      }
    }
    for (; e < n; e += OOVV_T) {
      f(base + e, ij, c, t);
      ij += OOVV_T;
      while (ij >= oo) { ij -= oo; ++c; }
    }
  }
}

inline int oovv_grid(int v) {
  const long long chunks = ((long long)v * v + OOVV_CHUNK - 1) / OOVV_CHUNK;
  return (int)std::max<long long>(1, std::min<long long>(chunks, 148LL * 8));
}

// Division.  The IEEE '/' costs this kernel a third of its bandwidth (63% of the copy peak against 93% with a
// multiplication in its place, k_divide_d2_probe): not the DFMA count but the slow-path check and call sequence around
// it, which serialises the four independent elements a thread has in flight.  fdiv() is branch-free instead: the
// hardware reciprocal seed (MUFU.RCP64H, relative error 2^-23), two Newton steps to full precision, and one residual
// correction of the quotient -- 1 MUFU + 7 DFMA, |error| <= 1 ulp (the reference itself is built with -ffast-math).
// Denominators of the CC equations are sums of orbital-energy differences: finite, far from the subnormal range.
__device__ __forceinline__ double fdiv(double x, double d) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  double e = fma(-d, y, 1.0);
  y = fma(y, e, y);
  e = fma(-d, y, 1.0);
  y = fma(y, e, y);
  double q = x * y;
  const double r = fma(-d, q, x);
  return fma(r, y, q);
}

__global__ void __launch_bounds__(OOVV_T) k_divide_d2_fast(double* __restrict__ out, const double* __restrict__ x,
                                                            const double* __restrict__ eo, const double* __restrict__ ev,
                                                            int o, int v) {
  oovv_walk<true, 4>(o, v, eo, ev, [&](long long idx, int ij, int c, const OovvTables& t) {
    out[idx] = fdiv(x[idx], t.eoo[ij] - t.evv[c]);
  });
}

// Diagnostic twin of k_divide_d2_fast (afesp_gpu_bench_hbm "divide_probe"): the same walk and traffic with the FP64
// division replaced by a multiplication, to separate the cost of the division from the cost of the walk.
__global__ void __launch_bounds__(OOVV_T) k_divide_d2_probe(double* __restrict__ out, const double* __restrict__ x,
                                                             const double* __restrict__ eo, const double* __restrict__ ev,
                                                             int o, int v) {
  oovv_walk<true, 4>(o, v, eo, ev, [&](long long idx, int ij, int c, const OovvTables& t) {
    out[idx] = x[idx] * (t.eoo[ij] - t.evv[c]);
  });
}

__global__ void __launch_bounds__(OOVV_T) k_t2_plus_t1t1_fast(double* __restrict__ out, const double* __restrict__ t2,
                                                               const double* __restrict__ t1, int o, int v, double ca,
                                                               double cb) {
  oovv_walk<false, 4>(o, v, nullptr, nullptr, [&](long long idx, int ij, int c, const OovvTables& t) {
    const int i = t.i[ij], j = t.j[ij], a = t.a[c], b = t.b[c];
    double r = t2[idx] + ca * t1[i + o * a] * t1[j + o * b];
    if (cb != 0.0) r += cb * t1[i + o * b] * t1[j + o * a];
    out[idx] = r;
  });
}

__global__ void k_divide_d2(double* __restrict__ out, const double* __restrict__ x, const double* __restrict__ eo,
                            const double* __restrict__ ev, int o, int v) {
  const long long oo = (long long)o * o, total = oo * v * v;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int ij = (int)(idx % oo);
    long long ab = idx / oo;
    int i = ij % o, j = ij / o, a = (int)(ab % v), b = (int)(ab / v);
    out[idx] = x[idx] / (eo[i] + eo[j] - ev[a] - ev[b]);
  }
}

__global__ void k_divide_d1(double* __restrict__ out, const double* __restrict__ x, const double* __restrict__ eo,
                            const double* __restrict__ ev, int o, int v) {
  int total = o * v;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x)
    out[idx] = x[idx] / (eo[idx % o] - ev[idx / o]);
}

__global__ void k_t2_plus_t1t1(double* __restrict__ out, const double* __restrict__ t2, const double* __restrict__ t1,
                               int o, int v, double ca, double cb) {
  const long long oo = (long long)o * o, total = oo * v * v;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int ij = (int)(idx % oo);
    long long ab = idx / oo;
    int i = ij % o, j = ij / o, a = (int)(ab % v), b = (int)(ab / v);
    double r = t2[idx] + ca * t1[i + o * a] * t1[j + o * b];
    if (cb != 0.0) r += cb * t1[i + o * b] * t1[j + o * a];
    out[idx] = r;
  }
}

__global__ void k_diag_sum_nbma(double* __restrict__ out, const double* __restrict__ Z, int o, int v, double alpha) {
  const int total = v * v;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    int b = idx % v, a = idx / v;
    double s = 0.0;
    for (int m = 0; m < o; ++m) s += Z[m + (long long)o * (b + (long long)v * (m + (long long)o * a))];
    out[idx] = alpha * s;
  }
}

__global__ void k_axpby(long long n, double a, const double* __restrict__ x, double b, double* __restrict__ y) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = (b == 0.0) ? a * x[i] : a * x[i] + b * y[i];
}

__global__ void k_fill(long long n, double val, double* __restrict__ y) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = val;
}

template <int NV>
__device__ __forceinline__ void block_reduce_store(double (&v)[NV], double* partials) {
  __shared__ double sh[NV][RB / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double x = v[k];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xffffffffu, x, off);
    if (lane == 0) sh[k][warp] = x;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0.0;
    for (int w = 0; w < RB / 32; ++w) s += sh[threadIdx.x][w];
    partials[(long long)blockIdx.x * NV + threadIdx.x] = s;
  }
}

struct PtrPack { const double* p[8]; double c[8]; };

template <int NX>
__global__ void __launch_bounds__(RB) k_dotn(long long n, PtrPack xs, const double* __restrict__ y,
                                              double* __restrict__ partials) {
  double acc[NX];
#pragma unroll
  for (int k = 0; k < NX; ++k) acc[k] = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double yv = y[i];
#pragma unroll
    for (int k = 0; k < NX; ++k) acc[k] += xs.p[k][i] * yv;
  }
  block_reduce_store<NX>(acc, partials);
}

// out[v] = sum_b partials[b*nvals + v]; one block per value, fixed summation order (deterministic)
__global__ void __launch_bounds__(RB) k_finish(const double* __restrict__ partials, int nblocks, int nvals,
                                                double* __restrict__ out) {
  __shared__ double sh[RB];
  const int v = blockIdx.x;
  double s = 0.0;
  for (int b = threadIdx.x; b < nblocks; b += RB) s += partials[(long long)b * nvals + v];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int w = RB / 2; w > 0; w >>= 1) {
    if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[v] = sh[0];
}

__global__ void __launch_bounds__(RB) k_energy_restricted(const double* __restrict__ vo, const double* __restrict__ t2,
                                                           const double* __restrict__ t1,
                                                           const double* __restrict__ t2_old, int o, int v,
                                                           double* __restrict__ partials) {
  const long long oo = (long long)o * o, total = oo * v * v;
  double acc[2] = {0.0, 0.0};
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int ij = (int)(idx % oo);
    long long ab = idx / oo;
    int i = ij % o, j = ij / o, a = (int)(ab % v), b = (int)(ab / v);
    double t = t2[idx];
    double vx = vo[ij + oo * (b + (long long)v * a)];
    acc[0] += (2.0 * vo[idx] - vx) * (t + t1[i + o * a] * t1[j + o * b]);
    double d = t - t2_old[idx];
    acc[1] += d * d;
  }
  block_reduce_store<2>(acc, partials);
}

__global__ void __launch_bounds__(RB) k_energy_spinorb(const double* __restrict__ vo, const double* __restrict__ t2,
                                                        const double* __restrict__ t1,
                                                        const double* __restrict__ t2_old, int o, int v,
                                                        double* __restrict__ partials) {
  const long long oo = (long long)o * o, total = oo * v * v;
  double acc[2] = {0.0, 0.0};
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int ij = (int)(idx % oo);
    long long ab = idx / oo;
    int i = ij % o, j = ij / o, a = (int)(ab % v), b = (int)(ab / v);
    double t = t2[idx];
    acc[0] += 0.25 * vo[idx] * (t + 2.0 * t1[i + o * a] * t1[j + o * b]);
    double d = t - t2_old[idx];
    acc[1] += d * d;
  }
  block_reduce_store<2>(acc, partials);
}

// Energy + amplitude-change norm in one pass over the (o,o,v,v) arrays, division-free walk (see oovv_walk).
template <bool SPINORB>
__global__ void __launch_bounds__(OOVV_T) k_energy_fast(const double* __restrict__ vo, const double* __restrict__ t2,
                                                         const double* __restrict__ t1, const double* __restrict__ t2_old,
                                                         int o, int v, double* __restrict__ partials) {
  double acc[2] = {0.0, 0.0};
  const long long oo = (long long)o * o;
  oovv_walk<false, 1>(o, v, nullptr, nullptr, [&](long long idx, int ij, int c, const OovvTables& t) {
    const int i = t.i[ij], j = t.j[ij], a = t.a[c], b = t.b[c];
    const double tv = t2[idx];
    if (SPINORB) {
      acc[0] += 0.25 * vo[idx] * (tv + 2.0 * t1[i + o * a] * t1[j + o * b]);
    } else {
      // <ij|ba> = <ji|ab>: the exchange partner sits in the SAME (a,b) plane at the transposed (i,j) position -- a second
      // read of the contiguous run this block is streaming (L1/L2 hit) instead of a gather from the (b,a) plane, which
      // doubled the DRAM traffic of v_oovv (the two elements are copies of one packed integral: bit-identical)
      const double vx = vo[idx - ij + (j + o * i)];
      acc[0] += (2.0 * vo[idx] - vx) * (tv + t1[i + o * a] * t1[j + o * b]);
    }
    const double d = tv - t2_old[idx];
    acc[1] += d * d;
  });
  __syncthreads();
  block_reduce_store<2>(acc, partials);
}

template <int NX>
__global__ void k_lincomb(long long n, PtrPack xs, double* __restrict__ y) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < NX; ++k) s += xs.c[k] * xs.p[k][i];
    y[i] = s;
  }
}

}  // namespace

void divide_d2(cudaStream_t st, double* out, const double* x, const double* eo, const double* ev, int o, int v) {
  if (o * o <= OOVV_MAX_OO && v < 65536) {
    k_divide_d2_fast<<<oovv_grid(v), OOVV_T, oovv_smem_bytes(o * o, true), st>>>(out, x, eo, ev, o, v);
    count_launch();
    return;
  }
  k_divide_d2<<<grid_for((long long)o * o * v * v), 256, 0, st>>>(out, x, eo, ev, o, v);
  count_launch();
}
void divide_d2_probe(cudaStream_t st, double* out, const double* x, const double* eo, const double* ev, int o, int v) {
  AFESP_REQUIRE(o * o <= OOVV_MAX_OO && v < 65536, "divide probe: shape outside the fast path");
  k_divide_d2_probe<<<oovv_grid(v), OOVV_T, oovv_smem_bytes(o * o, true), st>>>(out, x, eo, ev, o, v);
  count_launch();
}
void divide_d1(cudaStream_t st, double* out, const double* x, const double* eo, const double* ev, int o, int v) {
  k_divide_d1<<<grid_for((long long)o * v), 256, 0, st>>>(out, x, eo, ev, o, v);
  count_launch();
}
void t2_plus_t1t1(cudaStream_t st, double* out, const double* t2, const double* t1, int o, int v, double ca,
                  double cb) {
  if (o * o <= OOVV_MAX_OO && v < 65536) {
    k_t2_plus_t1t1_fast<<<oovv_grid(v), OOVV_T, oovv_smem_bytes(o * o, false), st>>>(out, t2, t1, o, v, ca, cb);
    count_launch();
    return;
  }
  k_t2_plus_t1t1<<<grid_for((long long)o * o * v * v), 256, 0, st>>>(out, t2, t1, o, v, ca, cb);
  count_launch();
}
void diag_sum_nbma(cudaStream_t st, double* out, const double* Z, int o, int v, double alpha) {
  k_diag_sum_nbma<<<grid_for((long long)v * v), 256, 0, st>>>(out, Z, o, v, alpha);
  count_launch();
}

void axpby(cudaStream_t st, long long n, double a, const double* x, double b, double* y) {
  if (n <= 0) return;
  k_axpby<<<grid_for(n), 256, 0, st>>>(n, a, x, b, y);
  count_launch();
}
void fill(cudaStream_t st, long long n, double val, double* y) {
  if (n <= 0) return;
  k_fill<<<grid_for(n), 256, 0, st>>>(n, val, y);
  count_launch();
}

double* reduce_scratch(Engine& e, size_t n) {
  if (e.red.n < n) e.red.alloc(std::max<size_t>(n, 1 << 16));
  return e.red.p;
}

void finish_partials(Engine& e, const double* partials, int nblocks, int nvals, double* out) {
  k_finish<<<nvals, RB, 0, e.stream>>>(partials, nblocks, nvals, out);
  count_launch();
  AFESP_CUDA_CHECK(cudaGetLastError());
}

void dotn(Engine& e, long long n, int nx, const double* const* xp, const double* y, double* out) {
  AFESP_REQUIRE(nx >= 1 && nx <= 8, "dotn: 1..8 vectors");
  PtrPack pk{};
  for (int k = 0; k < nx; ++k) pk.p[k] = xp[k];
  int nb = std::min(grid_for(n, RB), 1024);
  double* part = reduce_scratch(e, (size_t)nb * 8);
  switch (nx) {
#define C(NX) case NX: k_dotn<NX><<<nb, RB, 0, e.stream>>>(n, pk, y, part); break;
    C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8)
#undef C
  }
  count_launch();
  finish_partials(e, part, nb, nx, out);
}

void cc_energy_restricted(Engine& e, const double* v_oovv, const double* t2, const double* t1, const double* t2_old,
                          int o, int v, double* out) {
  int nb = std::min(grid_for((long long)o * o * v * v, RB), 1024);
  if (o * o <= OOVV_MAX_OO && v < 65536) nb = oovv_grid(v);
  double* part = reduce_scratch(e, (size_t)nb * 2);
  if (o * o <= OOVV_MAX_OO && v < 65536)
    k_energy_fast<false><<<nb, OOVV_T, oovv_smem_bytes(o * o, false), e.stream>>>(v_oovv, t2, t1, t2_old, o, v, part);
  else
    k_energy_restricted<<<nb, RB, 0, e.stream>>>(v_oovv, t2, t1, t2_old, o, v, part);
  count_launch();
  finish_partials(e, part, nb, 2, out);
}

void cc_energy_spinorb(Engine& e, const double* oovv, const double* t2, const double* t1, const double* t2_old, int o,
                       int v, double* out) {
  int nb = std::min(grid_for((long long)o * o * v * v, RB), 1024);
  if (o * o <= OOVV_MAX_OO && v < 65536) nb = oovv_grid(v);
  double* part = reduce_scratch(e, (size_t)nb * 2);
  if (o * o <= OOVV_MAX_OO && v < 65536)
    k_energy_fast<true><<<nb, OOVV_T, oovv_smem_bytes(o * o, false), e.stream>>>(oovv, t2, t1, t2_old, o, v, part);
  else
    k_energy_spinorb<<<nb, RB, 0, e.stream>>>(oovv, t2, t1, t2_old, o, v, part);
  count_launch();
  finish_partials(e, part, nb, 2, out);
}

void lincomb(cudaStream_t st, long long n, int nx, const double* const* xp, const double* c, double* y) {
  AFESP_REQUIRE(nx >= 1 && nx <= 8, "lincomb: 1..8 vectors");
  PtrPack pk{};
  for (int k = 0; k < nx; ++k) { pk.p[k] = xp[k]; pk.c[k] = c[k]; }
  int nb = grid_for(n);
  switch (nx) {
#define C(NX) case NX: k_lincomb<NX><<<nb, 256, 0, st>>>(n, pk, y); break;
    C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8)
#undef C
  }
  count_launch();
}

}  // namespace afesp
