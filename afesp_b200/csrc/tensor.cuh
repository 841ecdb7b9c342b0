// Device tensors, scratch pool and the einsum-style contraction front end used by the CC drivers.
#pragma once
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "common.cuh"

namespace afesp {

// Column-major (first index fastest, as in the Fortran reference) dense FP64 tensor view; does not own memory.
struct TView {
  double* p = nullptr;
  std::vector<int> dims;
  TView() = default;
  TView(double* ptr, std::vector<int> d) : p(ptr), dims(std::move(d)) {}
  long long size() const {
    long long s = 1;
    for (int d : dims) s *= d;
    return s;
  }
};

// Owning tensor.
struct Tensor {
  DBuf buf;
  std::vector<int> dims;
  Tensor() = default;
  explicit Tensor(std::vector<int> d) { init(std::move(d)); }
  void init(std::vector<int> d) {
    dims = std::move(d);
    long long s = 1;
    for (int x : dims) s *= x;
    buf.alloc((size_t)s);
  }
  void free() { buf.release(); dims.clear(); }
  long long size() const { return (long long)buf.n; }
  double* p() const { return buf.p; }
  TView view() const { return TView(buf.p, dims); }
  operator TView() const { return view(); }
};

// Scratch leases of the engine: a thin front over the caching device allocator (common.cuh).
class Pool {
 public:
  double* get(size_t n) { return device_alloc(n == 0 ? 1 : n); }
  void put(double* p) { device_free(p); }
  void clear() { device_trim(); }   // give the cached free blocks back to the driver (phase boundaries of large runs)
  size_t bytes_held() const { return device_cached_bytes(); }
};

struct Scratch {  // RAII lease from the pool
  Pool* pool;
  double* p;
  Scratch(Pool& pl, size_t n) : pool(&pl), p(pl.get(n)) {}
  ~Scratch() { if (p) pool->put(p); }
  Scratch(const Scratch&) = delete;
  Scratch& operator=(const Scratch&) = delete;
};

// Multi-GPU context of one rank (one process per GPU).  The CCSD state is replicated; the heavy GEMMs are sharded over
// output columns and the computed slabs are exchanged over NVLink with NCCL (installed by afesp_gpu_comm_init; the
// function pointers keep NCCL behind dlopen in capi.cu).  Every rank ends each exchange with bit-identical data, so the
// replicated state never diverges and the host-side convergence test gives the same answer on every rank.
struct Dist {
  int rank = 0, nranks = 1;
  void* comm = nullptr;
  int (*group_start)() = nullptr;
  int (*group_end)() = nullptr;
  int (*bcast)(const void* send, void* recv, size_t count, int root, void* comm, cudaStream_t st) = nullptr;  // doubles
  int (*allgather)(const void* send, void* recv, size_t count_per_rank, void* comm, cudaStream_t st) = nullptr;  // doubles
  int use_allgather = -1;      // sharded GEMMs: equal (padded) slabs + ONE in-place ncclAllGather instead of nranks grouped
                               // broadcasts.  -1 (default) = from 8 ranks on (measured: same speed at 2 ranks, 16.4 ms per
                               // CCSD iteration at 8 ranks, nbf=200), 0 / 1 = never / always (option "dist_allgather")
  bool allgather_on() const { return use_allgather < 0 ? nranks >= 8 : use_allgather != 0; }
  int (*send)(const void* buf, size_t count, int peer, void* comm, cudaStream_t st) = nullptr;
  int (*recv)(void* buf, size_t count, int peer, void* comm, cudaStream_t st) = nullptr;
  double min_flops = 4e9;   // GEMMs below this stay replicated (exchange latency would dominate)
  int overlap_chunks = 1;   // > 1: a rank's column slab is computed in this many pieces and the exchange of piece c runs on
                            // the communication stream while piece c+1 is computed (option "dist_overlap_chunks").  Measured
                            // on 2 x B200 at nbf=200 (profiles/r02_bench_n200_2gpu_overlap_chunks.json): 27.4 / 29.2 / 33.8 ms per
                            // CCSD iteration with 1 / 2 / 4 pieces -- the smaller GEMMs and the NCCL kernels competing for SMs
                            // cost more than the hidden transfer saves, so the default stays serial
  cudaStream_t comm_stream = nullptr;            // created on first use
  cudaEvent_t ev_piece[8] = {nullptr}, ev_done = nullptr;
  double exchanged_bytes = 0.0;  // bytes this rank received through slab exchanges (bench accounting)
  bool enabled = true;           // option dist_ccsd: 0 keeps CCSD / AO->MO replicated (only (T) is partitioned)
  bool active() const { return enabled && nranks > 1 && comm != nullptr; }
  // contiguous column range [c0, c1) of `ncols` owned by rank r, in multiples of `gran` columns
  void col_range(long long ncols, int r, long long* c0, long long* c1, long long gran = 64) const {
    const long long units = (ncols + gran - 1) / gran;
    const long long per = (units + nranks - 1) / nranks;
    *c0 = std::min(ncols, per * r * gran);
    *c1 = std::min(ncols, per * (r + 1) * gran);
  }
  // every rank broadcasts the element range it owns of a replicated array: ranges[r] = [begin, end) in doubles
  void exchange(double* base, const std::vector<std::pair<long long, long long>>& ranges, cudaStream_t st);
  void ensure_comm_stream();
};

struct Engine {
  cudaStream_t stream = nullptr;
  Pool pool;
  DBuf red;  // per-block partial sums of the deterministic reductions (kernels.cu)
  Dist dist;
};

// Wall-clock stage tracer (environment AFESP_TRACE=1 -> stderr); off by default: no synchronisation, no output.
struct Trace {
  bool on;
  cudaStream_t st;
  std::chrono::steady_clock::time_point t0;
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  explicit Trace(cudaStream_t s) : on(std::getenv("AFESP_TRACE") != nullptr), st(s) { t0 = std::chrono::steady_clock::now(); }
  void lap(int k) {
    if (!on) return;
    cudaStreamSynchronize(st);
    auto t1 = std::chrono::steady_clock::now();
    acc[k] += std::chrono::duration<double, std::milli>(t1 - t0).count();
    t0 = t1;
  }
  void report(const char* names[], int n) {
    if (!on) return;
    for (int k = 0; k < n; ++k) std::fprintf(stderr, "[afesp trace] %-14s %10.3f ms\n", names[k], acc[k]);
  }
};

// dgemm with ldc == M whose output columns are sharded over the ranks of e.dist (falls back to a plain dgemm when the
// context is inactive or the problem is small).  `b_local`: B already holds only this rank's column slab (tb = 'N').
void dgemm_sharded(Engine& e, char ta, char tb, int M, int N, int K, double alpha, const double* A, long long lda,
                   const double* B, long long ldb, double beta, double* C, bool b_local = false, bool force = false);

// C[ic] = alpha * sum_k A[ia] * B[ib] + beta * C[ic]; spec "ia,ib->ic" with one letter per axis.
// Every label appears in exactly two of the three tensors (no batch or trace labels).
// Lowered to (at most three) permutes + one DMMA GEMM ("TTGT"); operands already in GEMM order are used in place.
void einsum(Engine& e, const char* spec, double alpha, const TView& A, const TView& B, double beta, const TView& C);

// out[io] = alpha * in[ii] + beta * out[io]  (pure index permutation, labels as in einsum, spec "ii->io")
void transpose(Engine& e, const char* spec, double alpha, const TView& in, double beta, const TView& out);

}  // namespace afesp
