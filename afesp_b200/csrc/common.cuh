// Shared declarations for the AFESP B200 coupled-cluster engine (device side, sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

namespace afesp {

// ---- error handling: every failure becomes a C++ exception, turned into a status code at the C-ABI ----
struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define AFESP_CUDA_CHECK(expr)                                                                    \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      throw ::afesp::Error(2, std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " + \
                                  __FILE__ + ":" + std::to_string(__LINE__));                     \
  } while (0)

#define AFESP_REQUIRE(cond, msg)                                                               \
  do {                                                                                         \
    if (!(cond))                                                                               \
      throw ::afesp::Error(1, std::string(msg) + " (" #cond ") at " + __FILE__ + ":" +         \
                                  std::to_string(__LINE__));                                   \
  } while (0)

// ---- launch accounting (bench.py reports gpu_launches from this counter) ----
extern long long g_launch_count;
inline void count_launch(int n = 1) { g_launch_count += n; }

// ---- caching device allocator (contract.cu) ----
// cudaMalloc / cudaFree synchronise the device and cost up to ~1 s per call pattern once tens of GB are mapped, so
// every device block (named tensors and scratch alike) comes from a per-device size-bucketed cache: a freed block is
// kept and handed to the next request of a similar size.  All work of a handle is issued on one stream, so reuse is
// stream-ordered.  device_trim() returns the cached free blocks to the driver (phase boundaries of large runs; it is
// also what an allocation failure does before retrying).
double* device_alloc(size_t count);
void device_free(double* p);
void device_trim();
size_t device_cached_bytes();

// ---- device buffer with ownership ----
struct DBuf {
  double* p = nullptr;
  size_t n = 0;
  DBuf() = default;
  explicit DBuf(size_t count) { alloc(count); }
  DBuf(const DBuf&) = delete;
  DBuf& operator=(const DBuf&) = delete;
  DBuf(DBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DBuf& operator=(DBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DBuf() { release(); }
  void alloc(size_t count) {
    release();
    n = count;
    if (count) p = device_alloc(count);
  }
  void release() {
    if (p) device_free(p);
    p = nullptr; n = 0;
  }
};

// ---- GEMM (gemm.cu): C = alpha*op(A)*op(B) + beta*C, column-major, FP64 DMMA tiles ----
struct GemmBatch {
  int count = 1;
  long long strideA = 0, strideB = 0, strideC = 0;          // strided batching
  const double* const* Aptr = nullptr;                      // or device pointer arrays (override strides)
  const double* const* Bptr = nullptr;
  double* const* Cptr = nullptr;
  bool ptr_aligned16 = false;  // caller guarantees every A/B pointer in the arrays is 16-byte aligned
  // Optional "gather" description of the pointer arrays (enables the TMA path): Aptr[g] == Abase + Aidx[g]*Ablock, etc.
  const double* Abase = nullptr;
  const double* Bbase = nullptr;
  long long Ablock = 0, Bblock = 0;     // elements between consecutive blocks
  long long Anblocks = 0, Bnblocks = 0; // number of blocks addressable through the index arrays
  const int* Aidx = nullptr;            // device arrays of `count` block indices
  const int* Bidx = nullptr;
  // Optional second K segment ("dual" batches, the (T) driver): C[g] = A[Aidx[g]] B[Bidx[g]] + A[Aidx2[g]] B2[Bidx2[g]],
  // both products of depth K, A blocks from the same base, B2 blocks with the geometry of B.  One launch of the
  // TMA-staged kernel walks both segments in its k loop; the cp.async fallback runs two launches (beta, then 1).
  const double* const* Aptr2 = nullptr;
  const double* const* Bptr2 = nullptr;
  const double* Bbase2 = nullptr;
  const int* Aidx2 = nullptr;
  const int* Bidx2 = nullptr;
  bool dual() const { return Aptr2 != nullptr; }
};
void dgemm(cudaStream_t st, char transA, char transB, int M, int N, int K, double alpha, const double* A,
           long long lda, const double* B, long long ldb, double beta, double* C, long long ldc,
           const GemmBatch* batch = nullptr);
// Executed DMMA flop counter (2*M*N*K per gemm), for "% of FP64 tensor peak" from executed flops.
extern double g_gemm_flops;
// Optional per-launch device timing of the GEMM kernel (CUDA events on the launching stream), used by bench.py for
// the live roofline figure.  gemm_timing_collect() synchronises and returns the accumulated milliseconds.
void gemm_timing_enable(bool on);
void gemm_force_config(int cfg);  // tuning aid: -1 = automatic tile selection
int gemm_force_config_get();
// TMA-staged operand path: scope 0 = off, 1 = gathered batches only (the (T) contraction; default), 2 = every aligned
// problem the 64x64 tile is chosen for.  See DESIGN.md section 4.1.
void gemm_tma_scope(int scope);
int gemm_tma_scope_get();
void gemm_tma_edge(int on);   // 1 (default): balanced last M tile + short K tail in the TMA kernel; 0: pad (A/B tests)
// One-off consistency check of the TMA kernel against the cp.async kernel on the current device (gemm_tma.cu); a
// failure switches the TMA path off for the process.  State: 0 not run, 1 passed, -1 failed.
bool gemm_tma_selftest(cudaStream_t st);
int gemm_tma_selftest_state();
void gemm_crosscheck(cudaStream_t st, char ta, char tb, int M, int N, int K, int nbatch, double beta, int reps,
                     unsigned long long* bad, double* ms_tma, double* ms_ref);
bool dgemm_tma(cudaStream_t st, bool ak, bool bk, int M, int N, int K, double alpha, const double* A, long long lda,
               const double* B, long long ldb, double beta, double* C, long long ldc, const GemmBatch* batch, int cvec);
double gemm_timing_collect(double* flops_out, long long* launches_out = nullptr);

// ---- permute (permute.cu): out = alpha * permute(in) + beta * out for rank <= 6 ----
// perm[d] = which input axis becomes output axis d (numpy transpose convention), dims = input extents,
// column-major (axis 0 fastest) on both sides.
void permute(cudaStream_t st, int rank, const int* dims, const int* perm, double alpha, const double* in,
             double beta, double* out);
void permute_strided(cudaStream_t st, int rank, const int* dims, const int* perm, double alpha, const double* in,
                     double beta, double* out, const long long* ostride);

}  // namespace afesp
