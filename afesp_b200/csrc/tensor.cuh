// Device tensors, scratch pool and the einsum-style contraction front end used by the CC drivers.
#pragma once
#include <map>
#include <string>
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

// Size-bucketed cache of device scratch blocks (cudaMalloc/cudaFree synchronise; the CC loop reuses blocks).
class Pool {
 public:
  ~Pool() { clear(); }
  double* get(size_t n);
  void put(double* p);
  void clear();
  size_t bytes_held() const { return held_; }

 private:
  std::multimap<size_t, double*> free_;
  std::map<double*, size_t> live_;
  size_t held_ = 0;
};

struct Scratch {  // RAII lease from the pool
  Pool* pool;
  double* p;
  Scratch(Pool& pl, size_t n) : pool(&pl), p(pl.get(n)) {}
  ~Scratch() { if (p) pool->put(p); }
  Scratch(const Scratch&) = delete;
  Scratch& operator=(const Scratch&) = delete;
};

struct Engine {
  cudaStream_t stream = nullptr;
  Pool pool;
  DBuf red;  // per-block partial sums of the deterministic reductions (kernels.cu)
};

// C[ic] = alpha * sum_k A[ia] * B[ib] + beta * C[ic]; spec "ia,ib->ic" with one letter per axis.
// Every label appears in exactly two of the three tensors (no batch or trace labels).
// Lowered to (at most three) permutes + one DMMA GEMM ("TTGT"); operands already in GEMM order are used in place.
void einsum(Engine& e, const char* spec, double alpha, const TView& A, const TView& B, double beta, const TView& C);

// out[io] = alpha * in[ii] + beta * out[io]  (pure index permutation, labels as in einsum, spec "ii->io")
void transpose(Engine& e, const char* spec, double alpha, const TView& in, double beta, const TView& out);

}  // namespace afesp
