// einsum front end: lowers a pairwise tensor contraction to permutes + one FP64 DMMA GEMM.
//
// The reference hand-codes every contraction of src/ccsd.f90 as "omp_reshape, dgemm_wrapper, omp_reshape" or as a
// naive OpenMP loop nest.  Here each one is a single labelled statement; operands that are
// already laid out as a GEMM matrix (possibly transposed) are consumed in place, so e.g. the particle-particle
// ladder c(ij,ef)*v(ef,ab) runs with no data movement besides the GEMM itself.
#include <algorithm>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>

#include "kernels.cuh"
#include "tensor.cuh"

namespace afesp {

// ---------------------------------------------------------------- caching device allocator
namespace {
struct DeviceCache {
  std::multimap<size_t, double*> free_;   // size (doubles) -> block
  std::map<double*, size_t> live_;
  size_t cached_ = 0;                     // bytes sitting in free_
  void trim() {
    for (auto& kv : free_) cudaFree(kv.second);
    free_.clear();
    cached_ = 0;
  }
};
// Leaked singletons: buffers with static storage in other translation units (split-K workspace, pointer tables) are
// released during static destruction and must still find the cache alive.
std::mutex& g_cache_mu = *new std::mutex;
std::map<int, DeviceCache>& g_cache = *new std::map<int, DeviceCache>;   // one cache per device ordinal
DeviceCache& cache_here() {
  int dev = 0;
  cudaGetDevice(&dev);
  return g_cache[dev];
}
}  // namespace

double* device_alloc(size_t n) {
  if (n == 0) n = 1;
  std::lock_guard<std::mutex> lk(g_cache_mu);
  DeviceCache& c = cache_here();
  // reuse a cached block that is not much larger than the request (big blocks: <= 1.25x, small ones: <= 2x)
  auto it = c.free_.lower_bound(n);
  const size_t limit = n >= (8u << 20) ? n + n / 4 : 2 * n + 1024;
  if (it != c.free_.end() && it->first <= limit) {
    double* p = it->second;
    c.live_[p] = it->first;
    c.cached_ -= it->first * sizeof(double);
    c.free_.erase(it);
    return p;
  }
  double* p = nullptr;
  cudaError_t err = cudaMalloc(&p, n * sizeof(double));
  if (err != cudaSuccess) {
    cudaGetLastError();
    c.trim();
    AFESP_CUDA_CHECK(cudaMalloc(&p, n * sizeof(double)));
  }
  c.live_[p] = n;
  return p;
}

void device_free(double* p) {
  if (!p) return;
  std::lock_guard<std::mutex> lk(g_cache_mu);
  DeviceCache& c = cache_here();
  auto it = c.live_.find(p);
  if (it == c.live_.end()) {
    // allocated while another device was current: search the other caches
    for (auto& kv : g_cache) {
      auto jt = kv.second.live_.find(p);
      if (jt != kv.second.live_.end()) {
        kv.second.free_.emplace(jt->second, p);
        kv.second.cached_ += jt->second * sizeof(double);
        kv.second.live_.erase(jt);
        return;
      }
    }
    cudaFree(p);
    return;
  }
  c.free_.emplace(it->second, p);
  c.cached_ += it->second * sizeof(double);
  c.live_.erase(it);
}

void device_trim() {
  std::lock_guard<std::mutex> lk(g_cache_mu);
  cudaDeviceSynchronize();
  cache_here().trim();
}

size_t device_cached_bytes() {
  std::lock_guard<std::mutex> lk(g_cache_mu);
  return cache_here().cached_;
}

// ---------------------------------------------------------------- column-sharded GEMM over the ranks of one node
void Dist::exchange(double* base, const std::vector<std::pair<long long, long long>>& ranges, cudaStream_t st) {
  AFESP_REQUIRE(active() && (int)ranges.size() == nranks, "exchange: no communicator");
  AFESP_REQUIRE(group_start() == 0, "ncclGroupStart failed");
  for (int r = 0; r < nranks; ++r) {
    const long long cnt = ranges[r].second - ranges[r].first;
    if (cnt <= 0) continue;
    double* q = base + ranges[r].first;
    AFESP_REQUIRE(bcast(q, q, (size_t)cnt, r, comm, st) == 0, "ncclBroadcast failed");
    if (r != rank) exchanged_bytes += 8.0 * cnt;
  }
  AFESP_REQUIRE(group_end() == 0, "ncclGroupEnd failed");
}

void Dist::ensure_comm_stream() {
  if (comm_stream) return;
  AFESP_CUDA_CHECK(cudaStreamCreateWithFlags(&comm_stream, cudaStreamNonBlocking));
  for (auto& ev : ev_piece) AFESP_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  AFESP_CUDA_CHECK(cudaEventCreateWithFlags(&ev_done, cudaEventDisableTiming));
}

// C(M x N) = alpha op(A) op(B) + beta C with the output columns dealt to the ranks.  Every rank holds the same C on
// entry; it updates ITS column slab in place (beta included: no scratch copy, no second pass) and the updated slabs are
// then broadcast in place, so every rank leaves with the same C again.  The slab is computed in `overlap_chunks`
// pieces: the broadcasts of piece c (one grouped NCCL call on the communication stream, after an event) run over
// NVLink while the GEMM of piece c+1 runs on the compute stream; the compute stream waits for the last piece only.
void dgemm_sharded(Engine& e, char ta, char tb, int M, int N, int K, double alpha, const double* A, long long lda,
                   const double* B, long long ldb, double beta, double* C, bool b_local, bool force) {
  Dist& d = e.dist;
  const bool shard = d.active() && (force || (2.0 * M * (double)N * K >= d.min_flops && N >= 64LL * d.nranks));
  if (!shard) {
    AFESP_REQUIRE(!b_local, "dgemm_sharded: a local B slab needs the sharded path (communicator detached or disabled)");
    dgemm(e.stream, ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, M);
    return;
  }
  const bool tB = (tb == 'T' || tb == 't');
  int nch = std::max(1, std::min(d.overlap_chunks, 8));
  if (nch == 1 && d.allgather_on() && d.allgather) {
    // Equal slabs of `per` columns (the col_range unit, padded past N on the last ranks) in a scratch matrix
    // G(M x per*nranks): every rank computes its slab straight into its place (beta: its columns of C are copied in
    // first), ONE in-place ncclAllGather moves all slabs, and the N valid columns are copied back into C.
    const long long units = (N + 63) / 64, per = (units + d.nranks - 1) / d.nranks * 64;
    long long c0, c1;
    d.col_range(N, d.rank, &c0, &c1);
    Scratch G(e.pool, (size_t)M * per * d.nranks);
    double* mine = G.p + (size_t)M * per * d.rank;
    if (c1 > c0) {
      if (beta != 0.0)
        AFESP_CUDA_CHECK(cudaMemcpyAsync(mine, C + c0 * M, (size_t)(c1 - c0) * M * 8, cudaMemcpyDeviceToDevice, e.stream));
      const double* Bs = b_local ? B : (tB ? B + c0 : B + c0 * ldb);
      dgemm(e.stream, ta, tb, M, (int)(c1 - c0), K, alpha, A, lda, Bs, ldb, beta, mine, M);
    }
    AFESP_REQUIRE(d.allgather(mine, G.p, (size_t)M * per, d.comm, e.stream) == 0, "ncclAllGather failed");
    d.exchanged_bytes += 8.0 * M * per * (d.nranks - 1);
    AFESP_CUDA_CHECK(cudaMemcpyAsync(C, G.p, (size_t)M * N * 8, cudaMemcpyDeviceToDevice, e.stream));
    return;
  }
  if (nch > 1) d.ensure_comm_stream();
  // piece c of rank r: columns [lo, hi) of r's slab, split in multiples of 64 columns (the same rule on every rank)
  auto piece = [&](int r, int c, long long* lo, long long* hi) {
    long long a, b;
    d.col_range(N, r, &a, &b);
    const long long units = (b - a + 63) / 64, per = (units + nch - 1) / nch;
    *lo = std::min(b, a + per * c * 64);
    *hi = std::min(b, a + per * (c + 1) * 64);
  };
  long long c0, c1;
  d.col_range(N, d.rank, &c0, &c1);
  for (int c = 0; c < nch; ++c) {
    long long lo, hi;
    piece(d.rank, c, &lo, &hi);
    if (hi > lo) {
      const double* Bs = b_local ? (tB ? B + (lo - c0) : B + (lo - c0) * ldb) : (tB ? B + lo : B + lo * ldb);
      dgemm(e.stream, ta, tb, M, (int)(hi - lo), K, alpha, A, lda, Bs, ldb, beta, C + lo * M, M);
    }
    std::vector<std::pair<long long, long long>> ranges(d.nranks);
    for (int r = 0; r < d.nranks; ++r) {
      long long a, b;
      piece(r, c, &a, &b);
      ranges[r] = {a * M, b * M};
    }
    if (nch == 1) {
      d.exchange(C, ranges, e.stream);
    } else {
      AFESP_CUDA_CHECK(cudaEventRecord(d.ev_piece[c], e.stream));
      AFESP_CUDA_CHECK(cudaStreamWaitEvent(d.comm_stream, d.ev_piece[c], 0));
      d.exchange(C, ranges, d.comm_stream);
    }
  }
  if (nch > 1) {
    AFESP_CUDA_CHECK(cudaEventRecord(d.ev_done, d.comm_stream));
    AFESP_CUDA_CHECK(cudaStreamWaitEvent(e.stream, d.ev_done, 0));
  }
}

// ---------------------------------------------------------------- helpers
namespace {

std::string filter(const std::string& s, const std::string& set) {
  std::string r;
  for (char c : s) if (set.find(c) != std::string::npos) r.push_back(c);
  return r;
}

void split_spec(const char* spec, std::vector<std::string>& ins, std::string& out) {
  std::string s(spec);
  s.erase(std::remove(s.begin(), s.end(), ' '), s.end());
  size_t arrow = s.find("->");
  AFESP_REQUIRE(arrow != std::string::npos, "einsum spec needs '->'");
  out = s.substr(arrow + 2);
  std::string lhs = s.substr(0, arrow);
  size_t pos = 0;
  while (true) {
    size_t c = lhs.find(',', pos);
    ins.push_back(lhs.substr(pos, c == std::string::npos ? std::string::npos : c - pos));
    if (c == std::string::npos) break;
    pos = c + 1;
  }
}

void do_permute(Engine& e, const std::string& from, const std::string& to, const std::vector<int>& dims_from,
                double alpha, const double* in, double beta, double* out) {
  int rank = (int)from.size();
  AFESP_REQUIRE(to.size() == from.size(), "transpose: label count mismatch");
  if (rank == 0) {  // scalar
    int one = 1, zero = 0;
    permute(e.stream, 1, &one, &zero, alpha, in, beta, out);
    return;
  }
  int perm[6], dims[6];
  AFESP_REQUIRE(rank <= 6, "transpose: rank > 6");
  for (int d = 0; d < rank; ++d) {
    size_t k = from.find(to[d]);
    AFESP_REQUIRE(k != std::string::npos, "transpose: unknown output label");
    perm[d] = (int)k;
    dims[d] = dims_from[d];
  }
  permute(e.stream, rank, dims, perm, alpha, in, beta, out);
}

}  // namespace

void transpose(Engine& e, const char* spec, double alpha, const TView& in, double beta, const TView& out) {
  std::vector<std::string> ins;
  std::string so;
  split_spec(spec, ins, so);
  AFESP_REQUIRE(ins.size() == 1, "transpose: exactly one input");
  AFESP_REQUIRE(ins[0].size() == in.dims.size() && so.size() == out.dims.size(), "transpose: rank mismatch");
  for (size_t d = 0; d < so.size(); ++d) {
    size_t k = ins[0].find(so[d]);
    AFESP_REQUIRE(k != std::string::npos && in.dims[k] == out.dims[d], "transpose: extent mismatch");
  }
  do_permute(e, ins[0], so, in.dims, alpha, in.p, beta, out.p);
}

void einsum(Engine& e, const char* spec, double alpha, const TView& A_, const TView& B_, double beta,
            const TView& C) {
  std::vector<std::string> ins;
  std::string sc;
  split_spec(spec, ins, sc);
  AFESP_REQUIRE(ins.size() == 2, "einsum: exactly two inputs");
  std::string sa = ins[0], sb = ins[1];
  TView A = A_, B = B_;
  AFESP_REQUIRE(sa.size() == A.dims.size() && sb.size() == B.dims.size() && sc.size() == C.dims.size(),
                std::string("einsum rank mismatch in ") + spec);
  int ext[256];
  std::memset(ext, 0, sizeof(ext));
  auto reg = [&](const std::string& s, const std::vector<int>& d) {
    for (size_t i = 0; i < s.size(); ++i) {
      unsigned char c = (unsigned char)s[i];
      AFESP_REQUIRE(ext[c] == 0 || ext[c] == d[i], std::string("einsum extent mismatch in ") + spec);
      ext[c] = d[i];
    }
  };
  reg(sa, A.dims); reg(sb, B.dims); reg(sc, C.dims);
  std::string I, J, K;
  for (char c : sa) {
    bool inb = sb.find(c) != std::string::npos, inc = sc.find(c) != std::string::npos;
    AFESP_REQUIRE(inb != inc, std::string("einsum: label must be in exactly two tensors: ") + spec);
    (inc ? I : K).push_back(c);
  }
  for (char c : sb) {
    bool ina = sa.find(c) != std::string::npos, inc = sc.find(c) != std::string::npos;
    AFESP_REQUIRE(ina != inc, std::string("einsum: label must be in exactly two tensors: ") + spec);
    if (inc) J.push_back(c);
  }
  AFESP_REQUIRE(I.size() + J.size() == sc.size(), std::string("einsum: output labels unmatched: ") + spec);

  std::string cI = filter(sc, I), cJ = filter(sc, J);
  bool direct = (sc == cI + cJ);
  if (!direct && sc == cJ + cI) {  // C is (J,I): solve the transposed problem C^T = B^T A^T
    std::swap(A, B); std::swap(sa, sb); std::swap(I, J); std::swap(cI, cJ);
    direct = true;
  }
  std::string Iord = direct ? cI : filter(sa, I);
  std::string Jord = direct ? cJ : filter(sb, J);
  auto prod = [&](const std::string& s) { long long p = 1; for (char c : s) p *= ext[(unsigned char)c]; return p; };
  const long long M = prod(Iord), N = prod(Jord), Kd = prod(K);
  AFESP_REQUIRE(M < (1LL << 31) && N < (1LL << 31) && Kd < (1LL << 31), "einsum: matrix extent overflows int32");

  std::string kA = filter(sa, K), kB = filter(sb, K), Kord = kA;
  auto a_ok = [&](const std::string& ko) { return sa == Iord + ko || sa == ko + Iord; };
  auto b_ok = [&](const std::string& ko) { return sb == ko + Jord || sb == Jord + ko; };
  {
    long long costA = (a_ok(kA) ? 0 : A.size()) + (b_ok(kA) ? 0 : B.size());
    long long costB = (a_ok(kB) ? 0 : A.size()) + (b_ok(kB) ? 0 : B.size());
    if (costB < costA) Kord = kB;
  }

  // operand A as (I x K): 'N' when stored (I,K), 'T' when stored (K,I), otherwise permuted into (I,K)
  char ta = 'N';
  long long lda = M;
  const double* pa = A.p;
  std::unique_ptr<Scratch> sA, sB, sT;
  const bool a_nk = (sa == Iord + Kord), a_kn = (sa == Kord + Iord);
  if (a_kn && (M == 1 || !a_nk)) {
    ta = 'T'; lda = Kd;
  } else if (a_nk) {
    ta = 'N'; lda = M;
  } else {
    sA.reset(new Scratch(e.pool, (size_t)A.size()));
    do_permute(e, sa, Iord + Kord, A.dims, 1.0, A.p, 0.0, sA->p);
    pa = sA->p; ta = 'N'; lda = M;
  }
  // operand B as (K x J): 'N' when stored (K,J), 'T' when stored (J,K), otherwise permuted into (K,J)
  char tb = 'N';
  long long ldb = Kd;
  const double* pb = B.p;
  const bool b_kn = (sb == Kord + Jord), b_nk = (sb == Jord + Kord);
  if (b_kn && !(Kd == 1 && b_nk)) {
    tb = 'N'; ldb = Kd;
  } else if (b_nk) {
    tb = 'T'; ldb = N;
  } else {
    sB.reset(new Scratch(e.pool, (size_t)B.size()));
    do_permute(e, sb, Kord + Jord, B.dims, 1.0, B.p, 0.0, sB->p);
    pb = sB->p; tb = 'N'; ldb = Kd;
  }
  if (lda < 1) lda = 1;
  if (ldb < 1) ldb = 1;

  if (direct) {
    dgemm_sharded(e, ta, tb, (int)M, (int)N, (int)Kd, alpha, pa, lda, pb, ldb, beta, C.p);
  } else {
    sT.reset(new Scratch(e.pool, (size_t)(M * N)));
    dgemm_sharded(e, ta, tb, (int)M, (int)N, (int)Kd, alpha, pa, lda, pb, ldb, 0.0, sT->p);
    std::string st = Iord + Jord;
    std::vector<int> tdims;
    for (char c : st) tdims.push_back(ext[(unsigned char)c]);
    do_permute(e, st, sc, tdims, 1.0, sT->p, beta, C.p);
  }
}

}  // namespace afesp
