// FP64 tensor-core GEMM for sm_100a:  C = alpha * op(A) * op(B) + beta * C   (column-major, BLAS semantics)
//
// Replaces the reference's dgemm_wrapper -> OpenBLAS dgemm (src/linalg.fpp:58-89) for every dense contraction
// of the coupled-cluster path (call-site list: SURVEY.md §2.3).
//
// Hardware mapping.  On Blackwell the FP64 tensor path is the warp-level DMMA: ptxas lowers every
// mma.sync .f64 shape (m16n8k4/8/16 included) to DMMA.8x8x4 on sm_100a, so the kernel issues the native
// m8n8k4 shape directly.  tcgen05/TMEM have no f64 kind.  One DMMA.8x8x4 = 256 FMA per warp instruction.
//
// Tiling.  A CTA owns a BM x BN tile of C; its warps own WM x WN sub-tiles made of 8x8 DMMA accumulators
// kept in registers.  Operand tiles (BM x BK and BK x BN doubles) are staged global -> shared by an
// asynchronous-copy ring of STAGES slots so that loads of tile k+STAGES-1 overlap the DMMAs of tile k.
// Shared tiles keep the operand's own contiguous direction ("MN-major" for A^N / B^T, "K-major" for A^T / B^N)
// and are padded so that the row stride is 4 (mod 16) 8-byte words: a DMMA fragment load by a half-warp
// (4 groups x 4 lanes) then touches 16 distinct banks -- conflict-free for both orientations.
//
// Skinny problems (M*N tile count below the SM count with a long K) are split along K into a workspace and
// reduced by a second kernel, deterministically (no atomics).
#include <cmath>

#include <cmath>
#include <cstdlib>

#include "common.cuh"

namespace afesp {

long long g_launch_count = 0;
double g_gemm_flops = 0.0;

namespace {

constexpr int PAD = 4;  // doubles of row padding -> row stride = 4 (mod 16) words

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// Shared-memory operand tile: logical (mn, k), mn in [0,BMN), k in [0,BK).
template <int BMN, int BK, bool KMAJOR>
struct Tile {
  static constexpr int LD = KMAJOR ? (BK + PAD) : (BMN + PAD);
  static constexpr int SIZE = KMAJOR ? (BMN * LD) : (BK * LD);
  __device__ static __forceinline__ int off(int mn, int k) { return KMAJOR ? (mn * LD + k) : (k * LD + mn); }
};

// Asynchronous copy of one operand tile.  X(mn,k) lives at X[mn*1 + k*ld] (MN-major) or X[k*1 + mn*ld]
// (K-major).  Out-of-range elements are zero-filled by the copy unit (src-size 0).  VEC=2 moves 16-byte
// chunks and requires a 16-byte aligned base and an even ld; VEC=1 is the any-alignment path.
template <int BMN, int BK, bool KMAJOR, int VEC, int NT>
__device__ __forceinline__ void load_tile(double* smem, const double* __restrict__ X, long long ld, int mn0, int k0,
                                          int MN, int Kend, int tid) {
  using T = Tile<BMN, BK, KMAJOR>;
  constexpr int ROWLEN = KMAJOR ? BK : BMN;       // contiguous extent of one tile row
  constexpr int NROWS = KMAJOR ? BMN : BK;
  constexpr int CPR = ROWLEN / VEC;               // chunks per row
  constexpr int TOTAL = CPR * NROWS;
  constexpr int ITERS = (TOTAL + NT - 1) / NT;
#pragma unroll
  for (int i = 0; i < ITERS; ++i) {
    int c = tid + i * NT;
    if (TOTAL % NT != 0 && c >= TOTAL) break;
    int r = c / CPR;
    int cc = (c % CPR) * VEC;
    int mn = KMAJOR ? (mn0 + r) : (mn0 + cc);
    int k = KMAJOR ? (k0 + cc) : (k0 + r);
    int lim_c = KMAJOR ? (Kend - k) : (MN - mn);    // valid elements along the contiguous direction
    bool row_ok = KMAJOR ? (mn < MN) : (k < Kend);
    int valid = row_ok ? (lim_c < 0 ? 0 : (lim_c > VEC ? VEC : lim_c)) : 0;
    const double* src = valid ? (KMAJOR ? (X + (long long)mn * ld + k) : (X + (long long)k * ld + mn)) : X;
    double* dst = smem + r * T::LD + cc;
    if (VEC == 2) cp_async16(dst, src, valid * 8);
    else cp_async8(dst, src, valid * 8);
  }
}

struct Params {
  const double* A;
  const double* B;
  double* C;
  int M, N, K;
  long long lda, ldb, ldc;
  double alpha, beta;
  long long sA, sB, sC;  // batch strides
  const double* const* Ap;
  const double* const* Bp;
  double* const* Cp;
  int zfast;             // > 0: batch-fastest rasterisation on a 1-D grid (value = batch count), see the kernel
  int splitk;            // >1: write raw partials to ws[split][M*N]
  int kchunk;            // K extent per split (multiple of BK)
  double* ws;
  int cvec;              // 1: C base 16-byte aligned and ldc even -> 16-byte epilogue accesses
  int tiles_m;           // CTA tiles along M; blockIdx.x enumerates (m-tile fastest, n-tile)
};

template <int BM, int BN, int BK, int WM, int WN, int STAGES, int MINB, bool AK, bool BKM, int VEC>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32, MINB) gemm_f64_dmma(const Params p) {
  constexpr int NT = (BM / WM) * (BN / WN) * 32;
  using TA = Tile<BM, BK, AK>;
  using TB = Tile<BN, BK, BKM>;
  constexpr int STAGE_DOUBLES = TA::SIZE + TB::SIZE;
  extern __shared__ __align__(16) double smem[];

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int gid = lane >> 2, tig = lane & 3;
  const int wm0 = (warp % (BM / WM)) * WM;
  const int wn0 = (warp / (BM / WM)) * WN;
  // Rasterisation.  Pointer-array batches (the (T) driver sorts them by their B block) walk (m-tile, batch entry)
  // fastest and the n-tile slowest on a 1-D grid, so the CTAs sharing one K x BN tile of B run back to back and B comes
  // from DRAM once per distinct block; everything else uses (tile, batch/split) on grid.x / grid.z.
  int m0, n0, batch, split = 0;
  if (p.zfast > 0) {
    const unsigned per_n = (unsigned)p.tiles_m * (unsigned)p.zfast;
    const unsigned nt = blockIdx.x / per_n, rem = blockIdx.x - nt * per_n;
    batch = (int)(rem / (unsigned)p.tiles_m);
    m0 = (int)(rem - (unsigned)batch * p.tiles_m) * BM;
    n0 = (int)nt * BN;
  } else {
    m0 = (blockIdx.x % p.tiles_m) * BM; n0 = (blockIdx.x / p.tiles_m) * BN;
    batch = blockIdx.z;
    if (p.splitk > 1) { split = blockIdx.z % p.splitk; batch = blockIdx.z / p.splitk; }
  }
  const double* A = p.Ap ? p.Ap[batch] : p.A + batch * p.sA;
  const double* B = p.Bp ? p.Bp[batch] : p.B + batch * p.sB;
  double* C = p.Cp ? p.Cp[batch] : p.C + batch * p.sC;

  const int kbeg = split * p.kchunk;
  const int kend = (p.splitk > 1) ? min(p.K, kbeg + p.kchunk) : p.K;
  const int nk = (kend - kbeg + BK - 1) / BK;

  constexpr int MT = WM / 8, NTL = WN / 8;
  double acc[MT][NTL][2];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NTL; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  auto issue = [&](int kt) {
    if (kt < nk) {
      double* sa = smem + (kt % STAGES) * STAGE_DOUBLES;
      double* sb = sa + TA::SIZE;
      load_tile<BM, BK, AK, VEC, NT>(sa, A, p.lda, m0, kbeg + kt * BK, p.M, kend, tid);
      load_tile<BN, BK, BKM, VEC, NT>(sb, B, p.ldb, n0, kbeg + kt * BK, p.N, kend, tid);
    }
    cp_async_commit();
  };

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) issue(s);

  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    issue(kt + STAGES - 1);  // refills the slot consumed in iteration kt-1 (all warps are past it)
    const double* sa = smem + (kt % STAGES) * STAGE_DOUBLES;
    const double* sb = sa + TA::SIZE;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double af[MT], bf[NTL];
#pragma unroll
      for (int i = 0; i < MT; ++i) af[i] = sa[TA::off(wm0 + 8 * i + gid, kk + tig)];
#pragma unroll
      for (int j = 0; j < NTL; ++j) bf[j] = sb[TB::off(wn0 + 8 * j + gid, kk + tig)];
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NTL; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  cp_async_wait<0>();

  // Epilogue straight from the accumulator fragments: thread holds C(row = gid, cols = 2*tig, 2*tig+1).
  if (p.splitk > 1) {
    double* W = p.ws + ((long long)batch * p.splitk + split) * (long long)p.M * p.N;
#pragma unroll
    for (int i = 0; i < MT; ++i) {
      int m = m0 + wm0 + 8 * i + gid;
      if (m >= p.M) continue;
#pragma unroll
      for (int j = 0; j < NTL; ++j) {
        int n = n0 + wn0 + 8 * j + 2 * tig;
        if (n < p.N) W[m + (long long)n * p.M] = acc[i][j][0];
        if (n + 1 < p.N) W[m + (long long)(n + 1) * p.M] = acc[i][j][1];
      }
    }
    return;
  }
  const double alpha = p.alpha, beta = p.beta;
  if (p.cvec) {
    // 16-byte epilogue: lanes (gid, gid^1) swap one accumulator so that each thread owns two consecutive rows of one
    // column: even gid keeps column 2*tig for rows (gid, gid+1), odd gid keeps column 2*tig+1 for rows (gid-1, gid).
    const bool odd = gid & 1;
#pragma unroll
    for (int i = 0; i < MT; ++i) {
      const int m = m0 + wm0 + 8 * i + (gid & ~1);
      double2 oldv[NTL];
      if (beta != 0.0) {  // issue every read of this row block before the first dependent store (memory-level parallelism)
#pragma unroll
        for (int j = 0; j < NTL; ++j) {
          const int n = n0 + wn0 + 8 * j + 2 * tig + (odd ? 1 : 0);
          oldv[j] = make_double2(0.0, 0.0);
          if (n < p.N && m < p.M) {
            const double* c = C + m + (long long)n * p.ldc;
            if (m + 1 < p.M) oldv[j] = *reinterpret_cast<const double2*>(c);
            else oldv[j].x = *c;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < NTL; ++j) {
        const double send = odd ? acc[i][j][0] : acc[i][j][1];
        const double recv = __shfl_xor_sync(0xffffffffu, send, 4);
        const double lo = odd ? recv : acc[i][j][0];
        const double hi = odd ? acc[i][j][1] : recv;
        const int n = n0 + wn0 + 8 * j + 2 * tig + (odd ? 1 : 0);
        if (n < p.N && m < p.M) {
          double* c = C + m + (long long)n * p.ldc;
          double2 v = make_double2(alpha * lo, alpha * hi);
          if (beta != 0.0) { v.x += beta * oldv[j].x; v.y += beta * oldv[j].y; }
          if (m + 1 < p.M) *reinterpret_cast<double2*>(c) = v;
          else *c = v.x;
        }
      }
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < MT; ++i) {
    int m = m0 + wm0 + 8 * i + gid;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < NTL; ++j) {
      int n = n0 + wn0 + 8 * j + 2 * tig;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        if (n + e < p.N) {
          double* c = C + m + (long long)(n + e) * p.ldc;
          double v = alpha * acc[i][j][e];
          if (beta != 0.0) v += beta * (*c);
          *c = v;
        }
      }
    }
  }
}

// C = alpha * sum_s W[s] + beta * C  (split-K reduction; fixed summation order)
__global__ void splitk_reduce(const double* __restrict__ W, int splitk, int M, int N, double alpha, double beta,
                              double* __restrict__ C, long long ldc) {
  long long MN = (long long)M * N;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < MN;
       idx += (long long)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int k = 0; k < splitk; ++k) s += W[k * MN + idx];
    int m = (int)(idx % M);
    long long n = idx / M;
    double* c = C + m + n * ldc;
    double v = alpha * s;
    if (beta != 0.0) v += beta * (*c);
    *c = v;
  }
}

// C = beta * C for the degenerate K == 0 / alpha == 0 case
__global__ void scale_c(double* C, int M, int N, long long ldc, double beta) {
  long long MN = (long long)M * N;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < MN;
       idx += (long long)gridDim.x * blockDim.x) {
    double* c = C + (idx % M) + (idx / M) * ldc;
    *c = (beta == 0.0) ? 0.0 : beta * (*c);
  }
}

constexpr int kMaxDev = 64;
int g_num_sms[kMaxDev] = {0};   // per device ordinal
inline int current_device() {
  int dev = 0;
  AFESP_CUDA_CHECK(cudaGetDevice(&dev));
  AFESP_REQUIRE(dev >= 0 && dev < kMaxDev, "gemm: device ordinal out of range");
  return dev;
}

// per-launch event timing (off by default)
bool g_timing = false;
std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_ev_pool;
size_t g_ev_used = 0;
double g_timed_ms = 0.0, g_timed_flops = 0.0;
long long g_timed_launches = 0;
double drain_events() {
  double ms = 0.0;
  for (size_t i = 0; i < g_ev_used; ++i) {
    cudaEventSynchronize(g_ev_pool[i].second);
    float t = 0.f;
    cudaEventElapsedTime(&t, g_ev_pool[i].first, g_ev_pool[i].second);
    ms += t;
  }
  g_ev_used = 0;
  return ms;
}
std::pair<cudaEvent_t, cudaEvent_t>* next_events() {
  if (g_ev_used == 2048) g_timed_ms += drain_events();
  if (g_ev_used == g_ev_pool.size()) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    g_ev_pool.emplace_back(a, b);
  }
  return &g_ev_pool[g_ev_used++];
}

int num_sms() {
  const int dev = current_device();
  if (!g_num_sms[dev]) AFESP_CUDA_CHECK(cudaDeviceGetAttribute(&g_num_sms[dev], cudaDevAttrMultiProcessorCount, dev));
  return g_num_sms[dev];
}

template <int BM, int BN, int BK, int WM, int WN, int STAGES, int MINB, bool AK, bool BKM, int VEC>
void launch_cfg(cudaStream_t st, const Params& p, int nbatch) {
  constexpr int NT = (BM / WM) * (BN / WN) * 32;
  constexpr size_t SMEM = (size_t)STAGES * (Tile<BM, BK, AK>::SIZE + Tile<BN, BK, BKM>::SIZE) * sizeof(double);
  auto kern = gemm_f64_dmma<BM, BN, BK, WM, WN, STAGES, MINB, AK, BKM, VEC>;
  static bool attr_set[kMaxDev] = {false};   // the attribute is per device
  const int dev = current_device();
  if (!attr_set[dev]) {
    AFESP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    attr_set[dev] = true;
  }
  Params q = p;
  q.tiles_m = (p.M + BM - 1) / BM;
  const long long tiles = (long long)q.tiles_m * ((p.N + BN - 1) / BN);
  AFESP_REQUIRE(tiles < (1LL << 31), "gemm: too many tiles");
  dim3 grid((unsigned)tiles, 1, nbatch * (p.splitk > 1 ? p.splitk : 1));
  q.zfast = 0;
  if (nbatch > 1 && p.splitk <= 1 && p.Bp != nullptr && tiles * nbatch < (1LL << 31)) {
    q.zfast = nbatch;
    grid = dim3((unsigned)(tiles * nbatch), 1, 1);
  }
  AFESP_REQUIRE(grid.z <= 65535, "gemm: batch too large");
  kern<<<grid, NT, SMEM, st>>>(q);
  count_launch();
  AFESP_CUDA_CHECK(cudaGetLastError());
}

template <int BM, int BN, int WM, int WN, int STAGES, int MINB>
void launch_tile(cudaStream_t st, const Params& p, int nbatch, bool ak, bool bk, bool vec2) {
  constexpr int BK = 16;
#define AFESP_GEMM_CASE(AKv, BKv, V) launch_cfg<BM, BN, BK, WM, WN, STAGES, MINB, AKv, BKv, V>(st, p, nbatch)
  if (vec2) {
    if (ak) { if (bk) AFESP_GEMM_CASE(true, true, 2); else AFESP_GEMM_CASE(true, false, 2); }
    else    { if (bk) AFESP_GEMM_CASE(false, true, 2); else AFESP_GEMM_CASE(false, false, 2); }
  } else {
    if (ak) { if (bk) AFESP_GEMM_CASE(true, true, 1); else AFESP_GEMM_CASE(true, false, 1); }
    else    { if (bk) AFESP_GEMM_CASE(false, true, 1); else AFESP_GEMM_CASE(false, false, 1); }
  }
#undef AFESP_GEMM_CASE
}

// Tile menu.  `eff` is the measured fraction of the DMMA issue peak the config reaches on a large square problem
// (B200, profiles/); `occ` its resident CTAs per SM.  The dispatcher maximises eff * (useful / padded work) *
// (wave fill) over the menu.
struct TileCfg { int bm, bn, occ; double eff; };
constexpr int NCFG = 7;
const TileCfg kCfg[NCFG] = {
    {128, 128, 1, 0.83},  // 0: 8 warps of 64x32
    {64, 128, 2, 0.85},   // 1: 8 warps of 32x32
    {128, 64, 2, 0.86},   // 2: 8 warps of 32x32
    {64, 64, 3, 0.87},    // 3: 4 warps of 32x32  (three resident CTAs hide the cp.async prologue and the epilogue)
    {32, 128, 3, 0.84},   // 4: 4 warps of 32x32 (skinny M)
    {128, 32, 3, 0.82},   // 5: 4 warps of 32x32 (skinny N)
    {96, 128, 1, 0.74},   // 6: 8 warps of 48x32
};
int g_force_cfg = -1;

void launch_by_cfg(int cfg, cudaStream_t st, const Params& p, int nbatch, bool ak, bool bk, bool vec2) {
  switch (cfg) {
    case 0: launch_tile<128, 128, 64, 32, 3, 1>(st, p, nbatch, ak, bk, vec2); break;
    case 1: launch_tile<64, 128, 32, 32, 3, 2>(st, p, nbatch, ak, bk, vec2); break;
    case 2: launch_tile<128, 64, 32, 32, 3, 2>(st, p, nbatch, ak, bk, vec2); break;
    case 3: launch_tile<64, 64, 32, 32, 3, 3>(st, p, nbatch, ak, bk, vec2); break;
    case 4: launch_tile<32, 128, 32, 32, 3, 3>(st, p, nbatch, ak, bk, vec2); break;
    case 5: launch_tile<128, 32, 32, 32, 3, 3>(st, p, nbatch, ak, bk, vec2); break;
    case 6: launch_tile<96, 128, 48, 32, 3, 1>(st, p, nbatch, ak, bk, vec2); break;
    default: throw Error(1, "gemm: bad tile config");
  }
}

int choose_cfg(int M, int N, int nbatch, long long* tiles_out, bool tma_ok) {
  int best = 0;
  double best_score = -1.0;
  const double sms = num_sms();
  for (int c = 0; c < NCFG; ++c) {
    const TileCfg& t = kCfg[c];
    const long long tm = (M + t.bm - 1) / t.bm, tn = (N + t.bn - 1) / t.bn;
    const long long tiles = tm * tn * nbatch;
    double useful = ((double)M * N) / ((double)tm * t.bm * tn * t.bn);
    // the TMA-staged kernel multiplies only ceil(M/8) row fragments (balanced last M tile, gemm_tma.cu)
    if (c == 3 && tma_ok) useful = ((double)M * N) / ((double)((M + 7) / 8 * 8) * tn * t.bn);
    const double slots = sms * t.occ;
    const double waves = std::ceil(tiles / slots);
    const double fill = tiles / (waves * slots);
    // the 64x64 tile runs through the TMA-staged kernel when the operands are aligned: 94-96% of the DMMA issue rate
    // on large problems instead of 87% (profiles/r02_tma_soak_*.json, r02_ncu_*)
    const double eff = (c == 3 && tma_ok) ? 0.95 : t.eff;
    const double score = eff * useful * fill;
    if (score > best_score) { best_score = score; best = c; if (tiles_out) *tiles_out = tiles; }
  }
  return best;
}

}  // namespace

void dgemm(cudaStream_t st, char transA, char transB, int M, int N, int K, double alpha, const double* A,
           long long lda, const double* B, long long ldb, double beta, double* C, long long ldc,
           const GemmBatch* batch) {
  AFESP_REQUIRE(M >= 0 && N >= 0 && K >= 0, "dgemm: negative dimension");
  if (M == 0 || N == 0) return;
  const int nbatch = batch ? batch->count : 1;
  if (nbatch == 0) return;
  const bool ta = (transA == 'T' || transA == 't'), tb = (transB == 'T' || transB == 't');
  AFESP_REQUIRE(ta || transA == 'N' || transA == 'n', "dgemm: transA must be N or T");
  AFESP_REQUIRE(tb || transB == 'N' || transB == 'n', "dgemm: transB must be N or T");
  if (K == 0 || alpha == 0.0) {
    AFESP_REQUIRE(nbatch == 1, "dgemm: K==0 batched not supported");
    scale_c<<<std::min<long long>(((long long)M * N + 255) / 256, 4096), 256, 0, st>>>(C, M, N, ldc, beta);
    count_launch();
    return;
  }
  Params p{};
  p.A = A; p.B = B; p.C = C; p.M = M; p.N = N; p.K = K;
  p.lda = lda; p.ldb = ldb; p.ldc = ldc; p.alpha = alpha; p.beta = beta;
  if (batch) {
    p.sA = batch->strideA; p.sB = batch->strideB; p.sC = batch->strideC;
    p.Ap = batch->Aptr; p.Bp = batch->Bptr; p.Cp = batch->Cptr;
  }
  p.splitk = 1;
  const bool ak = ta;    // op(A)(m,k) contiguous in k  <=> transA == 'T'
  const bool bk = !tb;   // op(B)(k,n) contiguous in k  <=> transB == 'N'
  auto aligned16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  bool vec2 = (lda % 2 == 0) && (ldb % 2 == 0);
  if (batch && (batch->Aptr || batch->Bptr)) vec2 = vec2 && batch->ptr_aligned16;  // alignment is the caller's promise
  else vec2 = vec2 && aligned16(A) && aligned16(B) && (!batch || (batch->strideA % 2 == 0 && batch->strideB % 2 == 0));

  const bool dual = batch && batch->dual();   // two K segments per batch entry (common.cuh)
  const double flops = 2.0 * M * N * (double)K * nbatch * (dual ? 2 : 1);
  g_gemm_flops += flops;

  // tile selection (see kCfg)
  long long tiles = 0;
  const bool gathered = batch && batch->Abase && batch->Bbase;
  const bool tma_ok = vec2 && (gemm_tma_scope_get() == 2 ? (!batch || !(batch->Aptr || batch->Bptr) || gathered)
                                                         : (gemm_tma_scope_get() == 1 && gathered));
  int cfg = choose_cfg(M, N, nbatch, &tiles, tma_ok);
  if (g_force_cfg >= 0 && g_force_cfg < NCFG) {
    cfg = g_force_cfg;
    tiles = (long long)((M + kCfg[cfg].bm - 1) / kCfg[cfg].bm) * ((N + kCfg[cfg].bn - 1) / kCfg[cfg].bn) * nbatch;
  }
  {
    const bool c_al = batch && batch->Cptr ? batch->ptr_aligned16 : (aligned16(C) && (!batch || batch->strideC % 2 == 0));
    p.cvec = (ldc % 2 == 0) && c_al ? 1 : 0;
  }
  // split-K for skinny outputs with a long reduction
  if (dual) AFESP_REQUIRE(batch->Aptr && batch->Bptr && batch->Bptr2 && batch->Cptr, "dgemm: dual batches need pointer arrays");
  if (nbatch == 1 && tiles * 2 <= num_sms() && K >= 512) {
    int want = (int)std::min<long long>(num_sms() * 2 / tiles, (K + 127) / 128);
    if (want > 1) {
      int kchunk = ((K + want - 1) / want + 15) / 16 * 16;
      int splits = (K + kchunk - 1) / kchunk;
      if (splits > 1) {
        p.splitk = splits; p.kchunk = kchunk;
        // workspace from the per-device caching allocator; released after the reduce kernel is queued (reuse is
        // stream-ordered: one stream per device)
        p.ws = device_alloc((size_t)splits * M * N);
      }
    }
  }
  std::pair<cudaEvent_t, cudaEvent_t>* evs = g_timing ? next_events() : nullptr;
  if (evs) { cudaEventRecord(evs->first, st); g_timed_flops += flops; g_timed_launches += 1; }
  // aligned problems on the 64x64 tile go through the TMA-staged kernel (gemm_tma.cu); everything else through cp.async
  bool used_tma = false;
  // Developer check (AFESP_GEMM_VERIFY=1, unbatched problems): run the cp.async kernel on a copy of C as well and report
  // any TMA result that differs from it by more than rounding.
  static const bool verify = std::getenv("AFESP_GEMM_VERIFY") != nullptr;
  double* vcopy = nullptr;
  const bool do_verify = verify && cfg == 3 && p.splitk == 1 && vec2 && nbatch == 1 && ldc == M;
  if (do_verify) {
    vcopy = device_alloc((size_t)M * N);
    AFESP_CUDA_CHECK(cudaMemcpyAsync(vcopy, C, (size_t)M * N * 8, cudaMemcpyDeviceToDevice, st));
  }
  if (cfg == 3 && p.splitk == 1 && vec2)
    used_tma = dgemm_tma(st, ak, bk, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, batch, p.cvec);
  if (!used_tma) {
    launch_by_cfg(cfg, st, p, nbatch, ak, bk, vec2);
    if (dual) {  // second K segment accumulates into the same C blocks
      Params q = p;
      q.Ap = batch->Aptr2; q.Bp = batch->Bptr2; q.beta = 1.0;
      launch_by_cfg(cfg, st, q, nbatch, ak, bk, vec2);
    }
  }
  if (do_verify) {
    if (used_tma) {
      Params q = p;
      q.C = vcopy;
      launch_by_cfg(cfg, st, q, nbatch, ak, bk, vec2);
      std::vector<double> a((size_t)M * N), b((size_t)M * N);
      AFESP_CUDA_CHECK(cudaMemcpyAsync(a.data(), C, a.size() * 8, cudaMemcpyDeviceToHost, st));
      AFESP_CUDA_CHECK(cudaMemcpyAsync(b.data(), vcopy, b.size() * 8, cudaMemcpyDeviceToHost, st));
      AFESP_CUDA_CHECK(cudaStreamSynchronize(st));
      double md = 0.0, mr = 0.0;
      long long where = -1, nbad = 0;
      for (size_t i = 0; i < a.size(); ++i) {
        const double d = std::fabs(a[i] - b[i]);
        if (d > md) { md = d; where = (long long)i; }
        mr = std::max(mr, std::fabs(b[i]));
      }
      for (size_t i = 0; i < a.size(); ++i) if (std::fabs(a[i] - b[i]) > 1e-11 * mr) ++nbad;
      static int nreport = 0;
      if (md > 1e-11 * mr && nreport < 3) {
        int shown = 0;
        for (size_t i = 0; i < a.size() && shown < 40; ++i)
          if (std::fabs(a[i] - b[i]) > 1e-11 * mr) {
            std::fprintf(stderr, "   bad (m=%lld [tile %lld +%lld], n=%lld [tile %lld +%lld]) tma=%.6e ref=%.6e diff=%.3e\n",
                         (long long)(i % M), (long long)((i % M) / 64), (long long)((i % M) % 64), (long long)(i / M),
                         (long long)((i / M) / 64), (long long)((i / M) % 64), a[i], b[i], a[i] - b[i]);
            ++shown;
          }
      }
      if (md > 1e-11 * mr && nreport++ < 40)
        std::fprintf(stderr, "[afesp gemm verify] TMA != cp.async: M=%d N=%d K=%d ak=%d bk=%d lda=%lld ldb=%lld alpha=%g beta=%g "
                             "maxdiff=%.3e maxref=%.3e at (m=%lld,n=%lld) nbad=%lld A%%128=%d B%%128=%d\n",
                     M, N, K, (int)ak, (int)bk, lda, ldb, alpha, beta, md, mr, where % M, where / M, nbad,
                     (int)(reinterpret_cast<uintptr_t>(A) & 127), (int)(reinterpret_cast<uintptr_t>(B) & 127));
    }
    AFESP_CUDA_CHECK(cudaStreamSynchronize(st));
    device_free(vcopy);
  }
  if (p.splitk > 1) {
    long long MN = (long long)M * N;
    splitk_reduce<<<(int)std::min<long long>((MN + 255) / 256, 2048), 256, 0, st>>>(p.ws, p.splitk, M, N, alpha,
                                                                                   beta, C, ldc);
    count_launch();
    device_free(p.ws);
    AFESP_CUDA_CHECK(cudaGetLastError());
  }
  if (evs) cudaEventRecord(evs->second, st);
}

void gemm_force_config(int cfg) { g_force_cfg = cfg; }
int gemm_force_config_get() { return g_force_cfg; }

void gemm_timing_enable(bool on) {
  if (on && !g_timing) { g_timed_ms = 0.0; g_timed_flops = 0.0; g_timed_launches = 0; g_ev_used = 0; }
  g_timing = on;
}

double gemm_timing_collect(double* flops_out, long long* launches_out) {
  g_timed_ms += drain_events();
  double ms = g_timed_ms;
  if (flops_out) *flops_out = g_timed_flops;
  if (launches_out) *launches_out = g_timed_launches;
  g_timed_ms = 0.0; g_timed_flops = 0.0; g_timed_launches = 0;
  return ms;
}

}  // namespace afesp
