// Tensor index permutation with accumulate:  out = alpha * permute(in) + beta * out   (rank <= 6, FP64)
//
// Replaces the reference's omp_reshape (24 fypp-generated 4-index permutations with optional beta,
// src/linalg.fpp:99-156) and the intrinsic reshape(..., order=) call sites in src/ccsd.f90.
// HBM-bound: 16 B moved per element (+8 B when beta != 0).  Adjacent axes that stay adjacent are merged first
// (e.g. (i,j,a,b)->(a,b,i,j) runs as a 2-D transpose); then one of four kernels:
//   * permute_rows          axis 0 preserved: threads walk contiguous rows, the row index is decoded once per row;
//   * permute_slab          leading axes shuffled among themselves, <= 4000 elements: contiguous slabs staged through
//                           shared memory, the in-slab permutation comes from a table built once per block;
//   * permute_transpose64   axis 0 moved, both transposed extents >= 48: 64x64 shared-memory tile, 512-byte rows;
//   * permute_transpose     axis 0 moved, small extents: 32x32 shared-memory tile.
#include <algorithm>

#include "common.cuh"

namespace afesp {
namespace {

constexpr int MAXR = 6;

struct TransParams {
  int n0, nb;                 // extents of input axis 0 and of the input axis that becomes output axis 0
  long long istr_b;           // input stride of that axis
  long long ostr_0;           // output stride of input axis 0
  int nrest;
  int rdims[MAXR];            // remaining axes (extents), with their input and output strides
  long long ristr[MAXR], rostr[MAXR];
  long long rest_total;
  double alpha, beta;
};

__global__ void permute_transpose(const TransParams p, const double* __restrict__ in, double* __restrict__ out) {
  __shared__ double tile[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int i0 = blockIdx.x * 32;
  const int tiles_b = (p.nb + 31) / 32;
  for (int by = blockIdx.y; by < tiles_b; by += gridDim.y)
  for (long long rest = blockIdx.z; rest < p.rest_total; rest += gridDim.z) {
    const int b0 = by * 32;
    long long rem = rest, ibase = 0, obase = 0;
    for (int d = 0; d < p.nrest; ++d) {
      long long q = rem / p.rdims[d];
      long long c = rem - q * p.rdims[d];
      ibase += c * p.ristr[d];
      obase += c * p.rostr[d];
      rem = q;
    }
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
      int i = i0 + tx, b = b0 + r;
      if (i < p.n0 && b < p.nb) tile[r][tx] = in[ibase + i + (long long)b * p.istr_b];
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
      int i = i0 + r, b = b0 + tx;
      if (i < p.n0 && b < p.nb) {
        double* o = out + obase + b + (long long)i * p.ostr_0;
        double v = p.alpha * tile[tx][r];
        if (p.beta != 0.0) v += p.beta * (*o);
        *o = v;
      }
    }
    __syncthreads();
  }
}

// Axis 0 preserved, row form: a "row" is one run of the (merged) fastest axis, contiguous in both arrays.  Threads
// walk along the row; the row index is decoded once per row and thread (32-bit arithmetic), not once per element.
struct RowParams {
  int n0;                    // row length
  int nrest;
  int rdims[MAXR];
  long long ristr[MAXR], rostr[MAXR];
  long long rows;
  double alpha, beta;
};

template <int RX>   // threads along the row; 256 / RX rows per block pass
__global__ void __launch_bounds__(256) permute_rows(const RowParams p, const double* __restrict__ in,
                                                    double* __restrict__ out) {
  constexpr int RY = 256 / RX;
  const int tx = threadIdx.x % RX, ty = threadIdx.x / RX;
  for (long long row = (long long)blockIdx.x * RY + ty; row < p.rows; row += (long long)gridDim.x * RY) {
    long long ib = 0, ob = 0;
    if (p.rows < (1LL << 31)) {
      unsigned rem = (unsigned)row;
      for (int d = 0; d < p.nrest; ++d) {
        const unsigned q = rem / (unsigned)p.rdims[d], c = rem - q * (unsigned)p.rdims[d];
        ib += c * p.ristr[d]; ob += c * p.rostr[d];
        rem = q;
      }
    } else {
      long long rem = row;
      for (int d = 0; d < p.nrest; ++d) {
        const long long q = rem / p.rdims[d], c = rem - q * p.rdims[d];
        ib += c * p.ristr[d]; ob += c * p.rostr[d];
        rem = q;
      }
    }
    const double* src = in + ib;
    double* dst = out + ob;
    if (p.beta == 0.0) {
      for (int x = tx; x < p.n0; x += RX) dst[x] = p.alpha * src[x];
    } else {
      for (int x = tx; x < p.n0; x += RX) dst[x] = p.alpha * src[x] + p.beta * dst[x];
    }
  }
}

// Leading axes shuffled among themselves, small slab: the first k (merged) axes of the output are a permutation of
// the first k axes of the input and hold S <= 4096 elements, so every slab is one contiguous run in both arrays
// (e.g. (i,j,a,b) -> (j,i,a,b) with o <= 64).  A block stages G slabs through shared memory with fully coalesced
// loads and stores; the in-slab permutation is a table built once per block; slab bases are decoded once per slab.
struct SlabParams {
  int k;
  int sdims[MAXR];     // output extents of the slab axes
  int sistr[MAXR];     // input stride (inside the slab) of the axis feeding output slab axis d
  int S, G;
  int pad_every;       // one padding word after every `pad_every` slab elements (0: none), see below
  int Sp;              // padded slab size in shared memory
  int nrest;
  int rdims[MAXR];
  long long ristr[MAXR], rostr[MAXR];
  long long rest_total;
  double alpha, beta;
};

// Shared-memory position of slab element `off` (input order).  The permuted read walks the slab with the input stride of
// the output-fastest axis -- e.g. stride o for (i,j)->(j,i): with o = 20 or 40 the 16 lanes of a half-warp would hit 4
// or 2 of the 16 eight-byte banks.  One padding word per run of the input-fastest axis makes that stride odd.
__device__ __forceinline__ int slab_pos(int off, int pad_every) { return pad_every ? off + off / pad_every : off; }
// the same without an integer division, for the streaming loop: off < 4096 and pad_every >= 8, so the float quotient of
// (off + 0.5) is at least 0.5/pad_every away from an integer -- far more than its rounding error
__device__ __forceinline__ int slab_pos_fast(int off, float inv_pad) { return off + (int)(((float)off + 0.5f) * inv_pad); }

__global__ void __launch_bounds__(256) permute_slab(const SlabParams p, const double* __restrict__ in,
                                                    double* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char slab_sm[];
  double* buf = reinterpret_cast<double*>(slab_sm);                       // [G*Sp]
  long long* ibase = reinterpret_cast<long long*>(buf + (size_t)p.G * p.Sp);  // [G]
  long long* obase = ibase + p.G;                                         // [G]
  int* tab = reinterpret_cast<int*>(obase + p.G);                         // [S] padded position read by output element s
  const int tid = threadIdx.x, S = p.S, Sp = p.Sp;
  const float inv_pad = p.pad_every ? 1.0f / (float)p.pad_every : 0.0f;
  for (int s = tid; s < S; s += 256) {
    int rem = s, off = 0;
    for (int d = 0; d < p.k; ++d) {
      const int q = rem / p.sdims[d], c = rem - q * p.sdims[d];
      off += c * p.sistr[d];
      rem = q;
    }
    tab[s] = slab_pos(off, p.pad_every);
  }
  const int s0 = tid % S, g0 = tid / S;
  for (long long slab0 = (long long)blockIdx.x * p.G; slab0 < p.rest_total; slab0 += (long long)gridDim.x * p.G) {
    const int ng = (int)min((long long)p.G, p.rest_total - slab0);
    __syncthreads();
    for (int t = tid; t < ng; t += 256) {
      long long rem = slab0 + t, ib = 0, ob = 0;
      for (int d = 0; d < p.nrest; ++d) {
        const long long q = rem / p.rdims[d], c = rem - q * p.rdims[d];
        ib += c * p.ristr[d]; ob += c * p.rostr[d];
        rem = q;
      }
      ibase[t] = ib; obase[t] = ob;
    }
    __syncthreads();
    const int n = ng * S;
    int s = s0, g = g0;
    int e = tid;
    for (; e + 3 * 256 < n; e += 4 * 256) {   // four independent loads in flight per thread
      double v[4];
      int pos[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        v[u] = in[ibase[g] + s];
        pos[u] = g * Sp + slab_pos_fast(s, inv_pad);
        s += 256;
        while (s >= S) { s -= S; ++g; }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) buf[pos[u]] = v[u];
    }
    for (; e < n; e += 256) {
      buf[g * Sp + slab_pos_fast(s, inv_pad)] = in[ibase[g] + s];
      s += 256;
      while (s >= S) { s -= S; ++g; }
    }
    __syncthreads();
    s = s0; g = g0;
    if (p.beta == 0.0) {
      for (int e2 = tid; e2 < n; e2 += 256) {
        out[obase[g] + s] = p.alpha * buf[g * Sp + tab[s]];
        s += 256;
        while (s >= S) { s -= S; ++g; }
      }
    } else {
      int e2 = tid;
      for (; e2 + 3 * 256 < n; e2 += 4 * 256) {   // the four reads of `out` in flight together
        double* o[4];
        double old[4], nv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          o[u] = out + obase[g] + s;
          old[u] = *o[u];
          nv[u] = buf[g * Sp + tab[s]];
          s += 256;
          while (s >= S) { s -= S; ++g; }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) *o[u] = p.alpha * nv[u] + p.beta * old[u];
      }
      for (; e2 < n; e2 += 256) {
        double* o = out + obase[g] + s;
        *o = p.alpha * buf[g * Sp + tab[s]] + p.beta * (*o);
        s += 256;
        while (s >= S) { s -= S; ++g; }
      }
    }
  }
}

// Transpose for a SHORT input-fastest axis (n0 < 48, e.g. the occupied index o = 20 / 40 of (i,j,a,b) -> (b,a,j,i)):
// a tile holds the whole run of n0 elements for 64 values of the axis that becomes output-fastest, so the stores are
// 512-byte rows and the loads are complete n0*8-byte runs (a 32x32 tile would run the second tile column 1/4 full).
__global__ void __launch_bounds__(256) permute_transpose_narrow(const TransParams p, const double* __restrict__ in,
                                                                double* __restrict__ out) {
  __shared__ double tile[64][49];
  const int n0 = p.n0;
  const int tiles_b = (p.nb + 63) / 64;
  const float inv = 1.0f / (float)n0;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  for (int by = blockIdx.y; by < tiles_b; by += gridDim.y)
  for (long long rest = blockIdx.z; rest < p.rest_total; rest += gridDim.z) {
    const int b0 = by * 64;
    long long rem = rest, ibase = 0, obase = 0;
    for (int d = 0; d < p.nrest; ++d) {
      long long q = rem / p.rdims[d];
      long long c = rem - q * p.rdims[d];
      ibase += c * p.ristr[d];
      obase += c * p.rostr[d];
      rem = q;
    }
    for (int e = threadIdx.x; e < 64 * n0; e += 256) {
      int r = (int)(((float)e + 0.5f) * inv);      // e / n0 (exact for e < 64 * 48)
      int i = e - r * n0;
      if (i < 0) { --r; i += n0; } else if (i >= n0) { ++r; i -= n0; }
      if (b0 + r < p.nb) tile[r][i] = in[ibase + i + (long long)(b0 + r) * p.istr_b];
    }
    __syncthreads();
    for (int i = ty; i < n0; i += 4) {
      const int b = b0 + tx;
      if (b < p.nb) {
        double* o = out + obase + b + (long long)i * p.ostr_0;
        double v = p.alpha * tile[tx][i];
        if (p.beta != 0.0) v += p.beta * (*o);
        *o = v;
      }
    }
    __syncthreads();
  }
}

// 64x64 shared-memory tile transpose (512-byte rows on both sides), used when both transposed extents are >= 48.
__global__ void __launch_bounds__(256) permute_transpose64(const TransParams p, const double* __restrict__ in,
                                                           double* __restrict__ out) {
  __shared__ double tile[64][65];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;   // 64 x 4
  const int i0 = blockIdx.x * 64;
  const int tiles_b = (p.nb + 63) / 64;
  for (int by = blockIdx.y; by < tiles_b; by += gridDim.y)
  for (long long rest = blockIdx.z; rest < p.rest_total; rest += gridDim.z) {
    const int b0 = by * 64;
    long long rem = rest, ibase = 0, obase = 0;
    for (int d = 0; d < p.nrest; ++d) {
      long long q = rem / p.rdims[d];
      long long c = rem - q * p.rdims[d];
      ibase += c * p.ristr[d];
      obase += c * p.rostr[d];
      rem = q;
    }
#pragma unroll 4
    for (int r = ty; r < 64; r += 4) {
      const int i = i0 + tx, b = b0 + r;
      if (i < p.n0 && b < p.nb) tile[r][tx] = in[ibase + i + (long long)b * p.istr_b];
    }
    __syncthreads();
#pragma unroll 4
    for (int r = ty; r < 64; r += 4) {
      const int i = i0 + r, b = b0 + tx;
      if (i < p.n0 && b < p.nb) {
        double* o = out + obase + b + (long long)i * p.ostr_0;
        double v = p.alpha * tile[tx][r];
        if (p.beta != 0.0) v += p.beta * (*o);
        *o = v;
      }
    }
    __syncthreads();
  }
}

}  // namespace

void permute(cudaStream_t st, int rank, const int* dims, const int* perm, double alpha, const double* in, double beta,
             double* out) {
  permute_strided(st, rank, dims, perm, alpha, in, beta, out, nullptr);
}

// Same, with explicit output strides (ostride[d] = stride of output axis d, ostride[0] must be 1): writes a permuted
// tensor into a sub-block of a larger array (used to assemble the K-concatenated (T) operands).
void permute_strided(cudaStream_t st, int rank, const int* dims, const int* perm, double alpha, const double* in,
                     double beta, double* out, const long long* ostride) {
  AFESP_REQUIRE(rank >= 1 && rank <= MAXR, "permute: rank must be 1..6");
  // validate + input strides
  long long istr_full[MAXR];
  bool seen[MAXR] = {false, false, false, false, false, false};
  long long total = 1;
  for (int d = 0; d < rank; ++d) {
    AFESP_REQUIRE(perm[d] >= 0 && perm[d] < rank && !seen[perm[d]], "permute: perm is not a permutation");
    seen[perm[d]] = true;
    istr_full[d] = total;
    total *= dims[d];
  }
  long long ostr_full[MAXR];
  {
    long long acc = 1;
    for (int d = 0; d < rank; ++d) {
      ostr_full[d] = ostride ? ostride[d] : acc;
      acc *= dims[perm[d]];
    }
    AFESP_REQUIRE(ostr_full[0] == 1, "permute: output axis 0 must have unit stride");
  }
  if (total == 0) return;
  // Canonical form in output order: drop extent-1 axes, merge output-adjacent axes that are input-adjacent.
  int r = 0;
  int odims[MAXR];
  long long istr[MAXR], ostr_d[MAXR];
  for (int d = 0; d < rank; ++d) {
    int a = perm[d];
    if (dims[a] == 1) continue;
    if (r > 0 && istr[r - 1] * odims[r - 1] == istr_full[a] && ostr_d[r - 1] * odims[r - 1] == ostr_full[d] &&
        (long long)odims[r - 1] * dims[a] < (1LL << 31)) {   // merged extents stay 32-bit: larger arrays keep two axes
      odims[r - 1] *= dims[a];
    } else {
      odims[r] = dims[a];
      istr[r] = istr_full[a];
      ostr_d[r] = ostr_full[d];
      ++r;
    }
  }
  if (r == 0) { odims[0] = 1; istr[0] = 1; ostr_d[0] = 1; r = 1; }
  AFESP_REQUIRE(ostr_d[0] == 1, "permute: leading output axis must have unit stride");
  if (istr[0] == 1) {
    // axis 0 preserved: rows of odims[0] contiguous elements on both sides
    RowParams p{};
    p.n0 = odims[0];
    p.nrest = r - 1;
    p.rows = 1;
    for (int d = 1; d < r; ++d) {
      p.rdims[d - 1] = odims[d]; p.ristr[d - 1] = istr[d]; p.rostr[d - 1] = ostr_d[d];
      p.rows *= odims[d];
    }
    p.alpha = alpha; p.beta = beta;
    // threads along the row: the width that wastes the fewest lanes in the last pass (ties go to the wider one)
    int rx = 32;
    double best = -1.0;
    for (int w : {256, 128, 64, 32, 16, 8}) {
      const double eff = (double)p.n0 / ((double)((p.n0 + w - 1) / w) * w);
      if (eff > best + 1e-9) { best = eff; rx = w; }
    }
    const long long passes = (p.rows + (256 / rx) - 1) / (256 / rx);
    const int blocks = (int)std::max<long long>(1, std::min<long long>(passes, 148LL * 16));
    switch (rx) {
      case 256: permute_rows<256><<<blocks, 256, 0, st>>>(p, in, out); break;
      case 128: permute_rows<128><<<blocks, 256, 0, st>>>(p, in, out); break;
      case 64: permute_rows<64><<<blocks, 256, 0, st>>>(p, in, out); break;
      case 32: permute_rows<32><<<blocks, 256, 0, st>>>(p, in, out); break;
      case 16: permute_rows<16><<<blocks, 256, 0, st>>>(p, in, out); break;   // short rows (e.g. o = 40): sub-warp groups
      default: permute_rows<8><<<blocks, 256, 0, st>>>(p, in, out); break;
    }
  } else {
    // slab form?  smallest k >= 2 whose k leading output axes are exactly the k leading input axes, densely packed
    int kslab = 0;
    long long S = 1;
    for (int k = 2; k <= r && kslab == 0; ++k) {
      bool dense_out = true;
      long long acc = 1;
      for (int d = 0; d < k; ++d) { dense_out = dense_out && (ostr_d[d] == acc); acc *= odims[d]; }
      if (!dense_out || acc > 3600) break;   // buffer (+ padding) + tables + bases must fit the 48 KB default dynamic shared memory
      // input side: the same axes, sorted by input stride, must tile [0, acc) densely
      int idx[MAXR];
      for (int d = 0; d < k; ++d) idx[d] = d;
      std::sort(idx, idx + k, [&](int a, int b) { return istr[a] < istr[b]; });
      long long exp = 1;
      bool dense_in = true;
      for (int m = 0; m < k; ++m) { dense_in = dense_in && (istr[idx[m]] == exp); exp *= odims[idx[m]]; }
      if (dense_in) { kslab = k; S = acc; }
    }
    if (kslab > 0) {
      SlabParams p{};
      p.k = kslab; p.S = (int)S;
      // padding: one word per run of the slab's input-fastest axis when that run length is even (odd strides already
      // spread over the banks)
      int run = 0;
      for (int d = 0; d < kslab; ++d) if (istr[d] == 1) run = odims[d];
      p.pad_every = (run >= 8 && run % 2 == 0 && run < S) ? run : 0;
      p.Sp = p.pad_every ? (int)(S + S / p.pad_every) : (int)S;
      p.G = (int)std::max<long long>(1, 4000 / p.Sp);
      for (int d = 0; d < kslab; ++d) { p.sdims[d] = odims[d]; p.sistr[d] = (int)istr[d]; }
      p.nrest = r - kslab; p.rest_total = 1;
      for (int d = kslab; d < r; ++d) {
        p.rdims[d - kslab] = odims[d]; p.ristr[d - kslab] = istr[d]; p.rostr[d - kslab] = ostr_d[d];
        p.rest_total *= odims[d];
      }
      p.alpha = alpha; p.beta = beta;
      const size_t smem = (size_t)p.G * p.Sp * 8 + (size_t)p.G * 16 + (size_t)p.S * 4;   // <= 3600 * (9 + 4) + ... < 48 KB
      const long long passes = (p.rest_total + p.G - 1) / p.G;
      const int blocks = (int)std::max<long long>(1, std::min<long long>(passes, 148LL * 8));
      permute_slab<<<blocks, 256, smem, st>>>(p, in, out);
      count_launch();
      AFESP_CUDA_CHECK(cudaGetLastError());
      return;
    }
    // find the output axis fed by input axis 0 (input stride 1)
    int d0 = -1;
    for (int d = 1; d < r; ++d) if (istr[d] == 1) d0 = d;
    AFESP_REQUIRE(d0 > 0, "permute: internal error (no unit-stride axis)");
    TransParams p{};
    p.n0 = odims[d0];
    p.nb = odims[0];
    p.istr_b = istr[0];
    p.ostr_0 = ostr_d[d0];
    p.nrest = 0; p.rest_total = 1;
    for (int d = 1; d < r; ++d) {
      if (d == d0) continue;
      p.rdims[p.nrest] = odims[d]; p.ristr[p.nrest] = istr[d]; p.rostr[p.nrest] = ostr_d[d];
      p.rest_total *= odims[d];
      ++p.nrest;
    }
    p.alpha = alpha; p.beta = beta;
    if (p.n0 < 48 && p.nb >= 64) {
      dim3 grid(1, (unsigned)std::min<long long>((p.nb + 63) / 64, 65535), (unsigned)std::min<long long>(p.rest_total, 65535));
      permute_transpose_narrow<<<grid, 256, 0, st>>>(p, in, out);
    } else if (p.n0 >= 48 && p.nb >= 48) {
      dim3 grid((p.n0 + 63) / 64, (unsigned)std::min<long long>((p.nb + 63) / 64, 65535),
                (unsigned)std::min<long long>(p.rest_total, 65535));
      permute_transpose64<<<grid, 256, 0, st>>>(p, in, out);
    } else {
      dim3 grid((p.n0 + 31) / 32, (unsigned)std::min<long long>((p.nb + 31) / 32, 65535),
                (unsigned)std::min<long long>(p.rest_total, 65535));
      permute_transpose<<<grid, dim3(32, 8), 0, st>>>(p, in, out);
    }
  }
  count_launch();
  AFESP_CUDA_CHECK(cudaGetLastError());
}

}  // namespace afesp
