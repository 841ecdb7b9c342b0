// Tensor index permutation with accumulate:  out = alpha * permute(in) + beta * out   (rank <= 6, FP64)
//
// Replaces the reference's omp_reshape (24 fypp-generated 4-index permutations with optional beta,
// src/linalg.fpp:99-156) and the intrinsic reshape(..., order=) call sites in src/ccsd.f90.
// HBM-bound: 16 B moved per element (+8 B when beta != 0).  Two kernels:
//   * axis 0 preserved  -> straight coalesced gather (reads and writes both run along the fastest axis);
//   * axis 0 moved      -> 32x32 shared-memory tile transpose over (input-fastest, output-fastest) axes so that
//                          both the global read and the global write are coalesced 256-byte rows.
// Adjacent axes that stay adjacent are merged first, so e.g. (i,j,a,b)->(a,b,i,j) runs as a 2-D transpose.
#include "common.cuh"

namespace afesp {
namespace {

constexpr int MAXR = 6;

struct PermParams {
  int rank;
  int odims[MAXR];          // output extents
  long long istr[MAXR];     // input stride of the input axis feeding output axis d
  long long ostr[MAXR];     // output stride of output axis d (ostr[0] == 1)
  long long total;
  double alpha, beta;
};

__global__ void permute_gather(const PermParams p, const double* __restrict__ in, double* __restrict__ out) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < p.total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long rem = idx, off = 0, oo = 0;
#pragma unroll
    for (int d = 0; d < MAXR; ++d) {
      if (d < p.rank) {
        long long q = rem / p.odims[d];
        long long c = rem - q * p.odims[d];
        off += c * p.istr[d];
        oo += c * p.ostr[d];
        rem = q;
      }
    }
    double v = p.alpha * in[off];
    if (p.beta != 0.0) v += p.beta * out[oo];
    out[oo] = v;
  }
}

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
    if (r > 0 && istr[r - 1] * odims[r - 1] == istr_full[a] && ostr_d[r - 1] * odims[r - 1] == ostr_full[d]) {
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
    PermParams p{};
    p.rank = r;
    for (int d = 0; d < r; ++d) { p.odims[d] = odims[d]; p.istr[d] = istr[d]; p.ostr[d] = ostr_d[d]; }
    p.total = total; p.alpha = alpha; p.beta = beta;
    int blocks = (int)std::min<long long>((total + 255) / 256, 148LL * 32);
    permute_gather<<<blocks, 256, 0, st>>>(p, in, out);
  } else {
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
    dim3 grid((p.n0 + 31) / 32, (unsigned)std::min<long long>((p.nb + 31) / 32, 65535),
              (unsigned)std::min<long long>(p.rest_total, 65535));
    permute_transpose<<<grid, dim3(32, 8), 0, st>>>(p, in, out);
  }
  count_launch();
  AFESP_CUDA_CHECK(cudaGetLastError());
}

}  // namespace afesp
