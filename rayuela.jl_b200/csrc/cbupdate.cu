// cbupdate.cu -- "next" row 1: the data-parallel half of the LSQ codebook update,
// fast_bin_matmul (src/codebook_update.jl:96-171): A = B'B + rho*I and b = B'X' for the one-hot code matrix.
//   cooc_kernel   co-occurrence counts of every codebook pair (exact integers; shared-memory privatised
//                 256x256 tile per pair, one block per (pair, slice of n))
//   bxt_kernel    b[(i,c), :] = sum of the vectors whose code in codebook i is c, accumulated in Float64 in
//                 ASCENDING vector index -- the reference's order (:151-161) -- so b is bit-identical.
//                 One block per (i, c): it scans codebook i's code column, compacts the matching indices in
//                 order (ballot + prefix), and d threads add the matching X rows sequentially.
//   assemble_kernel  counts -> symmetric double matrix with the histograms on the block diagonals and +rho.
// The (m*h)^2 dense solve stays with LAPACK on the caller's side, exactly as the reference does
// (getrf!/getrs!, :193-196).
#include <algorithm>

#include "common.cuh"

namespace ryl {

static constexpr int kH = 256;

// grid: (pairs incl. diagonal = m*(m+1)/2, slices).  counts[pair][cj*256 + ci] for i <= j.
__global__ void __launch_bounds__(512) cooc_kernel(const uint8_t* __restrict__ B, int64_t n, int m,
                                                   unsigned int* __restrict__ counts, int64_t per_slice) {
  extern __shared__ unsigned int tile[];
  // 256x256 32-bit counters are 256 KB; shared memory holds 227 KB, so privatise in two passes over ci halves.
  int pair = blockIdx.x, i = 0;
  while (pair >= m - i) {
    pair -= m - i;
    i++;
  }
  const int j = i + pair;
  const int64_t b0 = (int64_t)blockIdx.y * per_slice, b1 = min(n, b0 + per_slice);
  unsigned int* out = counts + (size_t)blockIdx.x * kH * kH;
  for (int half = 0; half < 2; half++) {
    for (int t = threadIdx.x; t < kH * 128; t += blockDim.x) tile[t] = 0;
    __syncthreads();
    for (int64_t l = b0 + threadIdx.x; l < b1; l += blockDim.x) {
      const int ci = B[l * m + i], cj = B[l * m + j];
      if ((ci >> 7) == half) atomicAdd(&tile[cj * 128 + (ci & 127)], 1u);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < kH * 128; t += blockDim.x) {
      const unsigned int v = tile[t];
      if (v) atomicAdd(&out[(t >> 7) * kH + half * 128 + (t & 127)], v);
    }
    __syncthreads();
  }
}

// A[(j,cj),(i,ci)] = A[(i,ci),(j,cj)] = counts; diagonal + rho.  A is (mh x mh) double (symmetric).
__global__ void assemble_kernel(const unsigned int* __restrict__ counts, int m, double rho, double* __restrict__ A) {
  const size_t mh = (size_t)m * kH;
  int pair = blockIdx.x, i = 0;
  while (pair >= m - i) {
    pair -= m - i;
    i++;
  }
  const int j = i + pair;
  const unsigned int* c = counts + (size_t)blockIdx.x * kH * kH;
  for (int t = threadIdx.x; t < kH * kH; t += blockDim.x) {
    const int cj = t >> 8, ci = t & 255;
    const size_t ri = (size_t)i * kH + ci, rj = (size_t)j * kH + cj;
    double v = (double)c[t];
    if (i == j) {
      // the diagonal block of B'B is the histogram of codebook i (a vector has one code per codebook)
      v = (ci == cj) ? v + rho : 0.0;
      A[ri * mh + rj] = v;
    } else {
      A[rj * mh + ri] = v;
      A[ri * mh + rj] = v;
    }
  }
}

// Bt[i][l] = B[l][i], rows padded to a multiple of 16 with 0xFF... handled by the n bound in bxt_kernel.
__global__ void transpose_codes_kernel(const uint8_t* __restrict__ B, uint8_t* __restrict__ Bt, int64_t n, int m,
                                       int64_t ld) {
  for (int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; l < n; l += (int64_t)gridDim.x * blockDim.x)
    for (int i = 0; i < m; i++) Bt[(size_t)i * ld + l] = B[l * m + i];
}

// grid: (256 codes, m codebooks); block: 256 threads.  b is column-major (mh x d): b[t*mh + i*256 + c].
// The block walks codebook i's code row 4096 codes at a time (one 16-byte load per thread, SIMD byte compare),
// compacts the matching vector indices IN ASCENDING ORDER into shared memory, then every thread adds its
// dimensions of those vectors sequentially in Float64 -- the reference's accumulation order (:151-161).
__global__ void __launch_bounds__(256) bxt_kernel(const float* __restrict__ X, const uint8_t* __restrict__ Bt,
                                                  int64_t ld, int64_t n, int d, int m, double* __restrict__ b) {
  __shared__ int list[4096];
  __shared__ int wsum[8];
  const int c = blockIdx.x, i = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const size_t mh = (size_t)m * kH;
  const uint8_t* row = Bt + (size_t)i * ld;
  const unsigned pat = 0x01010101u * (unsigned)c;
  double acc[8];
#pragma unroll
  for (int r = 0; r < 8; r++) acc[r] = 0.0;
  for (int64_t base = 0; base < n; base += 4096) {
    const int64_t l0 = base + (int64_t)tid * 16;
    unsigned hits = 0;                                    // bit e set: code l0 + e matches
    if (l0 < n) {
      const uint4 v = *reinterpret_cast<const uint4*>(row + l0);   // ld is a multiple of 16 and padded
      const unsigned wd[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const unsigned eq = __vcmpeq4(wd[q], pat);        // 0xFF per equal byte
        hits |= ((eq & 1u) | ((eq >> 7) & 2u) | ((eq >> 14) & 4u) | ((eq >> 21) & 8u)) << (4 * q);
      }
      if (l0 + 16 > n) hits &= (1u << (int)(n - l0)) - 1u;
    }
    const int cnt = __popc(hits);
    int inc = cnt;                                        // inclusive scan over the block, in thread order
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, inc, off);
      if (lane >= off) inc += t;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    int offs = inc - cnt, total = 0;
    for (int ww = 0; ww < 8; ww++) {
      if (ww < w) offs += wsum[ww];
      total += wsum[ww];
    }
    while (hits) {
      const int e = __ffs(hits) - 1;
      hits &= hits - 1;
      list[offs++] = tid * 16 + e;
    }
    __syncthreads();
    if (d <= 256) {
      // popular codes make this list long and the adds are one dependent Float64 chain per dimension: issue the
      // row loads eight at a time so the chain is not also serialised on memory latency
      if (tid < d) {
        int e = 0;
        for (; e + 8 <= total; e += 8) {
          float v[8];
#pragma unroll
          for (int u = 0; u < 8; u++) v[u] = __ldg(X + (size_t)(base + list[e + u]) * d + tid);
#pragma unroll
          for (int u = 0; u < 8; u++) acc[0] += (double)v[u];
        }
        for (; e < total; e++) acc[0] += (double)__ldg(X + (size_t)(base + list[e]) * d + tid);
      }
    } else {
      for (int e = 0; e < total; e++) {
        const float* x = X + (size_t)(base + list[e]) * d;
#pragma unroll
        for (int r = 0; r < 8; r++) {
          const int t = tid + 256 * r;
          if (t < d) acc[r] += (double)x[t];
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 8; r++) {
    const int t = tid + 256 * r;
    if (t < d) b[(size_t)t * mh + (size_t)i * kH + c] = acc[r];
  }
}

}  // namespace ryl

using namespace ryl;

extern "C" int rayuela_fast_bin_matmul(const float* X, const uint8_t* B, int64_t n, int d, int m, int h, double rho,
                                       double* A, double* b, unsigned flags, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  RYL_ARG(h == kH, "fast_bin_matmul: only h = 256 is supported");
  RYL_ARG(m >= 1 && m <= 16 && n >= 1 && d >= 1 && d <= 2048, "fast_bin_matmul: bad shape (m in 1..16, d <= 2048)");
  RYL_ARG(n < (1ll << 31), "fast_bin_matmul: n must be below 2^31");
  RYL_ARG(X && B && A && b, "fast_bin_matmul: null array");
  const bool dev = flags & RAYUELA_DEVICE_PTRS;
  const size_t mh = (size_t)m * kH;
  InArg<float> x_in;
  InArg<uint8_t> b_in;
  RYL_TRY(x_in.bind(X, (size_t)n * d, dev, s));
  RYL_TRY(b_in.bind(B, (size_t)n * m, dev, s));
  OutArg<double> a_out, bb_out;
  RYL_TRY(a_out.bind(A, mh * mh, dev, s));
  RYL_TRY(bb_out.bind(b, mh * d, dev, s));

  const int npairs = m * (m + 1) / 2;
  DevBuf counts;
  RYL_TRY(counts.alloc((size_t)npairs * kH * kH * sizeof(unsigned int), s));
  RYL_CUDA(cudaMemsetAsync(counts.p, 0, counts.bytes, s));
  const int slices = (int)std::max<int64_t>(1, std::min<int64_t>((n + 65535) / 65536, (4 * sm_count() + npairs - 1) / npairs));
  const int64_t per_slice = (n + slices - 1) / slices;
  const size_t smem = (size_t)kH * 128 * sizeof(unsigned int);
  RYL_CUDA(cudaFuncSetAttribute(cooc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  RYL_LAUNCH(cooc_kernel, dim3(npairs, slices), 512, smem, s, b_in.d, n, m, counts.as<unsigned int>(), per_slice);
  RYL_LAUNCH(assemble_kernel, npairs, 256, 0, s, counts.as<unsigned int>(), m, rho, a_out.d);
  DevBuf bt;
  const int64_t ld = (n + 15) / 16 * 16;
  RYL_TRY(bt.alloc((size_t)m * ld, s));
  RYL_LAUNCH(transpose_codes_kernel, (int)std::min<int64_t>((n + 255) / 256, 148 * 8), 256, 0, s, b_in.d,
             bt.as<uint8_t>(), n, m, ld);
  RYL_LAUNCH(bxt_kernel, dim3(kH, m), 256, 0, s, x_in.d, bt.as<uint8_t>(), ld, n, d, m, bb_out.d);
  RYL_TRY(a_out.flush(s));
  RYL_TRY(bb_out.flush(s));
  if (!dev) RYL_CUDA(cudaStreamSynchronize(s));
  return RAYUELA_OK;
}
