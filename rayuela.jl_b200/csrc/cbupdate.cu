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

// grid: (256 codes, m codebooks); block: 256 threads.  b is column-major (mh x d): b[t*mh + i*256 + c].
__global__ void __launch_bounds__(256) bxt_kernel(const float* __restrict__ X, const uint8_t* __restrict__ B,
                                                  int64_t n, int d, int m, double* __restrict__ b) {
  __shared__ int list[256];
  __shared__ int wcount[8];
  __shared__ int total_s;
  const int c = blockIdx.x, i = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const size_t mh = (size_t)m * kH;
  // every thread owns dimensions tid, tid+256, ... (up to 8 => d <= 2048)
  double acc[8];
#pragma unroll
  for (int r = 0; r < 8; r++) acc[r] = 0.0;
  int fill = 0;
  for (int64_t base = 0; base < n; base += 256) {
    const int64_t l = base + tid;
    const bool hit = l < n && B[l * m + i] == c;
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) wcount[w] = __popc(bal);
    __syncthreads();
    int off = fill;
    for (int ww = 0; ww < w; ww++) off += wcount[ww];
    if (hit) list[off + __popc(bal & ((1u << lane) - 1u))] = (int)(l - base);
    if (tid == 0) {
      int t = fill;
      for (int ww = 0; ww < 8; ww++) t += wcount[ww];
      total_s = t;
    }
    __syncthreads();
    fill = total_s;
    // consume this window's matches (ascending l) before moving on: on average one match per 256 codes
    if (fill > 0) {
      for (int e = 0; e < fill; e++) {
        const float* x = X + (size_t)(base + list[e]) * d;
#pragma unroll
        for (int r = 0; r < 8; r++) {
          const int t = tid + 256 * r;
          if (t < d) acc[r] += (double)x[t];
        }
      }
      fill = 0;
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
  RYL_LAUNCH(bxt_kernel, dim3(kH, m), 256, 0, s, x_in.d, b_in.d, n, d, m, bb_out.d);
  RYL_TRY(a_out.flush(s));
  RYL_TRY(bb_out.flush(s));
  if (!dev) RYL_CUDA(cudaStreamSynchronize(s));
  return RAYUELA_OK;
}
