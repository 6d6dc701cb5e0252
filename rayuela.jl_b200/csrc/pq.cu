// pq.cu -- K7: PQ / OPQ nearest-centroid encode (quantize_pq, src/PQ.jl:18-48).
// The reference computes Distances.pairwise(SqEuclidean(), C[i], X[subdims[i],:]) -- i.e.
// max(||c||^2 + ||x||^2 - 2<c,x>, 0) -- then Clustering.update_assignments! (first minimum, strict <).
// Here: one thread per (vector, subspace); the subspace's 256 centroids sit in shared memory and are read
// as warp-wide broadcasts; the running first-minimum is thread-local, so there is no reduction at all.
// Dots and norms are sequential-t fmaf chains (the oracle's order; the reference's BLAS order is unpinned).
#include <algorithm>

#include "common.cuh"

namespace ryl {

static constexpr int kH = 256;

template <int SUB>  // SUB > 0: subspace dim known at compile time (x in registers); 0: generic
__global__ void __launch_bounds__(256) pq_encode_kernel(const float* __restrict__ X, const float* __restrict__ Cpq,
                                                        int64_t n, int d, int m, int sub_rt,
                                                        uint8_t* __restrict__ B) {
  extern __shared__ __align__(16) float smem[];
  const int sub = SUB > 0 ? SUB : sub_rt;
  float* cs = smem;              // [256][sub]
  float* cn = smem + kH * sub;   // [256]
  const int k = blockIdx.y;
  for (int i = threadIdx.x; i < kH * sub; i += blockDim.x) cs[i] = Cpq[(size_t)k * kH * sub + i];
  __syncthreads();
  {
    const float* c = cs + threadIdx.x * sub;
    float s = 0.f;
    for (int t = 0; t < sub; t++) s = fmaf(c[t], c[t], s);
    cn[threadIdx.x] = s;
  }
  __syncthreads();
  for (int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; l < n; l += (int64_t)gridDim.x * blockDim.x) {
    const float* x = X + (size_t)l * d + (size_t)k * sub;
    float xr[SUB > 0 ? SUB : 1];
    float xn = 0.f;
    if (SUB > 0) {
#pragma unroll
      for (int t = 0; t < SUB; t++) { xr[t] = x[t]; xn = fmaf(xr[t], xr[t], xn); }
    } else {
      for (int t = 0; t < sub; t++) xn = fmaf(x[t], x[t], xn);
    }
    float best = 0.f;
    int bi = 0;
    for (int c = 0; c < kH; c++) {
      const float* cv = cs + c * sub;
      float r = 0.f;
      if (SUB > 0) {
#pragma unroll
        for (int t = 0; t < SUB; t++) r = fmaf(cv[t], xr[t], r);
      } else {
        for (int t = 0; t < sub; t++) r = fmaf(cv[t], x[t], r);
      }
      float v = __fsub_rn(__fadd_rn(cn[c], xn), __fmul_rn(2.0f, r));
      v = v > 0.0f ? v : 0.0f;
      if (c == 0 || v < best) { best = v; bi = c; }
    }
    B[(size_t)l * m + k] = (uint8_t)bi;
  }
}

}  // namespace ryl

using namespace ryl;

template <int SUB>
static int launch_pq(const float* X, const float* Cpq, int64_t n, int d, int m, int sub, uint8_t* B, cudaStream_t s) {
  size_t smem = ((size_t)kH * sub + kH) * sizeof(float);
  RYL_ARG(smem <= 200 * 1024, "quantize_pq: subspace dimension too large for shared memory");
  RYL_CUDA(cudaFuncSetAttribute(pq_encode_kernel<SUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int gx = (int)std::min<int64_t>((n + 255) / 256, (int64_t)sm_count() * 8);
  RYL_LAUNCH(pq_encode_kernel<SUB>, dim3(gx, m), 256, smem, s, X, Cpq, n, d, m, sub, B);
  return RAYUELA_OK;
}

extern "C" int rayuela_quantize_pq(const float* X, const float* Cpq, int64_t n, int d, int m, int h, uint8_t* B,
                                   unsigned flags, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  RYL_ARG(h == kH, "quantize_pq: only h = 256 is supported");
  RYL_ARG(m >= 1 && d >= m && d % m == 0, "quantize_pq: d must be a positive multiple of m");
  RYL_ARG(n >= 1 && X && Cpq && B, "quantize_pq: bad n or null array");
  const bool dev = flags & RAYUELA_DEVICE_PTRS;
  const int sub = d / m;
  InArg<float> x_in, c_in;
  RYL_TRY(x_in.bind(X, (size_t)n * d, dev, s));
  RYL_TRY(c_in.bind(Cpq, (size_t)m * kH * sub, dev, s));
  OutArg<uint8_t> b_out;
  RYL_TRY(b_out.bind(B, (size_t)n * m, dev, s));
  int rc;
  switch (sub) {
    case 4: rc = launch_pq<4>(x_in.d, c_in.d, n, d, m, sub, b_out.d, s); break;
    case 8: rc = launch_pq<8>(x_in.d, c_in.d, n, d, m, sub, b_out.d, s); break;
    case 16: rc = launch_pq<16>(x_in.d, c_in.d, n, d, m, sub, b_out.d, s); break;
    case 32: rc = launch_pq<32>(x_in.d, c_in.d, n, d, m, sub, b_out.d, s); break;
    default: rc = launch_pq<0>(x_in.d, c_in.d, n, d, m, sub, b_out.d, s); break;
  }
  RYL_TRY(rc);
  RYL_TRY(b_out.flush(s));
  if (!dev) RYL_CUDA(cudaStreamSynchronize(s));
  return RAYUELA_OK;
}
