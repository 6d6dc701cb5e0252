// Microbenchmark: what bounds the skewed-scan inner step on B200?  One 512-thread block per SM, 128 KB LUT tile.
//   V=0  4 LDS.64 + 4 FADD2 per step           (shared-memory pipe alone)
//   V=1  4 LDS.64 + 8 FFMA2 per step           (the scan's mix)
//   V=2  4 LDS.64 + 16 FFMA  per step          (same math, scalar)
//   V=3  8 FFMA2 per step, no LDS              (FFMA2 pipe alone)
//   V=4  16 FFMA per step, no LDS
//   V=5  4 LDS.64 + 4 FFMA2 + 4 FADD2
//   V=6  2 LDS.64 + 4 FFMA2                    (the pre-filter scan's mix: 4 queries per LDS.64)
//   V=7  2 LDS.64 + 2 FADD2                    (shared-memory pipe alone at that width)
//   V=8  2 LDS.64 + 4 IMAD + 4 LOP3            (the same step on packed 16-bit integers)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_ffma2 lds_ffma2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
template <int IMM> __device__ __forceinline__ uint64_t lds64(uint32_t addr) {
  uint64_t v; asm volatile("ld.shared.b64 %0, [%1+%2];" : "=l"(v) : "r"(addr), "n"(IMM)); return v; }
__device__ __forceinline__ uint64_t pack2(float lo, float hi) { return ((uint64_t)__float_as_uint(hi) << 32) | __float_as_uint(lo); }

template <int V>
__global__ void __launch_bounds__(512, 1) k(const uint32_t* __restrict__ codes, float* out, int iters, long long* clk) {
  extern __shared__ __align__(16) unsigned char smem[];
  for (int i = threadIdx.x; i < 32768; i += 512) reinterpret_cast<float*>(smem)[i] = 1.0f / (1 + (i & 1023));
  __syncthreads();
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
  const int lane = threadIdx.x & 31;
  const uint32_t off = base + ((lane & 15) << 3);
  uint64_t acc[4] = {0, 0, 0, 0}, done[4] = {0, 0, 0, 0};
  const uint64_t kp = pack2(lane == 40 ? 0.f : 1.f, lane == 40 ? 0.f : 1.f), cp = pack2(lane == 41 ? 1.f : 0.f, lane == 41 ? 1.f : 0.f);
  uint32_t w = codes[threadIdx.x + blockIdx.x * 512];
  uint32_t ki = lane == 40 ? 0u : 1u, ci = lane == 41 ? 0xFFFFFFFFu : 0u;
  asm volatile("" : "+r"(ki), "+r"(ci));
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int s = 0; s < 8; s++) {
      const uint32_t a = ((w >> (s * 3)) & 0x7F80u) + off;
      uint64_t v0, v1, v2, v3;
      if (V <= 2 || V == 5) { v0 = lds64<0>(a); v1 = lds64<32768>(a); v2 = lds64<65536>(a); v3 = lds64<98304>(a); }
      else if (V >= 6) { v0 = lds64<0>(a); v1 = lds64<32768>(a); v2 = 0; v3 = 0; }
      else { v0 = a; v1 = a + 1; v2 = a + 2; v3 = a + 3; }
      if (V == 6) { acc[0] = ffma2(acc[0], kp, v0); acc[1] = ffma2(acc[1], kp, v1); done[0] = ffma2(acc[0], cp, done[0]); done[1] = ffma2(acc[1], cp, done[1]); }
      if (V == 7) { acc[0] = fadd2(acc[0], v0); acc[1] = fadd2(acc[1], v1); }
      if (V == 8) {
        uint32_t x[4] = {(uint32_t)v0, (uint32_t)(v0 >> 32), (uint32_t)v1, (uint32_t)(v1 >> 32)};
        uint32_t* a32 = reinterpret_cast<uint32_t*>(acc); uint32_t* d32 = reinterpret_cast<uint32_t*>(done);
#pragma unroll
        for (int i = 0; i < 4; i++) { a32[i] = a32[i] * ki + x[i]; d32[i] |= a32[i] & ci; }
      }
      if (V == 0) { acc[0] = fadd2(acc[0], v0); acc[1] = fadd2(acc[1], v1); acc[2] = fadd2(acc[2], v2); acc[3] = fadd2(acc[3], v3); }
      if (V == 1 || V == 3) {
        acc[0] = ffma2(acc[0], kp, v0); acc[1] = ffma2(acc[1], kp, v1); acc[2] = ffma2(acc[2], kp, v2); acc[3] = ffma2(acc[3], kp, v3);
        done[0] = ffma2(acc[0], cp, done[0]); done[1] = ffma2(acc[1], cp, done[1]); done[2] = ffma2(acc[2], cp, done[2]); done[3] = ffma2(acc[3], cp, done[3]);
      }
      if (V == 5) {
        acc[0] = ffma2(acc[0], kp, v0); acc[1] = ffma2(acc[1], kp, v1); acc[2] = ffma2(acc[2], kp, v2); acc[3] = ffma2(acc[3], kp, v3);
        done[0] = fadd2(acc[0], done[0]); done[1] = fadd2(acc[1], done[1]); done[2] = fadd2(acc[2], done[2]); done[3] = fadd2(acc[3], done[3]);
      }
      if (V == 2 || V == 4) {
        const float kf = __uint_as_float((uint32_t)kp), cf = __uint_as_float((uint32_t)cp);
#pragma unroll
        for (int i = 0; i < 4; i++) {
          uint64_t v = i == 0 ? v0 : i == 1 ? v1 : i == 2 ? v2 : v3;
          float al = __uint_as_float((uint32_t)acc[i]), ah = __uint_as_float((uint32_t)(acc[i] >> 32));
          float dl = __uint_as_float((uint32_t)done[i]), dh = __uint_as_float((uint32_t)(done[i] >> 32));
          al = fmaf(al, kf, __uint_as_float((uint32_t)v)); ah = fmaf(ah, kf, __uint_as_float((uint32_t)(v >> 32)));
          dl = fmaf(al, cf, dl); dh = fmaf(ah, cf, dh);
          acc[i] = pack2(al, ah); done[i] = pack2(dl, dh);
        }
      }
    }
    w = w * 1664525u + 1013904223u;
  }
  long long t1 = clock64();
  uint64_t r = acc[0] ^ acc[1] ^ acc[2] ^ acc[3] ^ done[0] ^ done[1] ^ done[2] ^ done[3];
  out[blockIdx.x * 512 + threadIdx.x] = __uint_as_float((uint32_t)r ^ (uint32_t)(r >> 32));
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
template <int V> void run(const uint32_t* codes, float* out, long long* clk, int iters) {
  cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
  k<V><<<148, 512, 131072>>>(codes, out, iters, clk);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<V><<<148, 512, 131072>>>(codes, out, iters, clk);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long c[148]; cudaMemcpy(c, clk, sizeof c, cudaMemcpyDeviceToHost);
  double steps = (double)iters * 8;   // per warp
  printf("V=%d: %.3f ms, %.2f clk per warp-step-of-4-LDS.64 per SM (16 warps => %.2f clk per LDS.64 warp-instr), err=%s\n", V, ms,
         c[0] / steps / 16.0, c[0] / steps / 16.0 / 4.0, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  uint32_t* codes; float* out; long long* clk;
  cudaMalloc(&codes, 148 * 512 * 4); cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&clk, 148 * 8);
  uint32_t* h = new uint32_t[148 * 512];
  for (int i = 0; i < 148 * 512; i++) h[i] = i * 2654435761u;
  cudaMemcpy(codes, h, 148 * 512 * 4, cudaMemcpyHostToDevice);
  const int iters = 20000;
  run<0>(codes, out, clk, iters); run<1>(codes, out, clk, iters); run<2>(codes, out, clk, iters);
  run<3>(codes, out, clk, iters); run<4>(codes, out, clk, iters); run<5>(codes, out, clk, iters);
  run<6>(codes, out, clk, iters); run<7>(codes, out, clk, iters); run<8>(codes, out, clk, iters);
  return 0;
}
