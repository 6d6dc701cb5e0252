#include <cstdint>
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__global__ void k(const uint64_t* in, uint64_t* out, const float* lut) {
  extern __shared__ uint64_t s[];
  uint64_t a = in[threadIdx.x], b = in[threadIdx.x + 32], c = in[threadIdx.x + 64];
  uint64_t v = s[threadIdx.x * 7 & 255];
  a = ffma2(a, b, v);
  c = ffma2(a, b, c);
  out[threadIdx.x] = fadd2(a, c);
}
