// icm.cu -- path (1): LSQ / LSQ++ ICM-ILS encoding on B200.
//   K0 sqnorm_kernel     ||c||^2 of every codebook entry
//   K1 unary_kernel      U[l][j][c] = -2<C_j[:,c], x_l> + ||C_j[:,c]||^2     (src/utils.jl:121-149)
//   K2 tables_kernel     T[j][k][b][c] = 2<C_j[:,c], C_k[:,b]>, both orientations (src/utils.jl:152-171,
//                        src/LSQ.jl:180-183), so every conditioning row is 1 KB contiguous
//   K3 icm_warp_kernel   the whole ILS loop of encode_icm_fully! (src/LSQ.jl:199-249) fused on device:
//                        perturb -> icmiter x m conditioning steps (deps/src/encode_icm.cpp:26-59) -> cost ->
//                        strict-< accept; one warp per vector
//   veccost_kernel       src/qerrors.jl:36-66
//   condition_kernel     exact-signature compat for deps/src/encode_icm.cpp:157-168
// Arithmetic orders are the oracle's (DESIGN.md "Arithmetic contract"): sequential-t fmaf chains for the
// dot products, ascending-k fp32 adds for the conditioning, sequential unfused sums for the cost.
#include <algorithm>
#include <cstring>

#include "common.cuh"

namespace ryl {

static constexpr int kH = 256;

// ---- K0 ---------------------------------------------------------------------------------------------
__global__ void sqnorm_kernel(const float* __restrict__ C, int d, int mh, float* __restrict__ nrm) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= mh) return;
  const float* c = C + (size_t)e * d;
  float s = 0.f;
  for (int t = 0; t < d; t++) s = fmaf(c[t], c[t], s);
  nrm[e] = s;
}

// ---- K1: fp32 SIMT GEMM, 128 entries x 128 vectors per block, 8x8 per thread, BK = 8 --------------------
// Each output is ONE sequential-t fmaf chain (no split-K), so it is bit-identical to the oracle's dot_seq.
// umax (optional): umax[l] = max over all entries of |U[l][.]|, as float bits via atomicMax (|.| >= 0 orders like
// an unsigned integer); it bounds the fp32 rounding slack of the quantised pre-filter in K3.
template <bool VEC4>
__global__ void __launch_bounds__(256) unary_kernel(const float* __restrict__ C, const float* __restrict__ X,
                                                    const float* __restrict__ nrm, float* __restrict__ U, int64_t n,
                                                    int d, int mh, unsigned int* __restrict__ umax) {
  constexpr int BM = 128, BN = 128, BK = 8;
  __shared__ __align__(16) float As[BK][BM];
  __shared__ __align__(16) float Bs[BK][BN];
  __shared__ float Rs[16][BN];                       // per-vector |U| maxima of the 16 entry groups (umax epilogue)
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int e0 = blockIdx.x * BM;
  const int64_t l0 = (int64_t)blockIdx.y * BN;
  const int lr = tid >> 1, lk = (tid & 1) * 4;  // loader: row, k-quad
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[i][j] = 0.f;

  // software pipeline: the global loads of tile k0+BK are issued before the FMAs of tile k0 and land in registers
  // while those run; each output is still ONE sequential-t fmaf chain
  const float* arow = C + (size_t)(e0 + lr) * d + lk;
  const int64_t lv = l0 + lr;
  const float* brow = X + (size_t)(lv < n ? lv : n - 1) * d + lk;
  float a[4], b[4];
  auto fetch = [&](int k0) {
    if (VEC4) {
      float4 av = make_float4(0.f, 0.f, 0.f, 0.f), bv = av;
      if (k0 + lk < d) {  // d % 4 == 0: a quad is entirely inside or entirely outside
        av = *reinterpret_cast<const float4*>(arow + k0);
        bv = *reinterpret_cast<const float4*>(brow + k0);
      }
      a[0] = av.x; a[1] = av.y; a[2] = av.z; a[3] = av.w;
      b[0] = bv.x; b[1] = bv.y; b[2] = bv.z; b[3] = bv.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; i++) {
        bool ok = k0 + lk + i < d;
        a[i] = ok ? arow[k0 + i] : 0.f;
        b[i] = ok ? brow[k0 + i] : 0.f;
      }
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < d; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      As[lk + i][lr] = a[i];
      Bs[lk + i][lr] = b[i];
    }
    __syncthreads();
    if (k0 + BK < d) fetch(k0 + BK);
    auto kstep = [&](int kk) {
      float av[8], bv[8];
      *reinterpret_cast<float4*>(&av[0]) = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      *reinterpret_cast<float4*>(&av[4]) = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      // vectors of thread tx: 4*tx .. 4*tx+3 and 64 + 4*tx .. 64 + 4*tx+3 -> consecutive lanes read consecutive
      // 16 B chunks (no bank conflicts; 8 consecutive floats per lane collide 2-way)
      *reinterpret_cast<float4*>(&bv[0]) = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      *reinterpret_cast<float4*>(&bv[4]) = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    };
    if (k0 + BK <= d) {                  // full tile: unrolled, so the shared loads run ahead of the FMAs
#pragma unroll
      for (int kk = 0; kk < BK; kk++) kstep(kk);
    } else {                             // the chain is exactly d long, like the oracle's dot_seq
      for (int kk = 0; kk < d - k0; kk++) kstep(kk);
    }
    __syncthreads();
  }
  float nr[8];
#pragma unroll
  for (int i = 0; i < 8; i++) nr[i] = nrm[e0 + ty * 8 + i];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int vt = (j < 4 ? 0 : 60) + tx * 4 + j;               // vector within the tile (see the Bs reads)
    int64_t l = l0 + vt;
    float mx = 0.f;
    if (l < n) {
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; i++) o[i] = fmaf(-2.0f, acc[i][j], nr[i]);  // -2*dot exact, one rounding
      float* dst = U + (size_t)l * mh + e0 + ty * 8;
      *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
      bool bad = false;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        mx = fmaxf(mx, fabsf(o[i]));
        bad |= o[i] != o[i];
      }
      if (bad) mx = __int_as_float(0x7f800000);                 // NaN unaries -> +inf slack -> exact path
    }
    if (umax) Rs[ty][vt] = mx;
  }
  if (umax) {   // one atomic per vector per block: max over the block's 128 entries
    __syncthreads();
    if (tid < BN) {
      float mx = 0.f;
#pragma unroll
      for (int t = 0; t < 16; t++) mx = fmaxf(mx, Rs[t][tid]);
      if (l0 + tid < n) atomicMax(umax + l0 + tid, __float_as_uint(mx));
    }
  }
}

// K1, d % 16 == 0: same outputs bit for bit (every output is still ONE fmaf chain over ascending t), restructured for
// the FMA pipe: BK = 16 with two shared-memory stages (one block barrier per 16 steps instead of two per 8), and the
// 8x8 register tile updated by 32 packed fma.rn.f32x2 per step (FFMA2: two independent chains per instruction, exact
// per element) instead of 64 FFMA -- the issue slots that frees go to the shared-memory loads.
__device__ __forceinline__ unsigned long long k1_fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long k1_dup(float a) {
  unsigned long long d;
  asm("mov.b64 %0, {%1, %1};" : "=l"(d) : "f"(a));
  return d;
}
__device__ __forceinline__ unsigned long long k1_pack(float lo, float hi) {
  unsigned long long d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
  return d;
}
__global__ void __launch_bounds__(256, 2) unary_kernel2(const float* __restrict__ C, const float* __restrict__ X,
                                                        const float* __restrict__ nrm, float* __restrict__ U, int64_t n,
                                                        int d, int mh, unsigned int* __restrict__ umax) {
  constexpr int BM = 128, BN = 128, BK = 16, LD = 132;
  __shared__ __align__(16) float As[2][BK][LD];
  __shared__ __align__(16) float Bs[2][BK][LD];
  __shared__ float Rs[16][BN];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int e0 = blockIdx.x * BM;
  const int64_t l0 = (int64_t)blockIdx.y * BN;
  const int lr = tid >> 1, lk = (tid & 1) * 8;       // loader: row, first of its 8 k
  unsigned long long acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0ull;    // (+0, +0)
  const float* arow = C + (size_t)(e0 + lr) * d + lk;
  const int64_t lv = l0 + lr;
  const float* brow = X + (size_t)(lv < n ? lv : n - 1) * d + lk;
  float4 ga[2], gb[2];
  auto fetch = [&](int k0) {
    ga[0] = *reinterpret_cast<const float4*>(arow + k0);
    ga[1] = *reinterpret_cast<const float4*>(arow + k0 + 4);
    gb[0] = *reinterpret_cast<const float4*>(brow + k0);
    gb[1] = *reinterpret_cast<const float4*>(brow + k0 + 4);
  };
  auto stage = [&](int buf) {
    const float av[8] = {ga[0].x, ga[0].y, ga[0].z, ga[0].w, ga[1].x, ga[1].y, ga[1].z, ga[1].w};
    const float bv[8] = {gb[0].x, gb[0].y, gb[0].z, gb[0].w, gb[1].x, gb[1].y, gb[1].z, gb[1].w};
#pragma unroll
    for (int i = 0; i < 8; i++) {
      As[buf][lk + i][lr] = av[i];
      Bs[buf][lk + i][lr] = bv[i];
    }
  };
  fetch(0);
  stage(0);
  __syncthreads();
  const int nt = d / BK;
  for (int kt = 0; kt < nt; kt++) {
    const int buf = kt & 1;
    if (kt + 1 < nt) fetch((kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; kk++) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      const unsigned long long bp[4] = {k1_pack(b0.x, b0.y), k1_pack(b0.z, b0.w), k1_pack(b1.x, b1.y), k1_pack(b1.z, b1.w)};
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const unsigned long long ad = k1_dup(av[i]);
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = k1_fma2(ad, bp[j], acc[i][j]);
      }
    }
    if (kt + 1 < nt) stage(buf ^ 1);
    __syncthreads();
  }
  float nr[8];
#pragma unroll
  for (int i = 0; i < 8; i++) nr[i] = nrm[e0 + ty * 8 + i];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int vt = (j < 4 ? 0 : 60) + tx * 4 + j;               // vector within the tile (see the Bs reads)
    int64_t l = l0 + vt;
    float mx = 0.f;
    if (l < n) {
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const unsigned long long pr = acc[i][j >> 1];
        const float a = __uint_as_float((j & 1) ? (unsigned)(pr >> 32) : (unsigned)pr);
        o[i] = fmaf(-2.0f, a, nr[i]);                            // -2*dot exact, one rounding
      }
      float* dst = U + (size_t)l * mh + e0 + ty * 8;
      *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
      bool bad = false;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        mx = fmaxf(mx, fabsf(o[i]));
        bad |= o[i] != o[i];
      }
      if (bad) mx = __int_as_float(0x7f800000);                 // NaN unaries -> +inf slack -> exact path
    }
    if (umax) Rs[ty][vt] = mx;
  }
  if (umax) {   // one atomic per vector per block: max over the block's 128 entries
    __syncthreads();
    if (tid < BN) {
      float mx = 0.f;
#pragma unroll
      for (int t = 0; t < 16; t++) mx = fmaxf(mx, Rs[t][tid]);
      if (l0 + tid < n) atomicMax(umax + l0 + tid, __float_as_uint(mx));
    }
  }
}

// ---- K2: pairwise tables, both orientations ------------------------------------------------------------
// T[((j*m + k)*256 + b)*256 + c] = 2 * <C_j[:,c], C_k[:,b]>; the product is commutative inside fmaf, so
// T[j][k][b][c] == T[k][j][c][b] bit for bit, i.e. this is binaries / binaries_t of the reference.
__global__ void __launch_bounds__(256) tables_kernel(const float* __restrict__ C, float* __restrict__ T, int d, int m) {
  constexpr int CH = 64;
  __shared__ float cs[32][CH + 1];
  __shared__ float bs[32][CH];
  const int j = blockIdx.z / m, k = blockIdx.z % m;
  if (j == k) return;
  const int c0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
  const int c = threadIdx.x & 31, bg = threadIdx.x >> 5;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int base = 0; base < d; base += CH) {
    const int chunk = min(CH, d - base);
    for (int i = threadIdx.x; i < 32 * CH; i += 256) {
      int r = i / CH, t = i % CH;
      if (t < chunk) {
        cs[r][t] = C[((size_t)j * kH + c0 + r) * d + base + t];
        bs[r][t] = C[((size_t)k * kH + b0 + r) * d + base + t];
      }
    }
    __syncthreads();
    for (int t = 0; t < chunk; t++) {
      float cv = cs[c][t];
#pragma unroll
      for (int i = 0; i < 4; i++) acc[i] = fmaf(cv, bs[bg + 8 * i][t], acc[i]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int b = b0 + bg + 8 * i;
    T[(((size_t)j * m + k) * kH + b) * kH + c0 + c] = 2.0f * acc[i];
  }
}

// ---- K2q: quantised copy of the tables for the K3 pre-filter ------------------------------------------------------
// tmax[j] = max |T[j][k][.][.]| over k != j;  scale_j = tmax[j] / Q;  Q = 8191 (14-bit fields; 32767 with RYL_K3_QBITS = 16)
// Tq[j][k][b][pos] = rint(T[j][k][b][c] / scale_j) + Q + 1 (1..2Q+1, two 16-bit containers per word), the row permuted so
// that the 16 bytes lane L loads are the 32-bit words w = 0..3 = { lo: c = 4L + w, hi: c = 128 + 4L + w } -- the same
// candidates the lane owns in the exact path.  |scale_j*(q - Q - 1) - T| <= 0.51*scale_j (rint + one fp32 division rounding).
#ifndef RYL_K3_QBITS
#define RYL_K3_QBITS 14
#endif
// 14-bit fields q + 8192 in 1..16383: up to four rows add inside their 16-bit containers without a carry, so the rows of a
// step are summed TWO candidates per integer add (pf_rows14 / the uniform loop of m > 8); the unit is 4x coarser than
// with 16-bit fields (more near-ties: ~5 % instead of ~1.5 % of the steps, resolved on the window's candidates) but a
// step has 32 instructions less.  RYL_K3_UQS16 = 0 keeps the round-2 kernel for m > 8 (16-bit fields, fp32 unaries).
#ifndef RYL_K3_UQS16
#define RYL_K3_UQS16 1
#endif
__host__ __device__ inline float pf_q(int m) {
  return ((m <= 8 || RYL_K3_UQS16) && RYL_K3_QBITS == 14) ? 8191.0f : 32767.0f;
}
__global__ void __launch_bounds__(256) tmax_kernel(const float* __restrict__ T, int m, unsigned int* __restrict__ tmax) {
  const int j = blockIdx.y / m, k = blockIdx.y % m;
  if (j == k) return;
  const float4* t = reinterpret_cast<const float4*>(T + ((size_t)j * m + k) * kH * kH);
  float mx = 0.f;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < kH * kH / 4; i += gridDim.x * 256) {
    const float4 v = t[i];
    mx = fmaxf(fmaxf(fmaxf(mx, fabsf(v.x)), fmaxf(fabsf(v.y), fabsf(v.z))), fabsf(v.w));
    if (v.x != v.x || v.y != v.y || v.z != v.z || v.w != v.w) mx = __int_as_float(0x7f800000);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  if ((threadIdx.x & 31) == 0) atomicMax(tmax + j, __float_as_uint(mx));
}

__global__ void __launch_bounds__(256) quant_tables_kernel(const float* __restrict__ T, int m,
                                                           const unsigned int* __restrict__ tmax,
                                                           uint32_t* __restrict__ Tq) {
  const int j = blockIdx.z / m, k = blockIdx.z % m;
  if (j == k) {   // the diagonal blocks are all-zero words: a row that adds nothing (K3's uniform row loop reads it)
    for (int i = blockIdx.x * 256 + threadIdx.x; i < kH * (kH / 2); i += gridDim.x * 256)
      Tq[(((size_t)j * m + k) * kH) * (kH / 2) + i] = 0u;
    return;
  }
  const float tm = __uint_as_float(tmax[j]);
  const float kQ = pf_q(m);
  const int off = (int)kQ + 1;
  const float scale = (tm > 0.f && tm < __int_as_float(0x7f800000)) ? __fdiv_rn(tm, kQ) : 1.0f;
  // one thread per 32-bit output word: row b = blockIdx.x*? ...  grid.x * 256 threads cover 256 rows * 128 words
  for (int i = blockIdx.x * 256 + threadIdx.x; i < kH * (kH / 2); i += gridDim.x * 256) {
    const int b = i >> 7, word = i & 127, L = word >> 2, w = word & 3;
    const float* row = T + (((size_t)j * m + k) * kH + b) * kH;
    const float qlo = rintf(__fdiv_rn(row[4 * L + w], scale)), qhi = rintf(__fdiv_rn(row[128 + 4 * L + w], scale));
    const uint32_t lo = (uint32_t)((int)fminf(fmaxf(qlo, -kQ), kQ) + off);
    const uint32_t hi = (uint32_t)((int)fminf(fmaxf(qhi, -kQ), kQ) + off);
    Tq[(((size_t)j * m + k) * kH + b) * (kH / 2) + word] = lo | (hi << 16);
  }
}

// {1/scale_j, W0_j}: W0_j = 2.002 * ((m-1)*0.51 + 0.75 + 2^-20*(m-1)*32767) is the part of the pre-filter window
// (in units of scale_j) that does not depend on the vector; degenerate tables (all zero, inf, NaN) get W0 = +inf,
// which sends every step of that codebook down the exact path.
__global__ void pf_consts_kernel(const unsigned int* __restrict__ tmax, int m, float2* __restrict__ pfc) {
  const int j = threadIdx.x;
  if (j >= m) return;
  const float tm = __uint_as_float(tmax[j]);
  const bool ok = m > 1 && tm > 0.f && tm < __int_as_float(0x7f800000);
  const float kQ = pf_q(m);
  const float w0 = 2.002f * ((float)(m - 1) * 0.51f + 0.75f + 9.5367431640625e-07f * (float)(m - 1) * kQ);
  pfc[j] = ok ? make_float2(__fdiv_rn(kQ, tm), w0) : make_float2(0.f, __int_as_float(0x7f800000));
}

// ---- helpers for codes held in two 64-bit registers (m <= 16) -------------------------------------------
struct Code {
  uint64_t lo, hi;
  __device__ __forceinline__ uint32_t get(int k) const {
    return (uint32_t)((k < 8 ? lo >> (8 * k) : hi >> (8 * (k - 8))) & 0xFFu);
  }
  __device__ __forceinline__ void set(int k, uint32_t v) {
    if (k < 8) lo = (lo & ~(0xFFull << (8 * k))) | ((uint64_t)v << (8 * k));
    else hi = (hi & ~(0xFFull << (8 * (k - 8)))) | ((uint64_t)v << (8 * (k - 8)));
  }
};

// M <= 8: the codes live in `lo` alone, so the k < 8 selects disappear from the per-step code path
template <int M>
__device__ __forceinline__ uint32_t code_get(const Code& c, int k) {
  if (M <= 8) return (uint32_t)(c.lo >> (8 * k)) & 0xFFu;
  return c.get(k);
}
template <int M>
__device__ __forceinline__ void code_set(Code& c, int k, uint32_t v) {
  if (M <= 8) c.lo = (c.lo & ~(0xFFull << (8 * k))) | ((uint64_t)v << (8 * k));
  else c.set(k, v);
}

template <int M>
__device__ __forceinline__ Code load_code(const uint8_t* b) {
  Code c{0, 0};
#pragma unroll
  for (int k = 0; k < M; k++) c.set(k, b[k]);
  return c;
}

// perturb_codes! (src/LSQ.jl:5-39, with replacement); identical draws to the oracle (DESIGN.md "RNG"): Philox4x32-10
// keyed by the seed, counter (g_lo, g_hi, it, block) for the positions (mulhi(w, m)) and block | 0x80000000 for the
// values (mulhi(w, 256)), applied in draw order so a later draw on the same position wins.
// Precomputed: one thread per (vector, ILS iteration, Philox block of 4 draws), packed as 4 positions
// (4 bits each, bits 0..15) + 4 values (8 bits each, bits 16..47).  K3 used to run both Philox blocks redundantly in all
// 32 lanes of the warp at the top of every ILS iteration (~240 instructions, 4 % of the kernel); now it loads 8 bytes.
__global__ void __launch_bounds__(256) perturb_draws_kernel(unsigned long long* __restrict__ P, int64_t nc, int64_t g0,
                                                            int ilsiter, int nblk, int m, uint64_t seed) {
  const int64_t total = nc * ilsiter * nblk;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int jb = (int)(i % nblk), it = (int)((i / nblk) % ilsiter);
    const uint64_t g = (uint64_t)(g0 + i / ((int64_t)nblk * ilsiter));
    uint32_t pw[4] = {(uint32_t)g, (uint32_t)(g >> 32), (uint32_t)it, (uint32_t)jb};
    uint32_t vw[4] = {(uint32_t)g, (uint32_t)(g >> 32), (uint32_t)it, 0x80000000u | (uint32_t)jb};
    philox4x32_10(pw, (uint32_t)seed, (uint32_t)(seed >> 32));
    philox4x32_10(vw, (uint32_t)seed, (uint32_t)(seed >> 32));
    unsigned long long pk = 0;
#pragma unroll
    for (int t = 0; t < 4; t++)
      pk |= (unsigned long long)mulhi32(pw[t], (uint32_t)m) << (4 * t) |
            (unsigned long long)mulhi32(vw[t], (uint32_t)kH) << (16 + 8 * t);
    P[i] = pk;
  }
}

// veccost for one vector, cooperatively by one warp; result uniform across the warp.
// cb = ((0 + C_0[t,b_0]) + C_1[t,b_1]) + ... ; cost = sequential sum over t of (cb - x[t])^2, unfused.
// the reference's sequential sum of the d squares in sq[] (every lane computes the same chain; 16-byte broadcast loads)
__device__ __forceinline__ float cost_seq(const float* sq, int d) {
  __syncwarp();
  float acc = 0.f;
  const int d4 = d & ~3;
  for (int t = 0; t < d4; t += 4) {
    const float4 v = *reinterpret_cast<const float4*>(sq + t);
    acc = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(acc, v.x), v.y), v.z), v.w);
  }
  for (int t = d4; t < d; t++) acc = __fadd_rn(acc, sq[t]);
  __syncwarp();
  return acc;
}
__device__ __noinline__ float cost_seq_call(const float* sq, int d) { return cost_seq(sq, d); }

// `reject_above`: the caller only needs to know whether the cost is < or == that value (the ILS accept test, strict <).
// The squares are first added as a tree (own values, then a 5-step butterfly: depth 8); both the tree sum and the
// reference's sequential sum of the same n = d non-negative terms are within gamma_{n-1} resp. gamma_8 of the real sum, so
// seq >= tree * (1 - gamma_{d-1}) / (1 + gamma_8): when tree > reject_above * (1 + 1.3e-7 * (d + 32)) (more than twice that
// bound) the sequential sum is strictly above reject_above too and its 128 dependent adds are skipped -- the value returned
// is then only known to compare greater.  NaN / inf on either side fail the test and take the exact chain.
template <int M>
__device__ __forceinline__ float warp_cost(const float* __restrict__ x, const float* __restrict__ C, const Code code,
                                           int d, float* sq, int lane, const float reject_above,
                                           const bool tree_only = false) {
  float part = 0.f;
  if ((d & 3) == 0) {   // four consecutive t per lane: one 16-byte load per codeword row (same arithmetic per element)
    for (int t4 = lane * 4; t4 < d; t4 += 128) {
      float4 cb = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < M; k++) {
        const float4 c = __ldg(reinterpret_cast<const float4*>(C + ((size_t)k * kH + code.get(k)) * d + t4));
        cb.x = __fadd_rn(cb.x, c.x); cb.y = __fadd_rn(cb.y, c.y); cb.z = __fadd_rn(cb.z, c.z); cb.w = __fadd_rn(cb.w, c.w);
      }
      const float4 xv = __ldg(reinterpret_cast<const float4*>(x + t4));
      const float dx = __fsub_rn(cb.x, xv.x), dy = __fsub_rn(cb.y, xv.y), dz = __fsub_rn(cb.z, xv.z), dw = __fsub_rn(cb.w, xv.w);
      const float4 o = make_float4(__fmul_rn(dx, dx), __fmul_rn(dy, dy), __fmul_rn(dz, dz), __fmul_rn(dw, dw));
      *reinterpret_cast<float4*>(sq + t4) = o;
      part = __fadd_rn(part, __fadd_rn(__fadd_rn(o.x, o.y), __fadd_rn(o.z, o.w)));
    }
  } else {
    for (int t = lane; t < d; t += 32) {
      float cb = 0.f;
#pragma unroll
      for (int k = 0; k < M; k++) cb = __fadd_rn(cb, __ldg(C + ((size_t)k * kH + code.get(k)) * d + t));
      float df = __fsub_rn(cb, __ldg(x + t));
      const float o = __fmul_rn(df, df);
      sq[t] = o;
      part = __fadd_rn(part, o);
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) part = __fadd_rn(part, __shfl_xor_sync(0xffffffffu, part, off));
  // (the per-lane chain over chunks / strided t is at most d / 32 deep: covered by the d-proportional margin)
  if (tree_only || part > reject_above * (1.0f + 1.3e-7f * (float)(d + 32))) return part;   // strictly above: see the header
  return cost_seq(sq, d);
}

// one out-of-line copy for K3 (called at the start of a vector and once per ILS iteration): the kernel's hot instruction
// footprint (8 specialised row loops + tail) sits at the instruction cache's capacity, every inlined copy costs throughput
#ifndef RYL_K3_COSTCALL
#define RYL_K3_COSTCALL 1
#endif
template <int M>
__device__ __noinline__ float warp_cost_call(const float* __restrict__ x, const float* __restrict__ C, uint64_t lo,
                                             uint64_t hi, int d, float* sq, int lane, float reject_above) {
  return warp_cost<M>(x, C, Code{lo, hi}, d, sq, lane, reject_above);
}

// L2 eviction-priority hints: with m = 16 the in-flight vectors' unaries (32 warps x 148 SMs x 16 KB = 78 MB) plus the
// quantised tables (34 MB) only just fit the 126 MB L2, and plain LRU lets the streaming fp32 rows of the rare exact
// steps push unary rows out (ncu: 307 GB of DRAM reads per 125k vectors).  Unaries are loaded evict_last, exact
// fp32 rows evict_first.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ float4 ldg_f4_hint(const float4* ptr, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(ptr), "l"(pol));
  return v;
}

struct IcmParams {
  const float* U;       // [nc][m][256]  unaries of this chunk
  const float* T;       // [m][m][256][256]
  const uint32_t* Tq;   // [m][m][256][128] quantised pairs (K2q), or null: exact path only
  const float2* pfc;    // [m] {1/scale_j, W0_j} (pf_consts_kernel)
  const unsigned int* umax;  // [nc] float bits of max |U[l][.]|
  const float* X;       // [nc][d]
  const float* C;       // [m*256][d]
  uint8_t* B;           // [nc][m] in/out
  float* cost;          // [nc] or null
  const int* orders;    // [ilsiter][m]
  const unsigned long long* orders_packed;  // [ilsiter]: the m visiting-order entries of an iteration, 4 bits each
  uint32_t one;         // == 1, opaque to the compiler (see pf_rows)
  const unsigned long long* draws;  // [nc][ilsiter][ceil(npert/4)] packed perturbation draws (perturb_draws_kernel)
  const int* snap_iters;  // [n_snap] (1-based ILS iteration counts)
  uint8_t* B_snap;      // [n_snap][n_total][m], already offset to this chunk's first vector
  int* stats;           // [ilsiter][2] (#equal, #better)
  unsigned long long* steps;  // [2] conditioning steps actually executed (memoisation skips the rest); of those,
                              //     steps the quantised pre-filter could not decide (exact path taken)
  unsigned long long* next;   // work counter of this launch (zeroed by the host): warps draw vectors from it
  int64_t nc, n_total, g0;
  uint64_t seed;
  int d, ilsiter, icmiter, npert, n_snap;
};

// ---- K3: one warp per vector, all ILS iterations of that vector back to back ----------------------------
// Lane owns candidates c = r*128 + lane*4 + e (r<2, e<4): every row (unary or pairwise) is one (quantised, 512 B) or two
// (fp32, 1 KB) coalesced 16-byte loads per lane.  The m*m*128 KB of quantised tables live in L2; the vector's unaries are
// read once and kept by the warp in shared memory as 16-bit integers (see the kernel); the fp32 tables and unaries are
// touched only by near-ties.  DESIGN.md 4 / 4d has the history and the measurements behind each choice.
//
// pf_rows: the row loop of the 16-bit-field build (RYL_K3_QBITS = 16; the shipped build uses pf_rows14 below).  The
// (M-1) quantised rows of one step with the conditioned codebook J as a LITERAL: no per-row predicate, no predicated-off
// row.  sl = sum lo + (sum hi << 16) (mod 2^32), sh = sum hi.  `one` is a kernel parameter equal to 1 that the compiler
// cannot fold, so x * one + s is emitted as IMAD (FMA pipe) instead of IADD3 (that build was bound by the integer ALU pipe).
template <int M, int J>
__device__ __forceinline__ void pf_rows(const char* tqj, const Code& nb, uint32_t (&sl)[4], uint32_t (&sh)[4],
                                        const uint32_t one) {
#pragma unroll
  for (int k = 0; k < M; k++) {
    if (k != J) {
      const uint32_t word = (uint32_t)((k < 8 ? nb.lo : nb.hi) >> (32 * ((k & 7) >> 2)));
      const uint32_t code = __byte_perm(word, 0, 0x4440 | (k & 3));
      const uint4 x = __ldg(reinterpret_cast<const uint4*>(tqj + (size_t)k * (kH * 512) + (size_t)code * 512));
      sl[0] = x.x * one + sl[0]; sl[1] = x.y * one + sl[1]; sl[2] = x.z * one + sl[2]; sl[3] = x.w * one + sl[3];
      sh[0] += x.x >> 16; sh[1] += x.y >> 16; sh[2] += x.z >> 16; sh[3] += x.w >> 16;
    }
  }
}

// 14-bit fields (m <= 8): the rows of codebooks k < 4 and k >= 4 are summed as whole 32-bit words into A and B -- at most
// four rows each, 4 * 16383 < 2^16, so neither 16-bit field carries -- and only the three words A, B and the unary are
// split into their halves: 8 instead of 16 instructions per word and step.
// compile-time variants kept for A/B builds (tools/build_variant.sh); the defaults are the measured optima
#ifndef RYL_K3_LITE
#define RYL_K3_LITE 1        // near-ties decided on the window's candidates
#endif
#ifndef RYL_K3_COLD
#define RYL_K3_COLD 1        // whole-row exact step out of line
#endif
#ifndef RYL_K3_REJECT
#define RYL_K3_REJECT 1      // certified early rejection in the cost evaluation
#endif
#ifndef RYL_K3_HALF
#define RYL_K3_HALF 0      // two register-indexed copies of the row loop instead of M literal ones
#endif
#ifndef RYL_K3_DIAG0ROW
#define RYL_K3_DIAG0ROW 0    // uniform row loop: one L1-resident zero row instead of the diagonal block's
#endif
// (adds on the FMA pipe -- x * one + acc -- measured slower here than IADD3, which fuses two adds: 194 vs 187 ms)
__device__ __forceinline__ uint32_t pf_add(uint32_t acc, uint32_t x, const uint32_t, int) { return acc + x; }
template <int M, int J>
__device__ __forceinline__ void pf_rows14(const char* tqj, const Code& nb, const uint4 xu, int (&S)[8],
                                          const uint32_t one) {
  uint32_t A[4] = {0, 0, 0, 0}, Bq[4] = {0, 0, 0, 0};
#pragma unroll
  for (int k = 0; k < M; k++) {
    if (k != J) {
      const uint32_t word = (uint32_t)((k < 8 ? nb.lo : nb.hi) >> (32 * ((k & 7) >> 2)));
      const uint32_t code = __byte_perm(word, 0, 0x4440 | (k & 3));
      const uint4 x = __ldg(reinterpret_cast<const uint4*>(tqj + (size_t)k * (kH * 512) + (size_t)code * 512));
      const bool first = (k == 0) || (k == 1 && J == 0) || (k == 4) || (k == 5 && J == 4);
      uint32_t (&G)[4] = k < 4 ? A : Bq;
      if (first) { G[0] = x.x; G[1] = x.y; G[2] = x.z; G[3] = x.w; }
      else {
        G[0] = pf_add(G[0], x.x, one, k); G[1] = pf_add(G[1], x.y, one, k);
        G[2] = pf_add(G[2], x.z, one, k); G[3] = pf_add(G[3], x.w, one, k);
      }
    }
  }
  const uint32_t xw[4] = {xu.x, xu.y, xu.z, xu.w};
#pragma unroll
  for (int w = 0; w < 4; w++) {
    uint32_t h = (xw[w] >> 16) + (A[w] >> 16);
    uint32_t t = pf_add(A[w], xw[w], one, 1);
    if (M > 4) { h += Bq[w] >> 16; t = pf_add(t, Bq[w], one, 0); }
    S[w] = (int)(t - (h << 16));
    S[4 + w] = (int)h;
  }
}

// Two copies of the row loop instead of M (RYL_K3_HALF, 5 <= M <= 8): the conditioned codebook j only selects WHICH
// three (or M-5) rows of its own group of four are read -- those rows take their codebook index from a register
// (k = (j + r) & 3 resp. 4 + (j - 4 + r) % (M - 4)), the other group's rows are literals.  +2 instructions per
// register-indexed row, but a quarter of the code of the per-j copies and a two-way branch instead of a jump table.
template <int M, bool LOW>
__device__ __forceinline__ void pf_rows14_half(const char* tqj, const Code& nb, const uint4 xu, int (&S)[8], const int j) {
  constexpr int GB = M - 4;                                 // rows of the second group
  uint32_t A[4] = {0, 0, 0, 0}, Bq[4] = {0, 0, 0, 0};
  auto row_lit = [&](const int k, uint32_t (&G)[4], const bool first) {
    const uint32_t word = (uint32_t)(nb.lo >> (32 * (k >> 2)));
    const uint32_t code = __byte_perm(word, 0, 0x4440 | (k & 3));
    const uint4 x = __ldg(reinterpret_cast<const uint4*>(tqj + (size_t)k * (kH * 512) + (size_t)code * 512));
    if (first) { G[0] = x.x; G[1] = x.y; G[2] = x.z; G[3] = x.w; }
    else { G[0] += x.x; G[1] += x.y; G[2] += x.z; G[3] += x.w; }
  };
  auto row_reg = [&](const int k, uint32_t (&G)[4], const bool first) {
    const uint32_t code = (uint32_t)(nb.lo >> (8 * k)) & 255u;
    const uint4 x = __ldg(reinterpret_cast<const uint4*>(tqj + (size_t)(k * kH + code) * 512));
    if (first) { G[0] = x.x; G[1] = x.y; G[2] = x.z; G[3] = x.w; }
    else { G[0] += x.x; G[1] += x.y; G[2] += x.z; G[3] += x.w; }
  };
  if (LOW) {                                                // j in 0..3
#pragma unroll
    for (int r = 1; r < 4; r++) row_reg((j + r) & 3, A, r == 1);
#pragma unroll
    for (int k = 4; k < M; k++) row_lit(k, Bq, k == 4);
  } else {                                                  // j in 4..M-1
#pragma unroll
    for (int k = 0; k < 4; k++) row_lit(k, A, k == 0);
#pragma unroll
    for (int r = 1; r < GB; r++) {
      int kk = j - 4 + r;
      if (kk >= GB) kk -= GB;
      row_reg(4 + kk, Bq, r == 1);
    }
  }
  const uint32_t xw[4] = {xu.x, xu.y, xu.z, xu.w};
#pragma unroll
  for (int w = 0; w < 4; w++) {
    const uint32_t h = (xw[w] >> 16) + (A[w] >> 16) + (Bq[w] >> 16);
    const uint32_t t = A[w] + xw[w] + Bq[w];
    S[w] = (int)(t - (h << 16));
    S[4 + w] = (int)h;
  }
}

// The exact step (fp32 rows in ascending k, first-minimum argmin -- encode_icm.cpp:28-58) out of line: with the windowed
// evaluation of near-ties it runs for ~0.1 % of the steps of the m <= 8 kernels, so it stays out of their instruction
// footprint and register allocation.
template <int M>
__device__ __noinline__ int exact_step_cold(const float* __restrict__ T, const float4* __restrict__ Uj, uint64_t codes,
                                            uint64_t codes_hi, int j, int lane) {
  float4 a0 = __ldg(Uj + lane), a1 = __ldg(Uj + 32 + lane);
#pragma unroll 1
  for (int kk = 0; kk < M - 1; kk++) {                  // ascending k != j
    const int k = kk + (kk >= j);
    const uint32_t b = (uint32_t)((k < 8 ? codes >> (8 * k) : codes_hi >> (8 * (k - 8))) & 255u);
    const float4* row = reinterpret_cast<const float4*>(T + (((size_t)j * M + k) * kH + b) * kH);
    const float4 r0 = __ldg(row + lane), r1 = __ldg(row + 32 + lane);
    a0.x = __fadd_rn(a0.x, r0.x); a0.y = __fadd_rn(a0.y, r0.y);
    a0.z = __fadd_rn(a0.z, r0.z); a0.w = __fadd_rn(a0.w, r0.w);
    a1.x = __fadd_rn(a1.x, r1.x); a1.y = __fadd_rn(a1.y, r1.y);
    a1.z = __fadd_rn(a1.z, r1.z); a1.w = __fadd_rn(a1.w, r1.w);
  }
  float bv = a0.x;
  int bc = lane * 4;
  if (a0.y < bv) { bv = a0.y; bc = lane * 4 + 1; }
  if (a0.z < bv) { bv = a0.z; bc = lane * 4 + 2; }
  if (a0.w < bv) { bv = a0.w; bc = lane * 4 + 3; }
  if (a1.x < bv) { bv = a1.x; bc = 128 + lane * 4; }
  if (a1.y < bv) { bv = a1.y; bc = 128 + lane * 4 + 1; }
  if (a1.z < bv) { bv = a1.z; bc = 128 + lane * 4 + 2; }
  if (a1.w < bv) { bv = a1.w; bc = 128 + lane * 4 + 3; }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, bv, off);
    int oc = __shfl_xor_sync(0xffffffffu, bc, off);
    if (ov < bv || (ov == bv && oc < bc)) { bv = ov; bc = oc; }
  }
  return __shfl_sync(0xffffffffu, bc, 0);   // NaN sums compare false everywhere: lane 0's view (c = 0 first)
}

// PF: quantised pre-filter.  The step's 256 sums are first formed in INTEGERS, in units of scale_j, from the quantised
// tables (half the bytes of the fp32 rows): S(c) = rint(U_j[c]/scale_j) + sum_k q_jk[b_k][c], with
//   |scale_j*S(c) - exact(c)| <= scale_j * ((M-1)*0.51 + 0.75)      quantisation of the rows (K2q) and of the unary
//                              + 2^-20 * (umax + (M-1)*tmax_j)       fp32 roundings of the exact chain
// =: delta.  Every candidate that can be the exact first-minimum has S <= min(S) + 2*delta/scale_j.  If exactly ONE
// candidate is inside that window it IS the reference's argmin and the step is done; otherwise (near-ties, ~5 % of
// the steps, tools/q16_prefilter_probe.py) the exact fp32 chains of the window's candidates decide (or, beyond 4 / 8
// candidates, the exact fp32 rows).  Bit-identical by construction; the window W0_j + slack*inv_j is prepared per codebook by pf_consts_kernel.
#ifndef RYL_K3_BLOCKS
#define RYL_K3_BLOCKS 4
#endif
// resident blocks per SM the register allocation is aimed at (m > 8 with shared-memory unaries: 3, 70 KB each)
template <int M, bool PF>
constexpr int k3_blocks() { return (PF && M <= 8) ? RYL_K3_BLOCKS : (PF && RYL_K3_UQS16) ? 3 : 4; }

template <int M, bool PF, bool JSPEC = false>
__global__ void __launch_bounds__(256, (k3_blocks<M, PF>())) icm_warp_kernel(IcmParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int dpad = (p.d + 3) & ~3;                         // per-warp slot, 16-byte aligned (warp_cost loads float4)
  // UQS (m <= 8): the warp keeps its vector's unaries in shared memory, already in the pre-filter's integer units
  // (16 bit, relative to the codebook's smallest one): [M][32 lanes] x 16 B in front of the cost scratch
  constexpr bool UQS = PF && (M <= 8 || RYL_K3_UQS16);
  constexpr bool P14 = UQS && RYL_K3_QBITS == 14;      // 14-bit table fields (pf_rows14)
  constexpr int PCB = M <= 8 ? 64 : 128;               // bytes of window constants per warp
  unsigned char* smem_f = smem_raw + (UQS ? (size_t)nwarps * (M * 512 + PCB) : 0);
  uint4* uqw = reinterpret_cast<uint4*>(smem_raw + (size_t)warp * (UQS ? M * 512 : 0)) + lane;
  // + the warp's own copy of {1/scale_j, W0_j}, W0 = +inf for a codebook whose unaries saturate for this vector
  float2* pcw = reinterpret_cast<float2*>(smem_raw + (UQS ? (size_t)nwarps * M * 512 + (size_t)warp * PCB : 0));
  float* sq = reinterpret_cast<float*>(smem_f) + (size_t)warp * dpad;
  int* stats_s = reinterpret_cast<int*>(reinterpret_cast<float*>(smem_f) + (size_t)nwarps * dpad);
  for (int i = threadIdx.x; i < 2 * p.ilsiter; i += blockDim.x) stats_s[i] = 0;
  __syncthreads();

  unsigned long long nsteps = 0, nexact = 0;
  const uint32_t one = p.one;
  const uint64_t pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
  // Dynamic schedule: a vector's time depends on how many of its steps the memoisation skips, so a fixed stride
  // leaves the slowest warp ~7 % behind the mean (1M vectors over 4736 warps); every warp draws its next vector
  // from one global counter instead (one atomic per vector).  Results do not depend on the schedule.
  for (;;) {
    unsigned long long lnext = 0;
    if (lane == 0) lnext = atomicAdd(p.next, 1ull);
    const int64_t l = (int64_t)__shfl_sync(0xffffffffu, lnext, 0);
    if (l >= p.nc) break;
    const float* x = p.X + (size_t)l * p.d;
    Code cur = load_code<M>(p.B + (size_t)l * M);
    float curcost = RYL_K3_COSTCALL ? warp_cost_call<M>(x, p.C, cur.lo, cur.hi, p.d, sq, lane, __int_as_float(0x7f800000))
                                    : warp_cost<M>(x, p.C, cur, p.d, sq, lane, __int_as_float(0x7f800000));   // prevcost, src/LSQ.jl:201
    const float4* Ul = reinterpret_cast<const float4*>(p.U + (size_t)l * M * kH);
    uint32_t vsteps = 0, vexact = 0;                            // this vector's step counters (32 bit in the hot loop)
    float slack = 0.f;                                          // 2.002 * 2^-20 * umax (PF)
    if (PF) slack = __uint_as_float(__ldg(p.umax + l)) * (2.002f * 9.5367431640625e-07f);
    if constexpr (UQS) {
      // The pre-filter only ever needs rint(u(c) / scale_j), and only up to a constant per (vector, j): the warp reads
      // its vector's 8 KB of fp32 unaries ONCE (streaming), rounds them exactly as the step used to (one fma with
      // 1.5 * 2^23), subtracts the codebook's minimum and keeps the result as 16-bit fields in the row layout of the
      // quantised tables -- a step then adds the unary like one more row (LDS.128 instead of two LDG.128 from L2:
      // -22 % of the step's L2 bytes, no per-step conversion).  If a candidate lies more than 65535 units above the
      // minimum (8 * tmax_j with 14-bit tables) the fields cannot hold it: the warp's copy of the window constant
      // is set to +inf for that codebook, which sends its steps down the exact path.  Garbage for a j whose window
      // test fails anyway (umax / scale_j >= 2^21, NaN) is never used: that test is per step and unchanged.
      __syncwarp();
#pragma unroll 1
      for (int j = 0; j < M; j++) {
        const float2 pcj = __ldg(p.pfc + j);
        const float inv = pcj.x;
        const float4 u0 = __ldcs(Ul + j * 64 + lane), u1 = __ldcs(Ul + j * 64 + 32 + lane);
        const float uu[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
        int q[8];
#pragma unroll
        for (int i = 0; i < 8; i++) q[i] = (int)(__float_as_uint(fmaf(uu[i], inv, 12582912.0f)) - 0x4B400000u);
        const int lm = min(min(min(q[0], q[1]), min(q[2], q[3])), min(min(q[4], q[5]), min(q[6], q[7])));
        const int lx = max(max(max(q[0], q[1]), max(q[2], q[3])), max(max(q[4], q[5]), max(q[6], q[7])));
        const int base = __reduce_min_sync(0xffffffffu, lm);
        const bool sat = (uint32_t)__reduce_max_sync(0xffffffffu, lx) - (uint32_t)base > 65535u;   // unsigned: garbage q may wrap
        if (lane == 0) pcw[j] = make_float2(inv, sat ? __int_as_float(0x7f800000) : pcj.y);
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; i++)
          w[i] = min((uint32_t)q[i] - (uint32_t)base, 65535u) | (min((uint32_t)q[4 + i] - (uint32_t)base, 65535u) << 16);
        uqw[j * 32] = make_uint4(w[0], w[1], w[2], w[3]);
      }
      __syncwarp();
    }

    for (int it = 0; it < p.ilsiter; it++) {
      Code nb = cur;                                            // copyto!(B, oldB), src/LSQ.jl:207
      {                                                         // perturb_codes!, src/LSQ.jl:225 (draws precomputed)
        const int nblk = (p.npert + 3) >> 2;
        for (int jb = 0; jb < nblk; jb++) {
          const unsigned long long pk = __ldg(p.draws + ((size_t)l * p.ilsiter + it) * nblk + jb);
#pragma unroll
          for (int t = 0; t < 4; t++)
            if (jb * 4 + t < p.npert)
              code_set<M>(nb, (int)(pk >> (4 * t)) & 15, (uint32_t)(pk >> (16 + 8 * t)) & 255u);
        }
      }
      const unsigned long long ordp = __ldg(p.orders_packed + it);   // the whole visiting order in a register
      // Memoised conditioning: the step for codebook j is a pure function of (U_j, codes of the others).
      // If none of the other codes changed since j was last evaluated in THIS iteration, its argmin is the
      // code it already holds, so the step is skipped -- bit-identical to running it (all icmiter sweeps
      // "always run" in the reference, encode_icm.cpp / src/LSQ.jl:64-78, they just cannot change anything).
      uint32_t dirty = (1u << M) - 1u;
      for (int sweep = 0; sweep < p.icmiter && dirty; sweep++) {
        for (int s = 0; s < M; s++) {
          const int j = (M <= 8) ? (int)(((uint32_t)ordp >> (4 * s)) & 15u) : (int)((ordp >> (4 * s)) & 15ull);
          if (!((dirty >> j) & 1u)) continue;
          vsteps++;
          float4 a0, a1;
          if constexpr (!UQS) {
            if (M > 8) {
              a0 = ldg_f4_hint(Ul + j * 64 + lane, pol_keep);
              a1 = ldg_f4_hint(Ul + j * 64 + 32 + lane, pol_keep);
            } else {
              a0 = __ldg(Ul + j * 64 + lane);
              a1 = __ldg(Ul + j * 64 + 32 + lane);
            }
          }
          int bc = -1;
          if (PF) {
            const float2 pc = UQS ? pcw[j] : __ldg(p.pfc + j);    // {1/scale_j, W0_j}
            const float wf = fmaf(slack, pc.x, pc.y);             // window, in units of scale_j
            // umax/scale_j < 2^21: integer sums fit (else exact).  UQS: the rows are fetched before the test is known
            // (it only fails for degenerate inputs) so that the shared-memory load is off the critical path
            if (UQS || wf < pc.y + 4.0f) {
              // ---- quantised pass: lane owns c = 4*lane + w (lo halves) and 128 + 4*lane + w (hi halves) ---------
              const char* tqj = reinterpret_cast<const char*>(p.Tq) + (size_t)j * (M * kH * 512) + lane * 16;
              asm volatile("" : "+l"(tqj));   // keep the base in a register pair: row address = one IMAD.WIDE
              int S[8];
              if constexpr (P14) {                                  // 14-bit fields: rows summed as whole words
                const uint4 xu = uqw[j * 32];
                if constexpr (JSPEC && RYL_K3_HALF && M >= 5) {     // two copies: j in its group of four
                  if (j < 4) pf_rows14_half<M, true>(tqj, nb, xu, S, j);
                  else pf_rows14_half<M, false>(tqj, nb, xu, S, j);
                } else if constexpr (JSPEC) {                       // one copy of the row loop per j (jump table)
                  switch (j) {
                    case 0: pf_rows14<M, 0>(tqj, nb, xu, S, one); break;
                    case 1: pf_rows14<M, (M > 1 ? 1 : 0)>(tqj, nb, xu, S, one); break;
                    case 2: pf_rows14<M, (M > 2 ? 2 : 0)>(tqj, nb, xu, S, one); break;
                    case 3: pf_rows14<M, (M > 3 ? 3 : 0)>(tqj, nb, xu, S, one); break;
                    case 4: pf_rows14<M, (M > 4 ? 4 : 0)>(tqj, nb, xu, S, one); break;
                    case 5: pf_rows14<M, (M > 5 ? 5 : 0)>(tqj, nb, xu, S, one); break;
                    case 6: pf_rows14<M, (M > 6 ? 6 : 0)>(tqj, nb, xu, S, one); break;
                    default: pf_rows14<M, (M > 7 ? 7 : 0)>(tqj, nb, xu, S, one); break;
                  }
                } else {
                  // uniform row loop: all M rows, the diagonal one (k == j) is a row of zero words -- no dispatch on j,
                  // one copy of the code, at the price of one more 512 B row per step
                  constexpr int NG = (M + 3) / 4;                  // groups of <= 4 rows: no carry out of a 14-bit field sum
                  uint32_t G[NG][4];
#if RYL_K3_DIAG0ROW
                  // measured variant (off): the diagonal row is always row 0 of block [j][j] -- M hot rows that stay in L1
                  // instead of one of 256 M zero rows in L2; no gain at m = 12 / 16, 3 % slower at m <= 8
                  Code nz = nb;
                  if (M <= 8 || j < 8) nz.lo &= ~(0xFFull << (8 * (j & 7)));
                  else nz.hi &= ~(0xFFull << (8 * (j & 7)));
#else
                  const Code nz = nb;
#endif
#pragma unroll
                  for (int k = 0; k < M; k++) {
                    const uint32_t word = (uint32_t)((k < 8 ? nz.lo : nz.hi) >> (32 * ((k & 7) >> 2)));
                    const uint32_t code = __byte_perm(word, 0, 0x4440 | (k & 3));
                    const uint4 x = __ldg(reinterpret_cast<const uint4*>(tqj + (size_t)k * (kH * 512) + (size_t)code * 512));
                    if ((k & 3) == 0) { G[k >> 2][0] = x.x; G[k >> 2][1] = x.y; G[k >> 2][2] = x.z; G[k >> 2][3] = x.w; }
                    else { G[k >> 2][0] += x.x; G[k >> 2][1] += x.y; G[k >> 2][2] += x.z; G[k >> 2][3] += x.w; }
                  }
                  const uint32_t xw[4] = {xu.x, xu.y, xu.z, xu.w};
#pragma unroll
                  for (int w = 0; w < 4; w++) {
                    uint32_t h = xw[w] >> 16, t = xw[w];
#pragma unroll
                    for (int g = 0; g < NG; g++) { h += G[g][w] >> 16; t += G[g][w]; }
                    S[w] = (int)(t - (h << 16));
                    S[4 + w] = (int)h;
                  }
                }
              } else {
                uint32_t sl[4] = {0, 0, 0, 0}, sh[4] = {0, 0, 0, 0};   // sl = sum lo + (sum hi << 16) (mod 2^32)
                if constexpr (UQS) {                                  // the unary is row 0 of the sum
                  const uint4 xu = uqw[j * 32];
                  sl[0] = xu.x; sl[1] = xu.y; sl[2] = xu.z; sl[3] = xu.w;
                  sh[0] = xu.x >> 16; sh[1] = xu.y >> 16; sh[2] = xu.z >> 16; sh[3] = xu.w >> 16;
                }
                if constexpr (M <= 8 && JSPEC) {                     // one copy of the row loop per j (jump table)
                  switch (j) {
                    case 0: pf_rows<M, 0>(tqj, nb, sl, sh, one); break;
                    case 1: pf_rows<M, 1>(tqj, nb, sl, sh, one); break;
                    case 2: pf_rows<M, 2>(tqj, nb, sl, sh, one); break;
                    case 3: pf_rows<M, 3>(tqj, nb, sl, sh, one); break;
                    case 4: pf_rows<M, 4>(tqj, nb, sl, sh, one); break;
                    case 5: pf_rows<M, 5>(tqj, nb, sl, sh, one); break;
                    case 6: pf_rows<M, 6>(tqj, nb, sl, sh, one); break;
                    default: pf_rows<M, 7>(tqj, nb, sl, sh, one); break;
                  }
                } else {
#pragma unroll
                  for (int k = 0; k < M; k++) {                      // k is a literal: byte extract + immediate offsets
                    if (k != j) {
                      const uint32_t word = (uint32_t)((k < 8 ? nb.lo : nb.hi) >> (32 * ((k & 7) >> 2)));
                      const uint32_t code = __byte_perm(word, 0, 0x4440 | (k & 3));
                      const uint4 x = __ldg(reinterpret_cast<const uint4*>(tqj + (size_t)k * (kH * 512) + (size_t)code * 512));
                      sl[0] += x.x; sl[1] += x.y; sl[2] += x.z; sl[3] += x.w;
                      sh[0] += x.x >> 16; sh[1] += x.y >> 16; sh[2] += x.z >> 16; sh[3] += x.w >> 16;
                    }
                  }
                }
                // S(c) = rint(u(c)/scale_j) + sum_k (q_k(c) - 32768): the unary is rounded by the 1.5*2^23 trick inside
                // one fma, whose integer image carries the constant 0x4B400000
                if constexpr (UQS) {                                  // constants per step do not move the argmin
#pragma unroll
                  for (int w4 = 0; w4 < 4; w4++) {
                    S[w4] = (int)(sl[w4] - (sh[w4] << 16));
                    S[4 + w4] = (int)sh[w4];
                  }
                } else {
                  constexpr int K = -0x4B400000 - (M - 1) * 32768;
                  const float uu[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                  for (int w4 = 0; w4 < 4; w4++) {
                    S[w4] = (int)(sl[w4] - (sh[w4] << 16)) + (__float_as_int(fmaf(uu[w4], pc.x, 12582912.0f)) + K);
                    S[4 + w4] = (int)sh[w4] + (__float_as_int(fmaf(uu[4 + w4], pc.x, 12582912.0f)) + K);
                  }
                }
              }
              // candidate keys S * 8 + (slot in the lane): the warp minimum names the winner (lane, slot) directly
              int K[8];
#pragma unroll
              for (int i = 0; i < 8; i++) K[i] = S[i] * 8 + i;
              // smallest (lm) and second smallest (m2) key of the lane: a tournament that does not wait for the warp
              // reduction; the keys of a lane are distinct
              const int p0 = min(K[0], K[1]), P0 = max(K[0], K[1]), p1 = min(K[2], K[3]), P1 = max(K[2], K[3]);
              const int p2 = min(K[4], K[5]), P2 = max(K[4], K[5]), p3 = min(K[6], K[7]), P3 = max(K[6], K[7]);
              const int q0 = min(p0, p1), Q0 = min(min(max(p0, p1), P0), P1);
              const int q1 = min(p2, p3), Q1 = min(min(max(p2, p3), P2), P3);
              const int m2 = min(min(max(q0, q1), Q0), Q1);
              const int lm = min(q0, q1);
              const int key = __reduce_min_sync(0xffffffffu, (lm << 5) | lane);      // |S| < 2^22
              const int thr8 = ((key >> 8) + (int)wf + 1) * 8;      // K < thr8  <=>  S <= min S + window
              const int wl = key & 31;
              // exactly one candidate inside the window: the winner's lane holds no second one, the others none
              const bool unique = __all_sync(0xffffffffu, (lane == wl ? m2 : lm) >= thr8);
              const bool pf_ok = !UQS || wf < pc.y + 4.0f;
              if (unique && pf_ok) bc = (key & 128) | (wl << 2) | ((key >> 5) & 3);
#if RYL_K3_LITE
              if (!unique && pf_ok) {
                // Near-tie: every candidate that can be the exact first-minimum is inside the window.  With at most
                // four (m > 8: eight) of them only THEIR exact sums are formed, 32 / G at a time -- G = 8 (16) lanes per candidate fetch
                // its unary and its M-1 table entries (32 B sectors instead of the M-1 whole fp32 rows), the chain
                // ((u + r_1) + r_2) + ... is added in ascending k by shuffles inside the group, and the smallest
                // (value, c) wins.
                constexpr int G = M <= 8 ? 8 : 16;
                uint32_t mm = 0;
#pragma unroll
                for (int i = 0; i < 8; i++) mm |= (K[i] < thr8) ? (1u << i) : 0u;
                const int total = __reduce_add_sync(0xffffffffu, __popc(mm));
                if (total <= (M <= 8 ? 4 : 8)) {                  // m <= 8: one round of four
                  unsigned long long cs = 0;                      // the candidates, 8 bits each
                  for (int r = 0; r < total; r++) {
                    const int L = __ffs(__ballot_sync(0xffffffffu, mm != 0)) - 1;
                    const int i = __shfl_sync(0xffffffffu, __ffs(mm) - 1, L);
                    if (lane == L) mm &= mm - 1;
                    cs |= (unsigned long long)(((i & 4) << 5) | (L << 2) | (i & 3)) << (8 * r);
                  }
                  float best_v = __int_as_float(0x7f800000);
                  int best_c = 512;
                  for (int base = 0; base < (M <= 8 ? 1 : total); base += 32 / G) {   // 32 / G candidates per round
                    const int r = base + lane / G, t = lane % G;
                    const int c = (int)(cs >> (8 * (r & 7))) & 255;
                    float v = 0.f;
                    if (r < total) {
                      if (t == G - 1) v = __ldg(p.U + ((size_t)l * M + j) * kH + c);
                      else if (t < M - 1) {
                        const int k = t + (t >= j);
                        v = __ldg(p.T + (((size_t)j * M + k) * kH + code_get<M>(nb, k)) * kH + c);
                      }
                    }
                    float acc = __shfl_sync(0xffffffffu, v, G - 1, G);
#pragma unroll
                    for (int tt = 0; tt < M - 1; tt++) acc = __fadd_rn(acc, __shfl_sync(0xffffffffu, v, tt, G));
                    float bv = r < total ? acc : __int_as_float(0x7f800000);
                    int bcc = r < total ? c : 256 + r;
#pragma unroll
                    for (int off = G; off <= 16; off <<= 1) {
                      const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
                      const int oc = __shfl_xor_sync(0xffffffffu, bcc, off);
                      if (ov < bv || (ov == bv && oc < bcc)) { bv = ov; bcc = oc; }
                    }
                    bv = __shfl_sync(0xffffffffu, bv, 0);
                    bcc = __shfl_sync(0xffffffffu, bcc, 0);
                    if (bv < best_v || (bv == best_v && bcc < best_c)) { best_v = bv; best_c = bcc; }
                  }
                  bc = best_c;
                }
              }
#endif
            }
          }
          if (bc < 0) {
            if (PF) vexact++;
            if constexpr (UQS && RYL_K3_COLD) {
              bc = exact_step_cold<M>(p.T, Ul + j * 64, nb.lo, nb.hi, j, lane);
            } else {
            if constexpr (UQS) {
              a0 = __ldg(Ul + j * 64 + lane);
              a1 = __ldg(Ul + j * 64 + 32 + lane);
            }
#pragma unroll
            for (int kk = 0; kk < M - 1; kk++) {                  // ascending k != j, encode_icm.cpp:28-45
              const int k = kk + (kk >= j);
              const float4* row =
                  reinterpret_cast<const float4*>(p.T + (((size_t)j * M + k) * kH + code_get<M>(nb, k)) * kH);
              float4 r0, r1;
              if (PF && M > 8) {
                r0 = ldg_f4_hint(row + lane, pol_stream);
                r1 = ldg_f4_hint(row + 32 + lane, pol_stream);
              } else {
                r0 = __ldg(row + lane);
                r1 = __ldg(row + 32 + lane);
              }
              a0.x = __fadd_rn(a0.x, r0.x); a0.y = __fadd_rn(a0.y, r0.y);
              a0.z = __fadd_rn(a0.z, r0.z); a0.w = __fadd_rn(a0.w, r0.w);
              a1.x = __fadd_rn(a1.x, r1.x); a1.y = __fadd_rn(a1.y, r1.y);
              a1.z = __fadd_rn(a1.z, r1.z); a1.w = __fadd_rn(a1.w, r1.w);
            }
            // first-minimum argmin (encode_icm.cpp:47-58): ascending c inside the lane, then a
            // (value, index)-lexicographic butterfly across lanes
            float bv = a0.x;
            bc = lane * 4;
            if (a0.y < bv) { bv = a0.y; bc = lane * 4 + 1; }
            if (a0.z < bv) { bv = a0.z; bc = lane * 4 + 2; }
            if (a0.w < bv) { bv = a0.w; bc = lane * 4 + 3; }
            if (a1.x < bv) { bv = a1.x; bc = 128 + lane * 4; }
            if (a1.y < bv) { bv = a1.y; bc = 128 + lane * 4 + 1; }
            if (a1.z < bv) { bv = a1.z; bc = 128 + lane * 4 + 2; }
            if (a1.w < bv) { bv = a1.w; bc = 128 + lane * 4 + 3; }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
              float ov = __shfl_xor_sync(0xffffffffu, bv, off);
              int oc = __shfl_xor_sync(0xffffffffu, bc, off);
              if (ov < bv || (ov == bv && oc < bc)) { bv = ov; bc = oc; }
            }
            bc = __shfl_sync(0xffffffffu, bc, 0);   // NaN sums compare false everywhere: take lane 0's view (c = 0
                                                    // first, like the sequential scan of encode_icm.cpp:47-58)
            }
          }
          dirty &= ~(1u << j);
          if ((uint32_t)bc != code_get<M>(nb, j)) {
            code_set<M>(nb, j, (uint32_t)bc);
            dirty |= ((1u << M) - 1u) & ~(1u << j);             // everyone conditioned on j must be redone
          }
        }
      }
      // newcost, src/LSQ.jl:237.  When ICM led back to the codes the iteration started from (the perturbation was
      // undone -- the common case late in the search) the cost is the same arithmetic on the same inputs: reuse it.
      const bool same = nb.lo == cur.lo && nb.hi == cur.hi;
      // m <= 7: squares + tree sum inline (the common, rejected case ends there), the reference-order chain as a call;
      // m = 8, whose hot loop sits at the instruction cache's capacity, calls the whole evaluation (measured: 176 vs 181 ms
      // at m = 8, 142 vs 140 ms at m = 7)
      float newcost = curcost;
      if (!same) {
        if constexpr (M <= 7 && RYL_K3_REJECT) {
          newcost = warp_cost<M>(x, p.C, nb, p.d, sq, lane, curcost, true);
          if (!(newcost > curcost * (1.0f + 1.3e-7f * (float)(p.d + 32)))) newcost = cost_seq_call(sq, p.d);
        } else {
          newcost = RYL_K3_COSTCALL ? warp_cost_call<M>(x, p.C, nb.lo, nb.hi, p.d, sq, lane, RYL_K3_REJECT ? curcost : __int_as_float(0x7f800000))
                                    : warp_cost<M>(x, p.C, nb, p.d, sq, lane, RYL_K3_REJECT ? curcost : __int_as_float(0x7f800000));
        }
      }
      if (lane == 0) {
        if (newcost == curcost) atomicAdd(&stats_s[2 * it], 1);
        if (newcost < curcost) atomicAdd(&stats_s[2 * it + 1], 1);
      }
      if (newcost < curcost) {                                   // strict <, src/LSQ.jl:242-247
        cur = nb;
        curcost = newcost;
      }
      for (int sidx = 0; sidx < p.n_snap; sidx++)                // ilsiters snapshots, src/LSQ_GPU.jl:193-204
        if (__ldg(p.snap_iters + sidx) == it + 1 && lane < M)
          p.B_snap[((size_t)sidx * p.n_total + l) * M + lane] = (uint8_t)cur.get(lane);
    }
    if (lane < M) p.B[(size_t)l * M + lane] = (uint8_t)cur.get(lane);
    if (p.cost && lane == 0) p.cost[l] = curcost;
    nsteps += vsteps;
    nexact += vexact;
  }
  if (lane == 0 && nsteps) atomicAdd(p.steps, nsteps);
  if (lane == 0 && nexact) atomicAdd(p.steps + 1, nexact);
  __syncthreads();
  if (p.stats)
    for (int i = threadIdx.x; i < 2 * p.ilsiter; i += blockDim.x)
      if (stats_s[i]) atomicAdd(p.stats + i, stats_s[i]);
}

template <int M>
__global__ void __launch_bounds__(256) veccost_kernel(const float* __restrict__ X, const uint8_t* __restrict__ B,
                                                      const float* __restrict__ C, int64_t n, int d,
                                                      float* __restrict__ cost) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  float* sq = reinterpret_cast<float*>(smem_raw) + (size_t)warp * ((d + 3) & ~3);
  for (int64_t l = (int64_t)blockIdx.x * nwarps + warp; l < n; l += (int64_t)gridDim.x * nwarps) {
    Code c = load_code<M>(B + (size_t)l * M);
    float v = warp_cost<M>(X + (size_t)l * d, C, c, d, sq, lane, __int_as_float(0x7f800000));
    if (lane == 0) cost[l] = v;
  }
}


// "next" row 2: quantize_norms (src/utils.jl:29-59).  One warp per vector:
// CB = sum_k C_k[:, b_k] from +0 in codebook order (reconstruct, src/qerrors.jl:6-33), norm = sequential sum
// of CB[t]^2 (unfused), code = first minimum of (norm - cbnorms[c])^2 over the 256 norm centroids.
template <int M>
__global__ void __launch_bounds__(256) norms_kernel(const uint8_t* __restrict__ B, const float* __restrict__ C,
                                                    const float* __restrict__ cbnorms, int64_t n, int d,
                                                    uint8_t* __restrict__ codes, float* __restrict__ norms) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  float* sq = reinterpret_cast<float*>(smem_raw) + (size_t)warp * ((d + 3) & ~3);
  for (int64_t l = (int64_t)blockIdx.x * nwarps + warp; l < n; l += (int64_t)gridDim.x * nwarps) {
    Code code = load_code<M>(B + (size_t)l * M);
    for (int t = lane; t < d; t += 32) {
      float cb = 0.f;
#pragma unroll
      for (int k = 0; k < M; k++) cb = __fadd_rn(cb, __ldg(C + ((size_t)k * kH + code.get(k)) * d + t));
      sq[t] = __fmul_rn(cb, cb);
    }
    __syncwarp();
    float nrm = 0.f;
    for (int t = 0; t < d; t++) nrm = __fadd_rn(nrm, sq[t]);
    __syncwarp();
    if (norms && lane == 0) norms[l] = nrm;
    if (cbnorms && codes) {
      float bv = 0.f;
      int bc = 0;
#pragma unroll
      for (int r = 0; r < 8; r++) {
        const int c = lane * 8 + r;
        const float df = __fsub_rn(nrm, __ldg(cbnorms + c));
        const float v = __fmul_rn(df, df);
        if (r == 0 || v < bv) { bv = v; bc = c; }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, bv, off);
        int oc = __shfl_xor_sync(0xffffffffu, bc, off);
        if (ov < bv || (ov == bv && oc < bc)) { bv = ov; bc = oc; }
      }
      if (lane == 0) codes[l] = (uint8_t)bc;
    }
  }
}

// "next" row 3: ChainQ Viterbi encode (quantize_chainq, src/ChainQ.jl:36-128; C++ twin
// deps/src/encode_icm.cpp:63-152).  Block = 256 threads (one per destination state j) x VB vectors.
// Forward pass i = 0..m-2: cost[k] = U_i[k] + bb_i[k -> j], first-minimum over ascending k (strict <);
// mincost[j] is added to U_{i+1}[j] (:97-100,123-125); final first-minimum over U_{m-1}; back-trace.
// TT[i] is the transposed chain table, TT[i][k*256 + j] = 2<C_i[:,k], C_{i+1}[:,j]>, so the block's reads are
// coalesced over j and every table element is read once per VB vectors.
template <int VB>
__global__ void __launch_bounds__(256) viterbi_kernel(const float* __restrict__ U, const float* __restrict__ TT,
                                                      int64_t tt_stride, int64_t n, int m,
                                                      uint8_t* __restrict__ B) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* ucur = reinterpret_cast<float*>(smem_raw);                      // [VB][256]
  uint8_t* minidx = reinterpret_cast<uint8_t*>(ucur + VB * kH);          // [VB][m-1][256]
  __shared__ float red_v[8][VB];
  __shared__ int red_i[8][VB];
  const int j = threadIdx.x, lane = j & 31, warp = j >> 5;
  for (int64_t v0 = (int64_t)blockIdx.x * VB; v0 < n; v0 += (int64_t)gridDim.x * VB) {
    float carry[VB];                                                     // mincost[j] of the previous stage
#pragma unroll
    for (int v = 0; v < VB; v++) carry[v] = 0.f;
    for (int i = 0; i < m; i++) {
      // U_i[j] (+ mincost[j] for i > 0: "add the precomputed costs")
      float ui[VB];
#pragma unroll
      for (int v = 0; v < VB; v++) {
        const int64_t l = min(v0 + v, n - 1);
        const float u = U[((size_t)l * m + i) * kH + j];
        ui[v] = i > 0 ? __fadd_rn(u, carry[v]) : u;
      }
      if (i == m - 1) {
        // final first-minimum over j, then the backward trace (one thread per vector)
#pragma unroll
        for (int v = 0; v < VB; v++) {
          float bv = ui[v];
          int bj = j;
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, bv, off);
            int oj = __shfl_xor_sync(0xffffffffu, bj, off);
            if (ov < bv || (ov == bv && oj < bj)) { bv = ov; bj = oj; }
          }
          if (lane == 0) { red_v[warp][v] = bv; red_i[warp][v] = bj; }
        }
        __syncthreads();
        if (j < VB && v0 + j < n) {
          float bv = red_v[0][j];
          int bj = red_i[0][j];
          for (int w = 1; w < 8; w++)
            if (red_v[w][j] < bv || (red_v[w][j] == bv && red_i[w][j] < bj)) { bv = red_v[w][j]; bj = red_i[w][j]; }
          uint8_t* out = B + (size_t)(v0 + j) * m;
          int state = bj;
          out[m - 1] = (uint8_t)state;
          for (int ii = m - 2; ii >= 0; ii--) {
            state = minidx[((size_t)j * (m - 1) + ii) * kH + state];
            out[ii] = (uint8_t)state;
          }
        }
        __syncthreads();
        break;
      }
#pragma unroll
      for (int v = 0; v < VB; v++) ucur[v * kH + j] = ui[v];
      __syncthreads();
      const float* tt = TT + (size_t)i * tt_stride;
      float best[VB];
      int bi[VB];
#pragma unroll 4
      for (int k = 0; k < kH; k++) {
        const float t = __ldg(tt + (size_t)k * kH + j);
#pragma unroll
        for (int v = 0; v < VB; v++) {
          const float c = __fadd_rn(ucur[v * kH + k], t);
          if (k == 0 || c < best[v]) { best[v] = c; bi[v] = k; }
        }
      }
#pragma unroll
      for (int v = 0; v < VB; v++) {
        carry[v] = best[v];
        minidx[((size_t)v * (m - 1) + i) * kH + j] = (uint8_t)bi[v];
      }
      __syncthreads();
    }
  }
}

// bb[i][j*256 + k] (the reference's binaries[i], column-major h-by-h) -> TT[i][k*256 + j]
__global__ void transpose_tables_kernel(const float* __restrict__ in, float* __restrict__ out, int count) {
  __shared__ float tile[32][33];
  const float* src = in + (size_t)blockIdx.z * kH * kH;
  float* dst = out + (size_t)blockIdx.z * kH * kH;
  const int x = blockIdx.x * 32 + threadIdx.x, y0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) tile[r][threadIdx.x] = src[(size_t)(y0 + r) * kH + x];
  __syncthreads();
  const int xo = blockIdx.y * 32 + threadIdx.x, yo0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += 8) dst[(size_t)(yo0 + r) * kH + xo] = tile[threadIdx.x][r];
}

// deterministic two-stage sum in double (qerror = mean(veccost), src/qerrors.jl:69-74)
__global__ void __launch_bounds__(256) sum_kernel(const float* __restrict__ v, int64_t n, double* __restrict__ partial) {
  __shared__ double sh[256];
  const int64_t per = (n + gridDim.x - 1) / gridDim.x;
  const int64_t b = (int64_t)blockIdx.x * per, e = min(n, b + per);
  double s = 0;
  for (int64_t i = b + threadIdx.x; i < e; i += 256) s += (double)v[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

// compat: one conditioning step with the reference's own argument layout (deps/src/encode_icm.cpp:3-61)
__global__ void __launch_bounds__(256) condition_kernel(uint8_t* __restrict__ B, float* __restrict__ ub,
                                                        const float* __restrict__ binaries,
                                                        const float* __restrict__ binaries_t,
                                                        const int* __restrict__ pair2idx,
                                                        const int* __restrict__ to_condition, int j, int n, int m) {
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t l = w; l < n; l += nw) {
    float4* u4 = reinterpret_cast<float4*>(ub + (size_t)l * kH);
    float4 a0 = u4[lane], a1 = u4[32 + lane];
    for (int kidx = 0; kidx < m - 1; kidx++) {
      const int k = to_condition[kidx];
      const int bidx = pair2idx[j * m + k];
      const float* bb = (j < k ? binaries : binaries_t) + (size_t)kH * kH * bidx;
      const float4* row = reinterpret_cast<const float4*>(bb + (size_t)B[(size_t)l * m + k] * kH);
      float4 r0 = __ldg(row + lane), r1 = __ldg(row + 32 + lane);
      a0.x = __fadd_rn(a0.x, r0.x); a0.y = __fadd_rn(a0.y, r0.y);
      a0.z = __fadd_rn(a0.z, r0.z); a0.w = __fadd_rn(a0.w, r0.w);
      a1.x = __fadd_rn(a1.x, r1.x); a1.y = __fadd_rn(a1.y, r1.y);
      a1.z = __fadd_rn(a1.z, r1.z); a1.w = __fadd_rn(a1.w, r1.w);
    }
    u4[lane] = a0;
    u4[32 + lane] = a1;
    float bv = a0.x;
    int bc = lane * 4;
    if (a0.y < bv) { bv = a0.y; bc = lane * 4 + 1; }
    if (a0.z < bv) { bv = a0.z; bc = lane * 4 + 2; }
    if (a0.w < bv) { bv = a0.w; bc = lane * 4 + 3; }
    if (a1.x < bv) { bv = a1.x; bc = 128 + lane * 4; }
    if (a1.y < bv) { bv = a1.y; bc = 128 + lane * 4 + 1; }
    if (a1.z < bv) { bv = a1.z; bc = 128 + lane * 4 + 2; }
    if (a1.w < bv) { bv = a1.w; bc = 128 + lane * 4 + 3; }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, bv, off);
      int oc = __shfl_xor_sync(0xffffffffu, bc, off);
      if (ov < bv || (ov == bv && oc < bc)) { bv = ov; bc = oc; }
    }
    __syncwarp();
    if (lane == 0) B[(size_t)l * m + j] = (uint8_t)bc;
    __syncwarp();
  }
}

// ---- any h <= 256: iterated_conditional_modes! (src/LSQ.jl:83-149), the pure-Julia path the reference takes with
// cpp=false (experiment_lsq, src/LSQ.jl:431).  Same arithmetic contract as the 256-entry kernels (sequential fmaf
// dots, ascending-k fp32 adds, first-minimum argmin), plain exact rows -- no quantised pre-filter, no tiling tricks:
// codebooks with h != 256 are a compatibility path, not the benchmarked one.
__global__ void __launch_bounds__(256) unary_generic_kernel(const float* __restrict__ C, const float* __restrict__ X,
                                                            const float* __restrict__ nrm, float* __restrict__ U,
                                                            int64_t n, int d, int mh) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* xs = reinterpret_cast<float*>(smem_raw);
  for (int64_t l = blockIdx.x; l < n; l += gridDim.x) {
    __syncthreads();
    for (int t = threadIdx.x; t < d; t += blockDim.x) xs[t] = X[(size_t)l * d + t];
    __syncthreads();
    for (int e = threadIdx.x; e < mh; e += blockDim.x) {
      const float* c = C + (size_t)e * d;
      float s = 0.f;
      for (int t = 0; t < d; t++) s = fmaf(c[t], xs[t], s);
      U[(size_t)l * mh + e] = fmaf(-2.0f, s, nrm[e]);
    }
  }
}

// T[((j*m + k)*h + b)*h + c] = 2 <C_j[:,c], C_k[:,b]>
__global__ void __launch_bounds__(256) tables_generic_kernel(const float* __restrict__ C, float* __restrict__ T, int d,
                                                             int m, int h) {
  const int64_t total = (int64_t)m * m * h * h;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % h), b = (int)((i / h) % h);
    const int k = (int)((i / ((int64_t)h * h)) % m), j = (int)(i / ((int64_t)h * h * m));
    if (j == k) continue;
    const float* cj = C + ((size_t)j * h + c) * d;
    const float* ck = C + ((size_t)k * h + b) * d;
    float s = 0.f;
    for (int t = 0; t < d; t++) s = fmaf(cj[t], ck[t], s);
    T[i] = 2.0f * s;
  }
}

__device__ __forceinline__ float warp_cost_generic(const float* __restrict__ x, const float* __restrict__ C,
                                                   const Code& code, int d, int m, int h, float* sq, int lane) {
  for (int t = lane; t < d; t += 32) {
    float cb = 0.f;
    for (int k = 0; k < m; k++) cb = __fadd_rn(cb, __ldg(C + ((size_t)k * h + code.get(k)) * d + t));
    const float df = __fsub_rn(cb, __ldg(x + t));
    sq[t] = __fmul_rn(df, df);
  }
  __syncwarp();
  float acc = 0.f;
  for (int t = 0; t < d; t++) acc = __fadd_rn(acc, sq[t]);
  __syncwarp();
  return acc;
}

__global__ void __launch_bounds__(256) veccost_generic_kernel(const float* __restrict__ X, const uint8_t* __restrict__ B,
                                                              const float* __restrict__ C, int64_t n, int d, int m,
                                                              int h, float* __restrict__ cost) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  float* sq = reinterpret_cast<float*>(smem_raw) + (size_t)warp * ((d + 3) & ~3);
  for (int64_t l = (int64_t)blockIdx.x * nwarps + warp; l < n; l += (int64_t)gridDim.x * nwarps) {
    Code c{0, 0};
    for (int k = 0; k < m; k++) c.set(k, B[(size_t)l * m + k]);
    const float v = warp_cost_generic(X + (size_t)l * d, C, c, d, m, h, sq, lane);
    if (lane == 0) cost[l] = v;
  }
}

struct IcmGenericParams {
  IcmParams p;      // U [nc][m][h], T [m][m][h][h]; Tq / pfc / umax unused
  int m, h;
};

__global__ void __launch_bounds__(256) icm_generic_kernel(IcmGenericParams gp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const IcmParams& p = gp.p;
  const int M = gp.m, h = gp.h;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int dpad = (p.d + 3) & ~3;
  float* sq = reinterpret_cast<float*>(smem_raw) + (size_t)warp * dpad;
  int* stats_s = reinterpret_cast<int*>(reinterpret_cast<float*>(smem_raw) + (size_t)nwarps * dpad);
  for (int i = threadIdx.x; i < 2 * p.ilsiter; i += blockDim.x) stats_s[i] = 0;
  __syncthreads();
  const float inf = __int_as_float(0x7f800000);
  unsigned long long nsteps = 0;
  for (int64_t l = (int64_t)blockIdx.x * nwarps + warp; l < p.nc; l += (int64_t)gridDim.x * nwarps) {
    const float* x = p.X + (size_t)l * p.d;
    Code cur{0, 0};
    for (int k = 0; k < M; k++) cur.set(k, p.B[(size_t)l * M + k]);
    float curcost = warp_cost_generic(x, p.C, cur, p.d, M, h, sq, lane);
    const float* Ul = p.U + (size_t)l * M * h;
    for (int it = 0; it < p.ilsiter; it++) {
      Code nb = cur;
      for (int jb = 0; jb * 4 < p.npert; jb++) {                          // perturb_codes!, src/LSQ.jl:5-39
        const uint64_t g = (uint64_t)(p.g0 + l);
        uint32_t pw[4] = {(uint32_t)g, (uint32_t)(g >> 32), (uint32_t)it, (uint32_t)jb};
        uint32_t vw[4] = {(uint32_t)g, (uint32_t)(g >> 32), (uint32_t)it, 0x80000000u | (uint32_t)jb};
        philox4x32_10(pw, (uint32_t)p.seed, (uint32_t)(p.seed >> 32));
        philox4x32_10(vw, (uint32_t)p.seed, (uint32_t)(p.seed >> 32));
#pragma unroll
        for (int t = 0; t < 4; t++)
          if (jb * 4 + t < p.npert) nb.set((int)mulhi32(pw[t], (uint32_t)M), mulhi32(vw[t], (uint32_t)h));
      }
      const int* order = p.orders + it * M;
      uint32_t dirty = (1u << M) - 1u;
      for (int sweep = 0; sweep < p.icmiter && dirty; sweep++) {
        for (int s = 0; s < M; s++) {
          const int j = __ldg(order + s);
          if (!((dirty >> j) & 1u)) continue;
          nsteps++;
          float a[8];                                                     // lane owns c = lane + 32*i
#pragma unroll
          for (int i = 0; i < 8; i++) a[i] = (lane + 32 * i < h) ? __ldg(Ul + (size_t)j * h + lane + 32 * i) : inf;
          for (int k = 0; k < M; k++) {                                   // ascending k != j, src/LSQ.jl:108-125
            if (k == j) continue;
            const float* row = p.T + (((size_t)j * M + k) * h + nb.get(k)) * h;
#pragma unroll
            for (int i = 0; i < 8; i++)
              if (lane + 32 * i < h) a[i] = __fadd_rn(a[i], __ldg(row + lane + 32 * i));
          }
          float bv = a[0];                                                // first minimum, src/LSQ.jl:128-142
          int bc = lane;
#pragma unroll
          for (int i = 1; i < 8; i++)
            if (a[i] < bv) { bv = a[i]; bc = lane + 32 * i; }
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
            const int oc = __shfl_xor_sync(0xffffffffu, bc, off);
            if (ov < bv || (ov == bv && oc < bc)) { bv = ov; bc = oc; }
          }
          bc = __shfl_sync(0xffffffffu, bc, 0);
          if (bc >= h) bc = 0;                                            // all sums NaN: the reference keeps index 1
          dirty &= ~(1u << j);
          if ((uint32_t)bc != nb.get(j)) {
            nb.set(j, (uint32_t)bc);
            dirty |= ((1u << M) - 1u) & ~(1u << j);
          }
        }
      }
      const float newcost = warp_cost_generic(x, p.C, nb, p.d, M, h, sq, lane);
      if (lane == 0) {
        if (newcost == curcost) atomicAdd(&stats_s[2 * it], 1);
        if (newcost < curcost) atomicAdd(&stats_s[2 * it + 1], 1);
      }
      if (newcost < curcost) {
        cur = nb;
        curcost = newcost;
      }
      for (int sidx = 0; sidx < p.n_snap; sidx++)
        if (__ldg(p.snap_iters + sidx) == it + 1 && lane < M)
          p.B_snap[((size_t)sidx * p.n_total + l) * M + lane] = (uint8_t)cur.get(lane);
    }
    if (lane < M) p.B[(size_t)l * M + lane] = (uint8_t)cur.get(lane);
    if (p.cost && lane == 0) p.cost[l] = curcost;
  }
  if (lane == 0 && nsteps) {
    atomicAdd(p.steps, nsteps);
    atomicAdd(p.steps + 1, nsteps);
  }
  __syncthreads();
  if (p.stats)
    for (int i = threadIdx.x; i < 2 * p.ilsiter; i += blockDim.x)
      if (stats_s[i]) atomicAdd(p.stats + i, stats_s[i]);
}

}  // namespace ryl

// ======================================================================================================
// host side
// ======================================================================================================
using namespace ryl;

static thread_local uint64_t g_icm_steps_done = 0, g_icm_steps_total = 0, g_icm_steps_exact = 0;
static thread_local float g_icm_ms[4] = {0, 0, 0, 0};   // setup (tables), unaries, ICM kernel, whole call -- CUDA events

extern "C" int rayuela_encode_icm_timings(float* ms4) {
  if (ms4) memcpy(ms4, g_icm_ms, sizeof g_icm_ms);
  return RAYUELA_OK;
}

extern "C" int rayuela_encode_icm_exact_steps(uint64_t* exact) {
  if (exact) *exact = g_icm_steps_exact;
  return RAYUELA_OK;
}

extern "C" int rayuela_encode_icm_steps(uint64_t* executed, uint64_t* total) {
  if (executed) *executed = g_icm_steps_done;
  if (total) *total = g_icm_steps_total;
  return RAYUELA_OK;
}

static bool env_off(const char* name) {
  const char* e = getenv(name);
  return e && *e && atoi(e) == 0;
}

// K1 dispatch: the FFMA2 kernel for d % 16 == 0 (RAYUELA_B200_K1_V2=0: the round-1 kernel), else the generic tiles
static int launch_unary(const float* C, const float* X, const float* nrm, float* U, int64_t n, int d, int mh,
                        unsigned int* umax, cudaStream_t s) {
  dim3 ug(mh / 128, (unsigned)((n + 127) / 128));
  if (d % 16 == 0 && !env_off("RAYUELA_B200_K1_V2"))
    RYL_LAUNCH(unary_kernel2, ug, 256, 0, s, C, X, nrm, U, n, d, mh, umax);
  else if (d % 4 == 0)
    RYL_LAUNCH(unary_kernel<true>, ug, 256, 0, s, C, X, nrm, U, n, d, mh, umax);
  else
    RYL_LAUNCH(unary_kernel<false>, ug, 256, 0, s, C, X, nrm, U, n, d, mh, umax);
  return RAYUELA_OK;
}

// tuning knobs: RAYUELA_B200_ICM_PF=0 disables the quantised pre-filter (every step reads the exact fp32 rows),
// RAYUELA_B200_ICM_JSPEC=0 the per-j specialised row loop (m <= 8: -3 % at m = 8, -7 % at m = 7)
template <int M, bool PF, bool JSPEC>
static int launch_icm_v(const IcmParams& p, size_t smem, cudaStream_t s) {
  const int warps = 8;
  RYL_CUDA(cudaFuncSetAttribute(icm_warp_kernel<M, PF, JSPEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t need = (p.nc + warps - 1) / warps;
  int per_sm = k3_blocks<M, PF>();                                          // blocks of 8 warps per SM
  if (const char* e = getenv("RAYUELA_B200_ICM_BLOCKS_PER_SM")) per_sm = std::max(1, std::min(per_sm, atoi(e)));   // knob
  const int grid = (int)std::min<int64_t>(need, (int64_t)sm_count() * per_sm);
  RYL_LAUNCH((icm_warp_kernel<M, PF, JSPEC>), grid, warps * 32, smem, s, p);
  return RAYUELA_OK;
}

template <int M>
static int launch_icm(const IcmParams& p, cudaStream_t s) {
  const int warps = 8;
  size_t smem = (size_t)warps * ((p.d + 3) & ~3) * sizeof(float) + (size_t)2 * p.ilsiter * sizeof(int);
  RYL_ARG(smem <= 160 * 1024, "encode_icm: d * 8 warps (+ ilsiter) exceeds shared memory");
  if (!p.Tq) return launch_icm_v<M, false, false>(p, smem, s);
  if (M <= 8 || RYL_K3_UQS16) smem += (size_t)warps * (M * 512 + (M <= 8 ? 64 : 128));   // quantised unaries + window constants (UQS)
  if constexpr (M <= 8) {
    if (!env_off("RAYUELA_B200_ICM_JSPEC")) return launch_icm_v<M, true, true>(p, smem, s);
  }
  return launch_icm_v<M, true, false>(p, smem, s);
}

template <int M>
static int launch_veccost(const float* X, const uint8_t* B, const float* C, int64_t n, int d, float* cost,
                          cudaStream_t s) {
  const int warps = 8;
  size_t smem = (size_t)warps * ((d + 3) & ~3) * sizeof(float);
  RYL_ARG(smem <= 200 * 1024, "veccost: d too large for shared memory");
  RYL_CUDA(cudaFuncSetAttribute(veccost_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = (int)std::min<int64_t>((n + warps - 1) / warps, (int64_t)sm_count() * 8);
  RYL_LAUNCH(veccost_kernel<M>, grid, warps * 32, smem, s, X, B, C, n, d, cost);
  return RAYUELA_OK;
}

#define RYL_M_SWITCH(m, CALL)                                                                     \
  switch (m) {                                                                                    \
    case 1: CALL(1); break;  case 2: CALL(2); break;  case 3: CALL(3); break;  case 4: CALL(4); break;    \
    case 5: CALL(5); break;  case 6: CALL(6); break;  case 7: CALL(7); break;  case 8: CALL(8); break;    \
    case 9: CALL(9); break;  case 10: CALL(10); break; case 11: CALL(11); break; case 12: CALL(12); break; \
    case 13: CALL(13); break; case 14: CALL(14); break; case 15: CALL(15); break; case 16: CALL(16); break; \
    default: return fail(RAYUELA_ERR_ARG, "m must be in 1..16");                                  \
  }

static int device_veccost(const float* X, const uint8_t* B, const float* C, int64_t n, int d, int m, float* cost,
                          cudaStream_t s) {
  int rc = RAYUELA_OK;
#define CALL(M) rc = launch_veccost<M>(X, B, C, n, d, cost, s)
  RYL_M_SWITCH(m, CALL)
#undef CALL
  return rc;
}

// mean of a device float array, deterministic; synchronises the stream
static int device_mean(const float* v, int64_t n, double* out, cudaStream_t s) {
  const int blocks = 256;
  DevBuf part;
  RYL_TRY(part.alloc(blocks * sizeof(double), s));
  RYL_LAUNCH(sum_kernel, blocks, 256, 0, s, v, n, part.as<double>());
  double h[blocks];
  RYL_CUDA(cudaMemcpyAsync(h, part.p, sizeof h, cudaMemcpyDeviceToHost, s));
  RYL_CUDA(cudaStreamSynchronize(s));
  double t = 0;
  for (int i = 0; i < blocks; i++) t += h[i];
  *out = t / (double)n;
  return RAYUELA_OK;
}

static size_t unary_budget_bytes() {
  const char* e = getenv("RAYUELA_B200_UNARY_BYTES");
  if (e && *e) return (size_t)strtoull(e, nullptr, 10);
  return (size_t)16 << 30;
}

// sum of a device float array in double, deterministic (256 fixed partial ranges); synchronises the stream
static int device_sum(const float* v, int64_t n, double* out, cudaStream_t s) {
  double mean = 0;
  RYL_TRY(device_mean(v, n, &mean, s));
  *out = mean * (double)n;
  return RAYUELA_OK;
}

// small RAII helpers for the chunk pipeline of encode_icm_single
struct StreamBox {
  cudaStream_t s = nullptr;
  ~StreamBox() { if (s) cudaStreamDestroy(s); }
  int create() { RYL_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)); return RAYUELA_OK; }
};
struct EventBox {
  cudaEvent_t e = nullptr;
  ~EventBox() { if (e) cudaEventDestroy(e); }
  int create() { RYL_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); return RAYUELA_OK; }
};

// One device, one call: everything rayuela_encode_icm documents.  snap_sums (host, n_snap doubles, may be null)
// receives the SUM of the per-vector costs at each snapshot, so a multi-device caller can form the global mean.
static int encode_icm_single(const float* X, const float* C, uint8_t* B, int64_t n, int d, int m, int ilsiter,
                             int icmiter, int npert, int randord, uint64_t seed, int64_t g0, const int* orders,
                             const int* snap_iters, int n_snap, uint8_t* B_snap, float* objs, double* snap_sums,
                             float* cost_out, int* stats, unsigned flags, cudaStream_t s) {
  const bool dev = flags & RAYUELA_DEVICE_PTRS;
  const int mh = m * kH;
  RYL_ARG(!dev || d % 4 != 0 || ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(C)) & 15) == 0,
          "encode_icm: device arrays X and C must be 16-byte aligned when d % 4 == 0 (rows are read with 16-byte loads)");

  // chunking of the base set: the unary buffer (m KB per vector) stays within budget (nsplits of
  // src/LSQ_GPU.jl:226-255, done inside the library); gridDim.y of K1 caps a chunk at 65535 * 128 vectors; with HOST
  // arrays the base is cut in >= 4 chunks so that the upload of chunk c+1 overlaps the kernels of chunk c
  const int64_t per_vec = (int64_t)mh * sizeof(float);
  int64_t chunk = std::max<int64_t>(1024, std::min<int64_t>(n, (int64_t)(unary_budget_bytes() / per_vec)));
  chunk = std::min<int64_t>(chunk, (int64_t)65535 * 128);
  if (!dev && n >= 4 * 65536 && !env_off("RAYUELA_B200_ICM_OVERLAP"))
    chunk = std::min<int64_t>(chunk, ((n + 3) / 4 + 1023) / 1024 * 1024);
  if (const char* e = getenv("RAYUELA_B200_ICM_CHUNKS"))      // tuning knob: force a chunk count
    if (atoi(e) >= 1) chunk = std::min<int64_t>(chunk, ((n + atoi(e) - 1) / atoi(e) + 1023) / 1024 * 1024);
  const int nchunks = (int)((n + chunk - 1) / chunk);
  const bool piped = nchunks > 1;

  // streams: chunks alternate between the caller's stream and s_alt (the tail of chunk c overlaps the head of chunk
  // c+1; chunks c and c+2 share a stream and a unary buffer, so the reuse is ordered for free); uploads on s_copy
  StreamBox alt_box, copy_box;
  EventBox ev_setup, ev_alt_done;
  std::vector<EventBox> ev_up(piped && !dev ? nchunks : 0);
  cudaStream_t s_alt = s, s_copy = s;
  if (piped) {
    RYL_TRY(alt_box.create());
    s_alt = alt_box.s;
    RYL_TRY(ev_setup.create());
    RYL_TRY(ev_alt_done.create());
    if (!dev) {
      RYL_TRY(copy_box.create());
      s_copy = copy_box.s;
      for (auto& e : ev_up) RYL_TRY(e.create());
    }
  }

  // CUDA-event timers of the phases, kept when `stats` is requested (rayuela_encode_icm_timings)
  EventBox t_begin, t_setup, t_end;
  std::vector<EventBox> t_u0(stats ? nchunks : 0), t_u1(stats ? nchunks : 0), t_k(stats ? nchunks : 0);
  auto timed_event = [](EventBox& e) -> int {
    RYL_CUDA(cudaEventCreate(&e.e));
    return RAYUELA_OK;
  };
  if (stats) {
    RYL_TRY(timed_event(t_begin));
    RYL_TRY(timed_event(t_setup));
    RYL_TRY(timed_event(t_end));
    for (int c = 0; c < nchunks; c++) {
      RYL_TRY(timed_event(t_u0[c]));
      RYL_TRY(timed_event(t_u1[c]));
      RYL_TRY(timed_event(t_k[c]));
    }
    RYL_CUDA(cudaEventRecord(t_begin.e, s));
  }

  InArg<float> x_in, c_in;
  if (dev || !piped) {
    RYL_TRY(x_in.bind(X, (size_t)n * d, dev, s));
  } else {
    RYL_TRY(x_in.own.alloc((size_t)n * d * sizeof(float), s));
    x_in.d = x_in.own.as<float>();
  }
  RYL_TRY(c_in.bind(C, (size_t)mh * d, dev, s));
  OutArg<uint8_t> b_io, snap_out;
  RYL_TRY(b_io.bind(B, (size_t)n * m, dev, s, /*copy_in=*/true));
  RYL_TRY(snap_out.bind(n_snap ? B_snap : nullptr, (size_t)n_snap * n * m, dev, s));
  DevBuf snap_tmp;  // objs wanted without B_snap: keep the snapshots on the device only
  if (n_snap && !snap_out.d && (objs || snap_sums)) {
    RYL_TRY(snap_tmp.alloc((size_t)n_snap * n * m, s));
    snap_out.d = snap_tmp.as<uint8_t>();
  }
  OutArg<float> cost_o;
  RYL_TRY(cost_o.bind(cost_out, (size_t)n, dev, s));

  // visiting orders (src/LSQ.jl:210-221), snapshot list, stats
  std::vector<int> ord((size_t)std::max(ilsiter, 1) * m);
  for (int it = 0; it < ilsiter; it++) {
    if (orders) memcpy(&ord[(size_t)it * m], orders + (size_t)it * m, sizeof(int) * m);
    else if (randord) philox_randperm(seed, it, m, &ord[(size_t)it * m]);
    else for (int i = 0; i < m; i++) ord[(size_t)it * m + i] = i;
    for (int i = 0; i < m; i++)
      RYL_ARG(ord[(size_t)it * m + i] >= 0 && ord[(size_t)it * m + i] < m, "encode_icm: order entry out of range");
  }
  std::vector<unsigned long long> ordp((size_t)std::max(ilsiter, 1), 0ull);
  for (int it = 0; it < ilsiter; it++)
    for (int i = 0; i < m; i++) ordp[it] |= (unsigned long long)ord[(size_t)it * m + i] << (4 * i);
  DevBuf ord_d, ordp_d, snapit_d, stats_d, nrm_d, T_d;
  RYL_TRY(ord_d.alloc(ord.size() * sizeof(int), s));
  RYL_CUDA(cudaMemcpyAsync(ord_d.p, ord.data(), ord.size() * sizeof(int), cudaMemcpyHostToDevice, s));
  RYL_TRY(ordp_d.alloc(ordp.size() * sizeof(unsigned long long), s));
  RYL_CUDA(cudaMemcpyAsync(ordp_d.p, ordp.data(), ordp.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, s));
  RYL_TRY(snapit_d.alloc((size_t)std::max(n_snap, 1) * sizeof(int), s));
  if (n_snap) RYL_CUDA(cudaMemcpyAsync(snapit_d.p, snap_iters, n_snap * sizeof(int), cudaMemcpyHostToDevice, s));
  RYL_TRY(stats_d.alloc((size_t)std::max(ilsiter, 1) * 2 * sizeof(int) + 16, s));   // + 16: executed / exact-path step counters
  RYL_CUDA(cudaMemsetAsync(stats_d.p, 0, stats_d.bytes, s));
  unsigned long long* steps_d = reinterpret_cast<unsigned long long*>(stats_d.as<int>() + (size_t)std::max(ilsiter, 1) * 2);

  // K0 + K2: ||c||^2 and the pairwise tables (get_binaries + binaries_t, src/LSQ.jl:288,180-183)
  RYL_TRY(nrm_d.alloc((size_t)mh * sizeof(float), s));
  RYL_LAUNCH(sqnorm_kernel, (mh + 255) / 256, 256, 0, s, c_in.d, d, mh, nrm_d.as<float>());
  RYL_TRY(T_d.alloc((size_t)m * m * kH * kH * sizeof(float), s));
  if (m > 1) RYL_LAUNCH(tables_kernel, dim3(kH / 32, kH / 32, m * m), 256, 0, s, c_in.d, T_d.as<float>(), d, m);
  // K2q: quantised copy for the pre-filter
  const bool pf = !env_off("RAYUELA_B200_ICM_PF");
  DevBuf Tq_d, tmax_d, pfc_d;
  if (pf) {
    RYL_TRY(Tq_d.alloc((size_t)m * m * kH * (kH / 2) * sizeof(uint32_t), s));
    RYL_TRY(tmax_d.alloc((size_t)m * sizeof(unsigned int), s));
    RYL_CUDA(cudaMemsetAsync(tmax_d.p, 0, tmax_d.bytes, s));
    if (m > 1) {
      RYL_LAUNCH(tmax_kernel, dim3(8, m * m), 256, 0, s, T_d.as<float>(), m, tmax_d.as<unsigned int>());
      RYL_LAUNCH(quant_tables_kernel, dim3(32, 1, m * m), 256, 0, s, T_d.as<float>(), m, tmax_d.as<unsigned int>(),
                 Tq_d.as<uint32_t>());
    }
    RYL_TRY(pfc_d.alloc((size_t)m * sizeof(float2), s));
    RYL_LAUNCH(pf_consts_kernel, 1, 32, 0, s, tmax_d.as<unsigned int>(), m, pfc_d.as<float2>());
  }

  // opt-in tensor-core unaries (RAYUELA_FAST_UNARIES): codebooks are split / packed once per call
  const char* fast_env = getenv("RAYUELA_B200_FAST_UNARIES");
  const bool fast = ((flags & RAYUELA_FAST_UNARIES) || (fast_env && atoi(fast_env) != 0)) && unary_tc_supported(d, mh);
  DevBuf Cp_d;
  if (fast) RYL_TRY(unary_tc_pack_codebooks(c_in.d, d, mh, &Cp_d, s));
  const int nbuf = piped ? 2 : 1;
  DevBuf U_d[2], umax_d[2], next_d, draws_d[2];
  const int nblk = (npert + 3) / 4;
  const bool predraw = ilsiter > 0 && npert > 0;
  RYL_TRY(next_d.alloc((size_t)nchunks * sizeof(unsigned long long), s));
  RYL_CUDA(cudaMemsetAsync(next_d.p, 0, next_d.bytes, s));
  for (int i = 0; i < nbuf; i++) {
    RYL_TRY(U_d[i].alloc((size_t)std::min(chunk, n) * per_vec, s));
    if (pf) RYL_TRY(umax_d[i].alloc((size_t)std::min(chunk, n) * sizeof(unsigned int), s));
    if (predraw) RYL_TRY(draws_d[i].alloc((size_t)std::min(chunk, n) * ilsiter * nblk * sizeof(unsigned long long), s));
  }
  if (stats) RYL_CUDA(cudaEventRecord(t_setup.e, s));
  if (piped) {
    RYL_CUDA(cudaEventRecord(ev_setup.e, s));               // allocations + tables are ordered before the other streams
    RYL_CUDA(cudaStreamWaitEvent(s_alt, ev_setup.e, 0));
    if (!dev) RYL_CUDA(cudaStreamWaitEvent(s_copy, ev_setup.e, 0));
  }
  // upload of chunk c on the copy stream.  Chunk c+1 is enqueued right AFTER chunk c's kernels: with pinned host memory
  // the order would not matter, but Julia arrays are pageable and a pageable cudaMemcpyAsync blocks the host while it
  // stages -- this way that staging overlaps the kernels already launched
  auto upload_chunk = [&](int c) -> int {
    const int64_t l0 = (int64_t)c * chunk, nc = std::min(chunk, n - l0);
    RYL_CUDA(cudaMemcpyAsync(const_cast<float*>(x_in.d) + (size_t)l0 * d, X + (size_t)l0 * d,
                             (size_t)nc * d * sizeof(float), cudaMemcpyHostToDevice, s_copy));
    RYL_CUDA(cudaEventRecord(ev_up[c].e, s_copy));
    return RAYUELA_OK;
  };
  if (piped && !dev) RYL_TRY(upload_chunk(0));
  for (int c = 0; c < nchunks; c++) {
    const int64_t l0 = (int64_t)c * chunk;
    const int64_t nc = std::min(chunk, n - l0);
    cudaStream_t cs = (c & 1) ? s_alt : s;
    float* U = U_d[c & (nbuf - 1)].as<float>();
    unsigned int* umax = umax_d[c & (nbuf - 1)].as<unsigned int>();
    if (piped && !dev) RYL_CUDA(cudaStreamWaitEvent(cs, ev_up[c].e, 0));
    if (pf) RYL_CUDA(cudaMemsetAsync(umax, 0, (size_t)nc * sizeof(unsigned int), cs));
    if (stats) RYL_CUDA(cudaEventRecord(t_u0[c].e, cs));
    if (fast)
      RYL_TRY(unary_tc_launch(x_in.d + (size_t)l0 * d, Cp_d, nrm_d.as<float>(), U, umax, nc, d, mh, cs));
    else
      RYL_TRY(launch_unary(c_in.d, x_in.d + (size_t)l0 * d, nrm_d.as<float>(), U, nc, d, mh, umax, cs));
    if (stats) RYL_CUDA(cudaEventRecord(t_u1[c].e, cs));
    unsigned long long* draws = draws_d[c & (nbuf - 1)].as<unsigned long long>();
    if (predraw)
      RYL_LAUNCH(perturb_draws_kernel, (int)std::min<int64_t>((nc * ilsiter * nblk + 255) / 256, (int64_t)sm_count() * 16), 256,
                 0, cs, draws, nc, g0 + l0, ilsiter, nblk, m, seed);
    IcmParams p;
    p.U = U;
    p.T = T_d.as<float>();
    p.Tq = pf ? Tq_d.as<uint32_t>() : nullptr;
    p.pfc = pfc_d.as<float2>();
    p.umax = umax;
    p.X = x_in.d + (size_t)l0 * d;
    p.C = c_in.d;
    p.B = b_io.d + (size_t)l0 * m;
    p.cost = cost_o.d ? cost_o.d + l0 : nullptr;
    p.orders = ord_d.as<int>();
    p.orders_packed = ordp_d.as<unsigned long long>();
    p.one = 1u;
    p.draws = predraw ? draws : nullptr;
    p.snap_iters = snapit_d.as<int>();
    p.B_snap = snap_out.d ? snap_out.d + (size_t)l0 * m : nullptr;
    p.stats = stats_d.as<int>();
    p.steps = steps_d;
    p.next = next_d.as<unsigned long long>() + c;
    p.nc = nc;
    p.n_total = n;
    p.g0 = g0 + l0;
    p.seed = seed;
    p.d = d;
    p.ilsiter = ilsiter;
    p.icmiter = icmiter;
    p.npert = npert;
    p.n_snap = snap_out.d ? n_snap : 0;
    int rc = RAYUELA_OK;
#define CALL(M) rc = launch_icm<M>(p, cs)
    RYL_M_SWITCH(m, CALL)
#undef CALL
    RYL_TRY(rc);
    if (stats) RYL_CUDA(cudaEventRecord(t_k[c].e, cs));
    if (piped && !dev && c + 1 < nchunks) RYL_TRY(upload_chunk(c + 1));
  }
  if (piped) {                                              // join: everything below is ordered after both streams
    RYL_CUDA(cudaEventRecord(ev_alt_done.e, s_alt));
    RYL_CUDA(cudaStreamWaitEvent(s, ev_alt_done.e, 0));
  }

  // objective at each snapshot: qerror(RX, B, C), src/LSQ_GPU.jl:197
  if (n_snap && (objs || snap_sums) && snap_out.d) {
    DevBuf tmp;
    RYL_TRY(tmp.alloc((size_t)n * sizeof(float), s));
    for (int si = 0; si < n_snap; si++) {
      RYL_TRY(device_veccost(x_in.d, snap_out.d + (size_t)si * n * m, c_in.d, n, d, m, tmp.as<float>(), s));
      double sum = 0;
      RYL_TRY(device_sum(tmp.as<float>(), n, &sum, s));
      if (objs) objs[si] = (float)(sum / (double)n);
      if (snap_sums) snap_sums[si] = sum;
    }
  }
  RYL_TRY(b_io.flush(s));
  RYL_TRY(snap_out.flush(s));
  RYL_TRY(cost_o.flush(s));
  if (stats) {
    unsigned long long done_steps[2] = {0, 0};
    RYL_CUDA(cudaMemcpyAsync(stats, stats_d.p, (size_t)ilsiter * 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
    RYL_CUDA(cudaMemcpyAsync(done_steps, steps_d, sizeof(done_steps), cudaMemcpyDeviceToHost, s));
    RYL_CUDA(cudaEventRecord(t_end.e, s));
    RYL_CUDA(cudaStreamSynchronize(s));
    g_icm_steps_done = done_steps[0];
    g_icm_steps_exact = pf ? done_steps[1] : done_steps[0];
    g_icm_steps_total = (uint64_t)n * ilsiter * icmiter * m;
    // phase times: chunks on alternating streams overlap, so unaries + ICM can exceed the whole-call figure
    float t = 0;
    g_icm_ms[1] = g_icm_ms[2] = 0;
    cudaEventElapsedTime(&g_icm_ms[0], t_begin.e, t_setup.e);
    for (int c = 0; c < nchunks; c++) {
      if (cudaEventElapsedTime(&t, t_u0[c].e, t_u1[c].e) == cudaSuccess) g_icm_ms[1] += t;
      if (cudaEventElapsedTime(&t, t_u1[c].e, t_k[c].e) == cudaSuccess) g_icm_ms[2] += t;
    }
    cudaEventElapsedTime(&g_icm_ms[3], t_begin.e, t_end.e);
  }
  if (!dev) RYL_CUDA(cudaStreamSynchronize(s));
  return RAYUELA_OK;
}

static int device_veccost_generic(const float* X, const uint8_t* B, const float* C, int64_t n, int d, int m, int h,
                                  float* cost, cudaStream_t s) {
  const int warps = 8;
  const size_t smem = (size_t)warps * ((d + 3) & ~3) * sizeof(float);
  RYL_ARG(smem <= 200 * 1024, "veccost: d too large for shared memory");
  RYL_CUDA(cudaFuncSetAttribute(veccost_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)std::min<int64_t>((n + warps - 1) / warps, (int64_t)sm_count() * 8);
  RYL_LAUNCH(veccost_generic_kernel, grid, warps * 32, smem, s, X, B, C, n, d, m, h, cost);
  return RAYUELA_OK;
}

// h != 256 (1..255): the reference's cpp=false path, iterated_conditional_modes! (src/LSQ.jl:83-149)
static int encode_icm_generic(const float* X, const float* C, uint8_t* B, int64_t n, int d, int m, int h, int ilsiter,
                              int icmiter, int npert, int randord, uint64_t seed, int64_t g0, const int* orders,
                              const int* snap_iters, int n_snap, uint8_t* B_snap, float* objs, float* cost_out,
                              int* stats, unsigned flags, cudaStream_t s) {
  const bool dev = flags & RAYUELA_DEVICE_PTRS;
  const int mh = m * h;
  InArg<float> x_in, c_in;
  RYL_TRY(x_in.bind(X, (size_t)n * d, dev, s));
  RYL_TRY(c_in.bind(C, (size_t)mh * d, dev, s));
  OutArg<uint8_t> b_io, snap_out;
  RYL_TRY(b_io.bind(B, (size_t)n * m, dev, s, /*copy_in=*/true));
  RYL_TRY(snap_out.bind(n_snap ? B_snap : nullptr, (size_t)n_snap * n * m, dev, s));
  DevBuf snap_tmp;
  if (n_snap && !snap_out.d && objs) {
    RYL_TRY(snap_tmp.alloc((size_t)n_snap * n * m, s));
    snap_out.d = snap_tmp.as<uint8_t>();
  }
  OutArg<float> cost_o;
  RYL_TRY(cost_o.bind(cost_out, (size_t)n, dev, s));
  std::vector<int> ord((size_t)std::max(ilsiter, 1) * m);
  for (int it = 0; it < ilsiter; it++) {
    if (orders) memcpy(&ord[(size_t)it * m], orders + (size_t)it * m, sizeof(int) * m);
    else if (randord) philox_randperm(seed, it, m, &ord[(size_t)it * m]);
    else for (int i = 0; i < m; i++) ord[(size_t)it * m + i] = i;
    for (int i = 0; i < m; i++)
      RYL_ARG(ord[(size_t)it * m + i] >= 0 && ord[(size_t)it * m + i] < m, "encode_icm: order entry out of range");
  }
  DevBuf ord_d, snapit_d, stats_d, nrm_d, T_d, U_d;
  RYL_TRY(ord_d.alloc(ord.size() * sizeof(int), s));
  RYL_CUDA(cudaMemcpyAsync(ord_d.p, ord.data(), ord.size() * sizeof(int), cudaMemcpyHostToDevice, s));
  RYL_TRY(snapit_d.alloc((size_t)std::max(n_snap, 1) * sizeof(int), s));
  if (n_snap) RYL_CUDA(cudaMemcpyAsync(snapit_d.p, snap_iters, n_snap * sizeof(int), cudaMemcpyHostToDevice, s));
  RYL_TRY(stats_d.alloc((size_t)std::max(ilsiter, 1) * 2 * sizeof(int) + 16, s));
  RYL_CUDA(cudaMemsetAsync(stats_d.p, 0, stats_d.bytes, s));
  unsigned long long* steps_d = reinterpret_cast<unsigned long long*>(stats_d.as<int>() + (size_t)std::max(ilsiter, 1) * 2);
  RYL_TRY(nrm_d.alloc((size_t)mh * sizeof(float), s));
  RYL_LAUNCH(sqnorm_kernel, (mh + 255) / 256, 256, 0, s, c_in.d, d, mh, nrm_d.as<float>());
  RYL_TRY(T_d.alloc((size_t)m * m * h * h * sizeof(float), s));
  if (m > 1) RYL_LAUNCH(tables_generic_kernel, sm_count() * 8, 256, 0, s, c_in.d, T_d.as<float>(), d, m, h);
  const int64_t per_vec = (int64_t)mh * sizeof(float);
  const int64_t chunk = std::max<int64_t>(1024, std::min<int64_t>(n, (int64_t)(unary_budget_bytes() / per_vec)));
  RYL_TRY(U_d.alloc((size_t)std::min(chunk, n) * per_vec, s));
  const int warps = 8;
  const size_t smem = (size_t)warps * ((d + 3) & ~3) * sizeof(float) + (size_t)2 * ilsiter * sizeof(int);
  RYL_ARG(smem <= 200 * 1024 && (size_t)d * sizeof(float) <= 200 * 1024, "encode_icm: d too large for shared memory");
  RYL_CUDA(cudaFuncSetAttribute(icm_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  RYL_CUDA(cudaFuncSetAttribute(unary_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(d * sizeof(float))));
  for (int64_t l0 = 0; l0 < n; l0 += chunk) {
    const int64_t nc = std::min(chunk, n - l0);
    RYL_LAUNCH(unary_generic_kernel, (int)std::min<int64_t>(nc, (int64_t)sm_count() * 8), 256, d * sizeof(float), s, c_in.d,
               x_in.d + (size_t)l0 * d, nrm_d.as<float>(), U_d.as<float>(), nc, d, mh);
    IcmGenericParams gp;
    IcmParams& p = gp.p;
    p.U = U_d.as<float>();
    p.T = T_d.as<float>();
    p.Tq = nullptr;
    p.pfc = nullptr;
    p.umax = nullptr;
    p.X = x_in.d + (size_t)l0 * d;
    p.C = c_in.d;
    p.B = b_io.d + (size_t)l0 * m;
    p.cost = cost_o.d ? cost_o.d + l0 : nullptr;
    p.orders = ord_d.as<int>();
    p.snap_iters = snapit_d.as<int>();
    p.B_snap = snap_out.d ? snap_out.d + (size_t)l0 * m : nullptr;
    p.stats = stats_d.as<int>();
    p.steps = steps_d;
    p.next = nullptr;
    p.draws = nullptr;
    p.nc = nc;
    p.n_total = n;
    p.g0 = g0 + l0;
    p.seed = seed;
    p.d = d;
    p.ilsiter = ilsiter;
    p.icmiter = icmiter;
    p.npert = npert;
    p.n_snap = snap_out.d ? n_snap : 0;
    gp.m = m;
    gp.h = h;
    const int grid = (int)std::min<int64_t>((nc + warps - 1) / warps, (int64_t)sm_count() * 4);
    RYL_LAUNCH(icm_generic_kernel, grid, warps * 32, smem, s, gp);
  }
  if (n_snap && objs && snap_out.d) {
    DevBuf tmp;
    RYL_TRY(tmp.alloc((size_t)n * sizeof(float), s));
    for (int si = 0; si < n_snap; si++) {
      RYL_TRY(device_veccost_generic(x_in.d, snap_out.d + (size_t)si * n * m, c_in.d, n, d, m, h, tmp.as<float>(), s));
      double mean = 0;
      RYL_TRY(device_mean(tmp.as<float>(), n, &mean, s));
      objs[si] = (float)mean;
    }
  }
  RYL_TRY(b_io.flush(s));
  RYL_TRY(snap_out.flush(s));
  RYL_TRY(cost_o.flush(s));
  if (stats) {
    unsigned long long done_steps[2] = {0, 0};
    RYL_CUDA(cudaMemcpyAsync(stats, stats_d.p, (size_t)ilsiter * 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
    RYL_CUDA(cudaMemcpyAsync(done_steps, steps_d, sizeof(done_steps), cudaMemcpyDeviceToHost, s));
    RYL_CUDA(cudaStreamSynchronize(s));
    g_icm_steps_done = done_steps[0];
    g_icm_steps_exact = done_steps[1];
    g_icm_steps_total = (uint64_t)n * ilsiter * icmiter * m;
  }
  if (!dev) RYL_CUDA(cudaStreamSynchronize(s));
  return RAYUELA_OK;
}

// Host arrays + a configured device set (rayuela_init / RAYUELA_B200_DEVICES): the base is cut in contiguous
// splitarray slices (src/utils.jl:179-203 -- the rule of the reference's serial `nsplits`, src/LSQ_GPU.jl:238-255),
// one per device slot, encoded concurrently by one host thread per slot.  Vectors are independent and the RNG is
// keyed on the global index g0 + l, so the codes are bit-identical to the single-device call; no collective.
static int encode_icm_multi(const std::vector<DeviceSlot>& slots, const float* X, const float* C, uint8_t* B, int64_t n,
                            int d, int m, int ilsiter, int icmiter, int npert, int randord, uint64_t seed, int64_t g0,
                            const int* orders, const int* snap_iters, int n_snap, uint8_t* B_snap, float* objs,
                            float* cost_out, int* stats, unsigned fast_flag) {
  const int D = (int)slots.size();
  std::vector<std::vector<uint8_t>> snaps(D);
  std::vector<std::vector<double>> sums(D, std::vector<double>((size_t)std::max(n_snap, 1), 0.0));
  std::vector<std::vector<int>> st(D, std::vector<int>((size_t)std::max(ilsiter, 1) * 2, 0));
  std::vector<uint64_t> done(D, 0), exact(D, 0);
  RYL_TRY(for_each_slot(slots, [&](int i) -> int {
    int64_t a, b;
    split_range(n, D, i, &a, &b);
    const int64_t ni = b - a;
    if (ni == 0) return RAYUELA_OK;
    if (n_snap && B_snap) snaps[i].resize((size_t)n_snap * ni * m);
    RYL_TRY(encode_icm_single(X + (size_t)a * d, C, B + (size_t)a * m, ni, d, m, ilsiter, icmiter, npert, randord, seed,
                              g0 + a, orders, snap_iters, n_snap, (n_snap && B_snap) ? snaps[i].data() : nullptr, nullptr,
                              (n_snap && objs) ? sums[i].data() : nullptr, cost_out ? cost_out + a : nullptr,
                              stats ? st[i].data() : nullptr, fast_flag, slots[i].stream));
    if (stats) {
      done[i] = g_icm_steps_done;        // this worker thread's counters
      exact[i] = g_icm_steps_exact;
    }
    return RAYUELA_OK;
  }));
  for (int i = 0; i < D; i++) {
    int64_t a, b;
    split_range(n, D, i, &a, &b);
    if (n_snap && B_snap)
      for (int si = 0; si < n_snap; si++)
        memcpy(B_snap + ((size_t)si * n + a) * m, snaps[i].data() + (size_t)si * (b - a) * m, (size_t)(b - a) * m);
  }
  if (n_snap && objs)
    for (int si = 0; si < n_snap; si++) {
      double t = 0;
      for (int i = 0; i < D; i++) t += sums[i][si];
      objs[si] = (float)(t / (double)n);
    }
  if (stats) {
    g_icm_steps_done = g_icm_steps_exact = 0;
    for (int i = 0; i < 2 * ilsiter; i++) stats[i] = 0;
    for (int i = 0; i < D; i++) {
      for (int t = 0; t < 2 * ilsiter; t++) stats[t] += st[i][t];
      g_icm_steps_done += done[i];
      g_icm_steps_exact += exact[i];
    }
    g_icm_steps_total = (uint64_t)n * ilsiter * icmiter * m;
  }
  return RAYUELA_OK;
}

extern "C" int rayuela_encode_icm(const float* X, const float* C, uint8_t* B, int64_t n, int d, int m, int h,
                                  int ilsiter, int icmiter, int npert, int randord, uint64_t seed, int64_t g0,
                                  const int* orders, const int* snap_iters, int n_snap, uint8_t* B_snap,
                                  float* objs, float* cost_out, int* stats, unsigned flags, void* stream) {
  RYL_ARG(h >= 1 && h <= kH, "encode_icm: h must be in 1..256 (codes are bytes below the boundary)");
  RYL_ARG(m >= 1 && m <= 16, "encode_icm: m must be in 1..16");
  RYL_ARG(n >= 0 && d >= 1, "encode_icm: bad n or d");
  RYL_ARG(ilsiter >= 0 && icmiter >= 0 && npert >= 0, "encode_icm: negative iteration count");
  RYL_ARG(n_snap >= 0 && (n_snap == 0 || snap_iters), "encode_icm: snap_iters missing");
  RYL_ARG(X && C && B, "encode_icm: null array");
  if (n == 0) return RAYUELA_OK;
  if (h != kH)   // the reference's cpp=false path (src/LSQ.jl:83-149): any h, plain exact kernels, one device
    return encode_icm_generic(X, C, B, n, d, m, h, ilsiter, icmiter, npert, randord, seed, g0, orders, snap_iters, n_snap,
                              B_snap, objs, cost_out, stats, flags, (cudaStream_t)stream);
  if (!(flags & RAYUELA_DEVICE_PTRS)) {
    const std::vector<DeviceSlot> slots = device_slots();
    if (slots.size() > 1 && n >= (int64_t)slots.size() * 1024)
      return encode_icm_multi(slots, X, C, B, n, d, m, ilsiter, icmiter, npert, randord, seed, g0, orders, snap_iters,
                              n_snap, B_snap, objs, cost_out, stats, flags & RAYUELA_FAST_UNARIES);
  }
  return encode_icm_single(X, C, B, n, d, m, ilsiter, icmiter, npert, randord, seed, g0, orders, snap_iters, n_snap,
                           B_snap, objs, nullptr, cost_out, stats, flags, (cudaStream_t)stream);
}

extern "C" int rayuela_get_unaries(const float* X, const float* C, int64_t n, int d, int m, int h, float* U,
                                   unsigned flags, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  RYL_ARG(h == kH && m >= 1 && m <= 16 && n >= 1 && d >= 1, "get_unaries: bad shape (h must be 256, m in 1..16)");
  RYL_ARG(X && C && U, "get_unaries: null array");
  RYL_ARG(n <= (int64_t)65535 * 128, "get_unaries: at most 8388480 vectors per call");
  const bool dev = flags & RAYUELA_DEVICE_PTRS;
  RYL_ARG(!dev || d % 4 != 0 || ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(C)) & 15) == 0,
          "get_unaries: device arrays X and C must be 16-byte aligned when d % 4 == 0");
  const int mh = m * kH;
  InArg<float> x_in, c_in;
  RYL_TRY(x_in.bind(X, (size_t)n * d, dev, s));
  RYL_TRY(c_in.bind(C, (size_t)mh * d, dev, s));
  OutArg<float> u_out;
  RYL_TRY(u_out.bind(U, (size_t)n * mh, dev, s));
  DevBuf nrm_d, Cp_d;
  RYL_TRY(nrm_d.alloc((size_t)mh * sizeof(float), s));
  RYL_LAUNCH(sqnorm_kernel, (mh + 255) / 256, 256, 0, s, c_in.d, d, mh, nrm_d.as<float>());
  const char* fast_env = getenv("RAYUELA_B200_FAST_UNARIES");
  const bool fast = ((flags & RAYUELA_FAST_UNARIES) || (fast_env && atoi(fast_env) != 0)) && unary_tc_supported(d, mh);
  if (fast) {
    RYL_TRY(unary_tc_pack_codebooks(c_in.d, d, mh, &Cp_d, s));
    RYL_TRY(unary_tc_launch(x_in.d, Cp_d, nrm_d.as<float>(), u_out.d, nullptr, n, d, mh, s));
  } else {
    RYL_TRY(launch_unary(c_in.d, x_in.d, nrm_d.as<float>(), u_out.d, n, d, mh, nullptr, s));
  }
  RYL_TRY(u_out.flush(s));
  if (!dev) RYL_CUDA(cudaStreamSynchronize(s));
  return RAYUELA_OK;
}

extern "C" int rayuela_veccost(const float* X, const uint8_t* B, const float* C, int64_t n, int d, int m, int h,
                               float* cost, double* mean_out, unsigned flags, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  RYL_ARG(h >= 1 && h <= kH && m >= 1 && m <= 16 && n >= 1 && d >= 1, "veccost: bad shape (h in 1..256, m in 1..16)");
  const bool dev = flags & RAYUELA_DEVICE_PTRS;
  RYL_ARG(!dev || d % 4 != 0 || ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(C)) & 15) == 0,
          "veccost: device arrays X and C must be 16-byte aligned when d % 4 == 0");
  InArg<float> x_in, c_in;
  InArg<uint8_t> b_in;
  RYL_TRY(x_in.bind(X, (size_t)n * d, dev, s));
  RYL_TRY(c_in.bind(C, (size_t)m * h * d, dev, s));
  RYL_TRY(b_in.bind(B, (size_t)n * m, dev, s));
  OutArg<float> cost_o;
  DevBuf tmp;
  float* cd = nullptr;
  if (cost) {
    RYL_TRY(cost_o.bind(cost, (size_t)n, dev, s));
    cd = cost_o.d;
  } else {
    RYL_TRY(tmp.alloc((size_t)n * sizeof(float), s));
    cd = tmp.as<float>();
  }
  if (h == kH) RYL_TRY(device_veccost(x_in.d, b_in.d, c_in.d, n, d, m, cd, s));
  else RYL_TRY(device_veccost_generic(x_in.d, b_in.d, c_in.d, n, d, m, h, cd, s));
  if (mean_out) RYL_TRY(device_mean(cd, n, mean_out, s));
  RYL_TRY(cost_o.flush(s));
  if (!dev) RYL_CUDA(cudaStreamSynchronize(s));
  return RAYUELA_OK;
}


template <int M>
static int launch_norms(const uint8_t* B, const float* C, const float* cbnorms, int64_t n, int d, uint8_t* codes,
                        float* norms, cudaStream_t s) {
  const int warps = 8;
  size_t smem = (size_t)warps * ((d + 3) & ~3) * sizeof(float);
  RYL_ARG(smem <= 200 * 1024, "quantize_norms: d too large for shared memory");
  RYL_CUDA(cudaFuncSetAttribute(norms_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = (int)std::min<int64_t>((n + warps - 1) / warps, (int64_t)sm_count() * 8);
  RYL_LAUNCH(norms_kernel<M>, grid, warps * 32, smem, s, B, C, cbnorms, n, d, codes, norms);
  return RAYUELA_OK;
}

extern "C" int rayuela_quantize_norms(const uint8_t* B, const float* C, const float* cbnorms, int64_t n, int d, int m,
                                      int h, uint8_t* norm_codes, float* norms_out, unsigned flags, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  RYL_ARG(h == kH && m >= 1 && m <= 16 && n >= 1 && d >= 1, "quantize_norms: bad shape (h must be 256, m in 1..16)");
  RYL_ARG(B && C && (norm_codes || norms_out), "quantize_norms: null array");
  RYL_ARG(!norm_codes || cbnorms, "quantize_norms: norm codes need the norms codebook");
  const bool dev = flags & RAYUELA_DEVICE_PTRS;
  InArg<uint8_t> b_in;
  InArg<float> c_in, cb_in;
  RYL_TRY(b_in.bind(B, (size_t)n * m, dev, s));
  RYL_TRY(c_in.bind(C, (size_t)m * kH * d, dev, s));
  RYL_TRY(cb_in.bind(cbnorms, (size_t)kH, dev, s));
  OutArg<uint8_t> codes_o;
  OutArg<float> norms_o;
  RYL_TRY(codes_o.bind(norm_codes, (size_t)n, dev, s));
  RYL_TRY(norms_o.bind(norms_out, (size_t)n, dev, s));
  int rc = RAYUELA_OK;
#define CALL(M) rc = launch_norms<M>(b_in.d, c_in.d, cb_in.d, n, d, codes_o.d, norms_o.d, s)
  RYL_M_SWITCH(m, CALL)
#undef CALL
  RYL_TRY(rc);
  RYL_TRY(codes_o.flush(s));
  RYL_TRY(norms_o.flush(s));
  if (!dev) RYL_CUDA(cudaStreamSynchronize(s));
  return RAYUELA_OK;
}

static int launch_viterbi(const float* U, const float* TT, int64_t tt_stride, int64_t n, int m, uint8_t* B,
                          cudaStream_t s) {
  constexpr int VB = 8;
  size_t smem = (size_t)VB * kH * sizeof(float) + (size_t)VB * std::max(m - 1, 1) * kH;
  RYL_CUDA(cudaFuncSetAttribute(viterbi_kernel<VB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = (int)std::min<int64_t>((n + VB - 1) / VB, (int64_t)sm_count() * 6);
  RYL_LAUNCH(viterbi_kernel<VB>, grid, 256, smem, s, U, TT, tt_stride, n, m, B);
  return RAYUELA_OK;
}

extern "C" int rayuela_quantize_chainq(const float* X, const float* C, int64_t n, int d, int m, int h, uint8_t* B,
                                       unsigned flags, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  RYL_ARG(h == kH, "quantize_chainq: only codebooks with 256 entries are supported (src/ChainQ.jl:18-21)");
  RYL_ARG(m >= 1 && m <= 16 && n >= 1 && d >= 1, "quantize_chainq: bad shape (m in 1..16)");
  RYL_ARG(X && C && B, "quantize_chainq: null array");
  const bool dev = flags & RAYUELA_DEVICE_PTRS;
  const int mh = m * kH;
  InArg<float> x_in, c_in;
  RYL_TRY(x_in.bind(X, (size_t)n * d, dev, s));
  RYL_TRY(c_in.bind(C, (size_t)mh * d, dev, s));
  OutArg<uint8_t> b_out;
  RYL_TRY(b_out.bind(B, (size_t)n * m, dev, s));
  DevBuf nrm_d, T_d, U_d;
  RYL_TRY(nrm_d.alloc((size_t)mh * sizeof(float), s));
  RYL_LAUNCH(sqnorm_kernel, (mh + 255) / 256, 256, 0, s, c_in.d, d, mh, nrm_d.as<float>());
  RYL_TRY(T_d.alloc((size_t)m * m * kH * kH * sizeof(float), s));
  if (m > 1) RYL_LAUNCH(tables_kernel, dim3(kH / 32, kH / 32, m * m), 256, 0, s, c_in.d, T_d.as<float>(), d, m);
  const int64_t per_vec = (int64_t)mh * sizeof(float);
  const int64_t chunk = std::min<int64_t>((int64_t)65535 * 128,   // gridDim.y of K1
      std::max<int64_t>(1024, std::min<int64_t>(n, (int64_t)(unary_budget_bytes() / per_vec))));
  RYL_TRY(U_d.alloc((size_t)std::min(chunk, n) * per_vec, s));
  // chain table i -> i+1 with k (state of i) major: T[(i+1)*m + i][b = k][c = j] = 2<C_{i+1}[:,j], C_i[:,k]>
  const float* TT = T_d.as<float>() + (size_t)(1 * m + 0) * kH * kH;
  const int64_t tt_stride = (int64_t)(m + 1) * kH * kH;
  for (int64_t l0 = 0; l0 < n; l0 += chunk) {
    const int64_t nc = std::min(chunk, n - l0);
    RYL_TRY(launch_unary(c_in.d, x_in.d + (size_t)l0 * d, nrm_d.as<float>(), U_d.as<float>(), nc, d, mh, nullptr, s));
    RYL_TRY(launch_viterbi(U_d.as<float>(), TT, tt_stride, nc, m, b_out.d + (size_t)l0 * m, s));
  }
  RYL_TRY(b_out.flush(s));
  if (!dev) RYL_CUDA(cudaStreamSynchronize(s));
  return RAYUELA_OK;
}

// exact-signature compat for the reference symbol (deps/src/encode_icm.cpp:170-178, called at src/ChainQ.jl:26-28)
extern "C" void viterbi_encoding(unsigned char* B, float* unaries, float* binaries, int n, int m) {
  auto body = [&]() -> int {
    cudaStream_t s = nullptr;
    RYL_ARG(m >= 1 && m <= 16 && n >= 1, "viterbi_encoding: bad shape");
    InArg<float> u_in, b_in;
    RYL_TRY(u_in.bind(unaries, (size_t)n * m * kH, false, s));
    RYL_TRY(b_in.bind(binaries, (size_t)std::max(m - 1, 1) * kH * kH, false, s));
    OutArg<uint8_t> b_out;
    RYL_TRY(b_out.bind(B, (size_t)n * m, false, s));
    DevBuf tt;
    RYL_TRY(tt.alloc((size_t)std::max(m - 1, 1) * kH * kH * sizeof(float), s));
    if (m > 1)
      RYL_LAUNCH(transpose_tables_kernel, dim3(kH / 32, kH / 32, m - 1), dim3(32, 8), 0, s, b_in.d, tt.as<float>(),
                 m - 1);
    RYL_TRY(launch_viterbi(u_in.d, tt.as<float>(), (int64_t)kH * kH, n, m, b_out.d, s));
    RYL_TRY(b_out.flush(s));
    RYL_CUDA(cudaStreamSynchronize(s));
    return RAYUELA_OK;
  };
  int rc = body();
  if (rc != RAYUELA_OK) {
    fprintf(stderr, "librayuela_b200: viterbi_encoding failed (%d): %s\n", rc, rayuela_last_error());
    abort();
  }
}

extern "C" void condition(unsigned char* B, float* ub, float* binaries, float* binaries_t, int* cbpair2binaryidx,
                          int* to_condition, int j, int n, int m) {
  auto body = [&]() -> int {
    cudaStream_t s = nullptr;
    RYL_ARG(m >= 2 && n >= 1 && j >= 0 && j < m, "condition: bad shape");
    const size_t ncbi = (size_t)m * (m - 1) / 2;
    InArg<float> bin, bin_t;
    InArg<int> p2i, tc;
    OutArg<uint8_t> b_io;
    OutArg<float> ub_io;
    RYL_TRY(bin.bind(binaries, ncbi * kH * kH, false, s));
    RYL_TRY(bin_t.bind(binaries_t, ncbi * kH * kH, false, s));
    RYL_TRY(p2i.bind(cbpair2binaryidx, (size_t)m * m, false, s));
    RYL_TRY(tc.bind(to_condition, (size_t)m - 1, false, s));
    RYL_TRY(b_io.bind(B, (size_t)n * m, false, s, true));
    RYL_TRY(ub_io.bind(ub, (size_t)n * kH, false, s, true));
    int grid = (int)std::min<int64_t>(((int64_t)n + 7) / 8, (int64_t)sm_count() * 8);
    RYL_LAUNCH(condition_kernel, grid, 256, 0, s, b_io.d, ub_io.d, bin.d, bin_t.d, p2i.d, tc.d, j, n, m);
    RYL_TRY(b_io.flush(s));
    RYL_TRY(ub_io.flush(s));
    RYL_CUDA(cudaStreamSynchronize(s));
    return RAYUELA_OK;
  };
  int rc = body();
  if (rc != RAYUELA_OK) {
    fprintf(stderr, "librayuela_b200: condition failed (%d): %s\n", rc, rayuela_last_error());
    abort();
  }
}
