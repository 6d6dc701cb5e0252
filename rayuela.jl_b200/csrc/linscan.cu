// linscan.cu -- path (2): asymmetric-distance linear scan on B200.
//   K4 lut_kernel     per-query m*256 lookup table, exact reference arithmetic order
//   K5 scanx_kernel   bank-conflict-free byte-code scan from a shared-memory LUT tile + streaming top-k
//                     (threshold filter, candidate buffers, event-driven block radix-select compaction)
//   K6 merge_kernel   k-way merge of sorted (dist,id) lists (DB slices of one GPU, or per-GPU shards)
// Replaces deps/src/linscan_aqd.cpp:37-102 and deps/src/linscan_aqd_pairwise_byte.cpp:14-176.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

#include "common.cuh"

namespace ryl {

static constexpr int kH = 256;

// ------------------------------------------------------------------------------------------------------
// K4: LUT build.  One thread owns one (query, entry) pair and walks the dimension sequentially with
// UNFUSED fp32 ops in the reference's order (the reference .so is built without FMA contraction):
//   LSQ  t -= (2*q[k])*c[k]         pairwise_byte.cpp:45-47
//   CQ   t += (q[k]-c[k])^2         pairwise_byte.cpp:127-130
//   PQ   t += (c[s]-q[kk*sub+s])^2  linscan_aqd.cpp:66-74
// Tile: 32 entries x 32 queries per 256-thread block, staged through shared memory (entry rows padded to
// an odd stride so the 32 lanes of a warp -- 32 different entries, same t -- hit 32 banks).
// ------------------------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(256) lut_kernel(const float* __restrict__ queries, const float* __restrict__ cb,
                                                  float* __restrict__ lut, int nq, int d, int len, int mh,
                                                  int tiled, int* __restrict__ bad, int h) {
  constexpr int CH = 64;
  __shared__ float cs[32][CH + 1];
  __shared__ float qs[32][CH];
  // blockIdx.x = (codebook kk, 32-entry group inside it): a block never straddles two codebooks, for any h <= 256
  const int gph = (h + 31) >> 5, kk = blockIdx.x / gph, c0 = (blockIdx.x % gph) * 32;
  const int e0 = kk * h + c0, q0 = blockIdx.y * 32;
  const int nvalid = min(32, h - c0);                    // entries of this block that exist
  const int e = threadIdx.x & 31, qg = threadIdx.x >> 5;
  const int qoff = (KIND == RAYUELA_SCAN_PQ) ? kk * len : 0;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int base = 0; base < len; base += CH) {
    const int chunk = min(CH, len - base);
    for (int i = threadIdx.x; i < 32 * CH; i += 256) {
      int r = i / CH, t = i % CH;
      if (t < chunk) {
        cs[r][t] = r < nvalid ? cb[(size_t)(e0 + r) * len + base + t] : 0.f;
        int q = min(q0 + r, nq - 1);
        qs[r][t] = queries[(size_t)q * d + qoff + base + t];
      }
    }
    __syncthreads();
    for (int t = 0; t < chunk; t++) {
      float c = cs[e][t];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        float q = qs[qg + 8 * i][t];
        if (KIND == RAYUELA_SCAN_LSQ) {
          acc[i] = __fsub_rn(acc[i], __fmul_rn(__fmul_rn(2.0f, q), c));
        } else if (KIND == RAYUELA_SCAN_CQ) {
          float df = __fsub_rn(q, c);
          acc[i] = __fadd_rn(acc[i], __fmul_rn(df, df));
        } else {
          float df = __fsub_rn(c, q);
          acc[i] = __fadd_rn(acc[i], __fmul_rn(df, df));
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int q = q0 + qg + 8 * i;
    if (q < nq && e < nvalid) {
      // The scan restarts a lane's accumulator by multiplying it by zero, so one non-finite partial sum would turn
      // every later code of that lane into NaN.  Entries are therefore required to be finite and small enough that a
      // sum of 16 cannot overflow; a violation fails the whole search instead of silently dropping neighbours.
      if (!(fabsf(acc[i]) <= 1e37f)) *bad = 1;
      if (tiled == 1) {
        // layout of scanx_kernel<8>'s shared-memory tile (see there): [q/16][(q%16)/4][c][((q%4)/2)*8 + k][q%2]
        const int k = kk, c = c0 + e;
        lut[(size_t)(q >> 4) * 32768 + ((q & 15) >> 2) * 8192 + c * 32 + ((((q & 3) >> 1) * 8 + k) << 1) + (q & 1)] =
            acc[i];
      } else if (tiled == 2) {
        // scanx_kernel<16>: [q/8][(q%8)/2][c][k][q%2]
        const int k = kk, c = c0 + e;
        lut[(size_t)(q >> 3) * 32768 + ((q & 7) >> 1) * 8192 + c * 32 + (k << 1) + (q & 1)] = acc[i];
      } else {
        lut[(size_t)q * mh + e0 + e] = acc[i];
      }
    }
  }
}

// OPT-IN fast LUT (RAYUELA_FAST_LUT, LSQ only): the LUT is the dense contraction -2 * Q C^T, computed row-major by the
// tcgen05 GEMM of unary_tc.cu (bf16x3) and re-laid-out here into the scan's tiled layout (same index maps as lut_kernel).
__global__ void __launch_bounds__(256) lut_relayout_kernel(const float* __restrict__ rowmajor, float* __restrict__ lut,
                                                           int nq, int mh, int tiled, int* __restrict__ bad) {
  const int64_t total = (int64_t)nq * mh;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(i / mh), ent = (int)(i % mh), k = ent >> 8, c = ent & 255;
    const float v = rowmajor[i];
    if (!(fabsf(v) <= 1e37f)) *bad = 1;
    if (tiled == 1)
      lut[(size_t)(q >> 4) * 32768 + ((q & 15) >> 2) * 8192 + c * 32 + ((((q & 3) >> 1) * 8 + k) << 1) + (q & 1)] = v;
    else
      lut[(size_t)(q >> 3) * 32768 + ((q & 7) >> 1) * 8192 + c * 32 + (k << 1) + (q & 1)] = v;
  }
}

// K4q: quantised twin of one fp32 LUT tile for the pre-filter scan (QPF, see scanx_kernel).  One block per query tile:
//   lo[q][k], hi[q][k] = min / max over the valid entries c < h of codebook k < m
//   s      = max_q (sum_k (hi - lo) + (norm cap - min norm)) / 2000   one scale per tile (1 if that is 0)
//   v      = rint((LUT - lo[q][k]) / s)  (+ 1 for k = 0)              two queries per fp32 word as v_hi * 4096 + v_lo, laid
//            out like the fp32 tile with 4 queries per 8-byte entry:
//            P = 8: [tt2][c][g*8 + k][e2], queries tt2*8 + g*4 + 2*e2 (+1);   P = 16: [tt2][c][k][e2], queries tt2*4 + 2*e2 (+1)
//   off_q  = sum_k lo[q][k] + norm0 (the norm that quantises to 0, <= the minimum);
//   mu_q   = ceil(0.5 (m + 1) + 0.1) + 1 + ceil((m+1) 2^-23 (sum_k max|LUT| + max|norm|) / s)
// so that A = sum_k v + min(rint((norm - norm0)/s), cap) <= 2000 + m/2 + 3 < 2^11 and every code with exact fp32 distance
// E <= tau has A <= floor((tau - off_q)/s) + mu_q: m + 1 roundings of 0.5 (+ the divisions'), 1 for the offset of
// codebook 0, and the fp32 roundings of the exact chain.
template <int P>
__global__ void __launch_bounds__(512) lut_quant_kernel(const float* __restrict__ lut, float* __restrict__ lutq,
                                                        float4* __restrict__ tilep, double* __restrict__ qoff,
                                                        int* __restrict__ qmu, int m, int h, float nmin, float nmax,
                                                        float ncap, int has_norms) {
  constexpr int QB = P == 8 ? 16 : 8;
  extern __shared__ __align__(16) float tile_s[];          // 32768 floats
  __shared__ unsigned int mn_s[16 * 16], mx_s[16 * 16];    // ordered-uint images, [q][k]
  __shared__ double rng_s[16], off_s[16], bmax_s[16];
  __shared__ float s_s;
  __shared__ double koff_s;
  const int tid = threadIdx.x;
  const float4* src = reinterpret_cast<const float4*>(lut + (size_t)blockIdx.x * 32768);
  for (int i = tid; i < 8192; i += 512) reinterpret_cast<float4*>(tile_s)[i] = __ldg(src + i);
  if (tid < 256) {
    mn_s[tid] = 0xFFFFFFFFu;
    mx_s[tid] = 0u;
  }
  __syncthreads();
  auto fidx = [](int q, int k, int c) -> int {
    return P == 8 ? (q >> 2) * 8192 + c * 32 + (((((q & 3) >> 1) << 3) + k) << 1) + (q & 1)
                  : (q >> 1) * 8192 + c * 32 + (k << 1) + (q & 1);
  };
  {
    // thread <-> (column of the 32-float row, tile quarter tt, quarter of the c range): conflict-free column walks
    const int col = tid & 31, tt = (tid >> 5) & 3, part = tid >> 7;
    const int bp = col >> 1, e = col & 1;
    const int k = P == 8 ? (bp & 7) : bp;
    const int q = P == 8 ? tt * 4 + (bp >> 3) * 2 + e : tt * 2 + e;
    if (k < m) {
      float lo = __int_as_float(0x7f800000), hi = -__int_as_float(0x7f800000);
      for (int c = part * 64; c < min(h, part * 64 + 64); c++) {
        const float v = tile_s[tt * 8192 + c * 32 + col];
        lo = fminf(lo, v);
        hi = fmaxf(hi, v);
      }
      if (lo <= hi) {
        atomicMin(&mn_s[q * 16 + k], f32_to_ordered(lo));
        atomicMax(&mx_s[q * 16 + k], f32_to_ordered(hi));
      }
    }
  }
  __syncthreads();
  if (tid < QB) {
    double rng = 0, off = 0, bm = 0;
    for (int k = 0; k < m; k++) {
      const float lo = ordered_to_f32(mn_s[tid * 16 + k]), hi = ordered_to_f32(mx_s[tid * 16 + k]);
      rng += (double)hi - (double)lo;
      off += (double)lo;
      bm += fmax(fabs((double)lo), fabs((double)hi));
    }
    rng_s[tid] = rng;
    off_s[tid] = off;
    bmax_s[tid] = bm;
  }
  __syncthreads();
  // norms are quantised up to ncap only (index_create: mean + 4 sigma, at most the maximum) and SATURATE above it: a
  // saturated code merely looks closer than it is (a spurious survivor at worst), and a few outliers no longer set the scale
  const double nrange = has_norms ? (double)ncap - (double)nmin : 0.0;
  if (tid == 0) {
    double r = 0;
    for (int q = 0; q < QB; q++) r = fmax(r, rng_s[q]);
    float sc = (float)((r + nrange) / 2000.0);
    if (!(sc > 0.f) || !(sc < __int_as_float(0x7f800000))) sc = 1.0f;
    s_s = sc;
    const float inv = __fdiv_rn(1.0f, sc);
    // rint(norm * inv - K) with K = floor(min norm * inv): the magic constant 1.5 * 2^23 - K is an exact integer, so the
    // conversion's only error is its own rint; the norm that quantises to 0 is K / inv (<= the minimum)
    const double K = has_norms ? floor((double)nmin * (double)inv) : 0.0;
    koff_s = K / (double)inv;
    const float cap_units = has_norms ? (float)(ceil((double)ncap * (double)inv - K) + 1.0) : 0.f;
    tilep[blockIdx.x] = make_float4(sc, inv, (float)(12582912.0 - K), cap_units);
  }
  __syncthreads();
  const float sc = s_s;
  if (tid < QB) {
    const double nb = has_norms ? fmax(fabs((double)nmin), fabs((double)nmax)) : 0.0;
    const double fp = ceil((double)(m + 1) * 1.1920928955078125e-07 * (bmax_s[tid] + nb) / (double)sc);
    qoff[(size_t)blockIdx.x * QB + tid] = off_s[tid] + (has_norms ? koff_s : 0.0);
    qmu[(size_t)blockIdx.x * QB + tid] = (int)ceil(0.5 * (m + (has_norms ? 1 : 0)) + 0.1) + 1 + (int)fmin(fp, 40000.0);
  }
  for (int o = tid; o < 16384; o += 512) {
    const int tt2 = o >> 13, c = (o >> 5) & 255, bp = (o >> 1) & 15, e2 = o & 1;
    const int k = P == 8 ? (bp & 7) : bp;
    const int q = P == 8 ? tt2 * 8 + (bp >> 3) * 4 + 2 * e2 : tt2 * 4 + 2 * e2;      // low digit; q + 1 is the high one
    float w = 0.f;
    if (k < m && c < h) {
      float v[2];
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const float f = tile_s[fidx(q + e, k, c)], lo = ordered_to_f32(mn_s[(q + e) * 16 + k]);
        v[e] = fminf(fmaxf(rintf(__fdiv_rn(__fsub_rn(f, lo), sc)), 0.f), 2010.f) + (k == 0 ? 1.f : 0.f);
      }
      w = v[1] * 4096.f + v[0];
    }
    lutq[(size_t)blockIdx.x * 16384 + o] = w;
  }
}

// min / max of the database norms (ordered-uint images; NaN norms are skipped -- such codes can never be returned)
__global__ void __launch_bounds__(256) norm_range_kernel(const float* __restrict__ v, int64_t n, unsigned int* __restrict__ out,
                                                         double* __restrict__ sums) {
  float lo = __int_as_float(0x7f800000), hi = -__int_as_float(0x7f800000);
  double s1 = 0, s2 = 0, cnt = 0;                       // of the finite norms: for the saturation point of the pre-filter
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const float x = v[i];
    lo = fminf(lo, x);
    hi = fmaxf(hi, x);
    if (fabsf(x) < __int_as_float(0x7f800000)) {
      s1 += x;
      s2 += (double)x * x;
      cnt += 1;
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, off);
    s2 += __shfl_xor_sync(0xffffffffu, s2, off);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(sums, s1);
    atomicAdd(sums + 1, s2);
    atomicAdd(sums + 2, cnt);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, off));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, off));
  }
  if ((threadIdx.x & 31) == 0 && lo <= hi) {
    atomicMin(out, f32_to_ordered(lo));
    atomicMax(out + 1, f32_to_ordered(hi));
  }
}

// Block-wide barrier (named barrier 1 with an explicit thread count; same semantics as __syncthreads()).  In
// scanx_kernel every block barrier between the prologue and the final phase lives inside service(), which is ONE
// non-inlined function, so the warps that call it from the period loop and the finished warps that call it from
// their wait loop execute the very same barrier instructions.
__device__ __forceinline__ void block_sync() { asm volatile("bar.sync 1, %0;" ::"r"(blockDim.x) : "memory"); }

// ------------------------------------------------------------------------------------------------------
// Block-wide bitonic sort of np2 (power of two) 64-bit keys in shared memory, ascending.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_bitonic_sort(uint64_t* s, int np2) {
  for (int size = 2; size <= np2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (np2 >> 1); t += blockDim.x) {
        int lo = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
        int hi = lo | stride;
        bool up = (lo & size) == 0;
        uint64_t a = s[lo], b = s[hi];
        if ((a > b) == up) {
          s[lo] = b;
          s[hi] = a;
        }
      }
      block_sync();
    }
  }
}

__device__ __forceinline__ int pow2ceil(int x) {
  int p = 2;
  while (p < x) p <<= 1;
  return p;
}

static constexpr int kChunkCodes = 1024;              // codes per warp chunk
static constexpr int kScanWarps = 16;
static constexpr int kScanSortKeys = 8192;           // shared-memory selection buffer (64 KB)
static constexpr int kLutTileBytes = 131072;          // 16 (m <= 8) or 8 (m <= 16) queries' tables

__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  return ((uint64_t)__float_as_uint(hi) << 32) | __float_as_uint(lo);
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ int lds_volatile(uint32_t addr) {
  int v;
  asm volatile("ld.volatile.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
template <int IMM>
__device__ __forceinline__ uint64_t lds64(uint32_t addr) {
  uint64_t v;
  asm("ld.shared.b64 %0, [%1+%2];" : "=l"(v) : "r"(addr), "n"(IMM));
  return v;
}

// Block-wide radix select over c distinct 64-bit keys in shared memory: returns the k-th smallest key
// (1 <= k <= c <= 8192).  Eight 8-bit passes from the most significant byte, skipping the leading bytes all keys share;
// per pass a 256-bin histogram of the keys that match the prefix decided so far, then EVERY warp finds the bin holding
// rank k by itself (same histogram, same answer) -- so a pass costs ONE block barrier: the histograms rotate through
// three buffers (pass p adds into buffer p % 3 and zeroes buffer (p+1) % 3, whose last readers finished before the
// previous barrier).  Bins are 16-bit counters, two per word (c <= 8192 cannot carry into the neighbour).
// hist: 3 x 128 words.  Replaces a 3-barrier-per-pass version whose serial bin search by warp 0 made the other 15 warps
// wait (ncu r1, k = 1000: 22 % of the samples of the scan kernel sat in this function).
__device__ __forceinline__ uint64_t block_radix_select(const uint64_t* buf, int c, int k, uint32_t* hist) {
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nt = blockDim.x;
  // first byte where any two keys differ: per-warp OR of (key ^ buf[0]) parked in buffer 2, combined by everyone
  uint64_t diff = 0;
  const uint64_t k0 = buf[0];
  for (int t = tid; t < c; t += nt) diff |= buf[t] ^ k0;
  const uint32_t dlo = __reduce_or_sync(0xffffffffu, (uint32_t)diff), dhi = __reduce_or_sync(0xffffffffu, (uint32_t)(diff >> 32));
  if (lane == 0) {
    hist[256 + 2 * w] = dlo;
    hist[256 + 2 * w + 1] = dhi;
  }
  if (tid < 128) hist[tid] = 0;                                  // buffer 0 for the first pass
  block_sync();
  {
    const uint32_t v = hist[256 + lane];                          // 16 warps x 2 words
    const uint32_t lo = __reduce_or_sync(0xffffffffu, (lane & 1) ? 0u : v), hi = __reduce_or_sync(0xffffffffu, (lane & 1) ? v : 0u);
    diff = ((uint64_t)hi << 32) | lo;
  }
  const int first = diff ? (__clzll((long long)diff) >> 3) : 8;
  uint64_t prefix = first ? (k0 >> (64 - 8 * first)) : 0;
  int krem = k;
  for (int pass = first; pass < 8; pass++) {
    const int shift = 56 - 8 * pass;
    uint32_t* h = hist + ((pass - first) % 3) * 128;
    uint32_t* hz = hist + ((pass - first + 1) % 3) * 128;
    if (tid < 128) hz[tid] = 0;
    if (pass == first) {
      // the first discriminating byte usually takes few distinct values: aggregate equal digits per warp
      for (int t0 = 0; t0 < c; t0 += nt) {        // uniform trip count: match.any needs converged warps
        const int t = t0 + tid;
        const bool act = t < c;
        const uint32_t digit = act ? (uint32_t)(buf[t] >> shift) & 255u : 256u + lane;
        const uint32_t peers = __match_any_sync(0xffffffffu, digit);
        if (act && lane == __ffs(peers) - 1) atomicAdd(&h[digit >> 1], (uint32_t)__popc(peers) << (16 * (digit & 1)));
      }
    } else {
      for (int t = tid; t < c; t += nt) {
        const uint64_t key = buf[t];
        if ((key >> (shift + 8)) == prefix) {
          const uint32_t digit = (uint32_t)(key >> shift) & 255u;
          atomicAdd(&h[digit >> 1], 1u << (16 * (digit & 1)));
        }
      }
    }
    block_sync();
    int loc[8], sum = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const uint32_t v = h[lane * 4 + i];
      loc[2 * i] = (int)(v & 0xFFFFu);
      loc[2 * i + 1] = (int)(v >> 16);
      sum += loc[2 * i] + loc[2 * i + 1];
    }
    int inc = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, inc, off);
      if (lane >= off) inc += v;
    }
    int run = inc - sum, bin = 0, newk = 0;
    const bool mine = run < krem && krem <= inc;
    if (mine) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        if (run < krem && krem <= run + loc[i]) {
          bin = lane * 8 + i;
          newk = krem - run;
        }
        run += loc[i];
      }
    }
    const int src = __ffs(__ballot_sync(0xffffffffu, mine)) - 1;
    bin = __shfl_sync(0xffffffffu, bin, src);
    krem = __shfl_sync(0xffffffffu, newk, src);
    prefix = (prefix << 8) | (uint32_t)bin;
  }
  return prefix;
}

// ------------------------------------------------------------------------------------------------------
// K5: bank-conflict-free scan, period P = 8 (m <= 8) or 16 (m = 9..16) codebooks.
//
// A 4-byte LUT lookup per byte of code makes the scan shared-memory-gather bound (one LDS.64 warp-instruction per
// 2 clk per SM, measured: tools/ubench/lds_ffma2.cu), and 32 lanes looking up random entries of the same
// 256-entry row collide ~3.6x (measured on the round-1 v1 kernel, profiles/r1_v1).  Here every lane of a half-warp
// is at a DIFFERENT bank-pair at any instant:
//   P = 8 : tile[tt][c][bp = g*8 + k][e]  (float; 4 queries per 32 KB tile: q = tt*4 + g*2 + e); lane = hw*16 + g*8 + j
//           handles code stream hw*8 + j of its warp's chunk and the query pair g of each of the 4 tiles (16 queries
//           per block, 16 streams per warp);
//   P = 16: tile[tt][c][bp = k][e]  (2 queries per 32 KB tile, 8 queries per block); lane = stream (32 per warp).
// At step s a lane is at codebook k = (s - j - 1) mod P (j = lane mod P), so the 16 lanes of a half-warp hit 16
// distinct bank-pairs with one LDS.64 each -> no conflicts, 2 queries per load.  The sum for one code must still be
// ((0 + t_0) + t_1) + ... in ascending k (pairwise_byte.cpp:70-73), so a lane's code simply starts j+1 steps
// "late": the index stores each lane's stream pre-skewed (skew_fields_kernel).  Accumulate / restart / capture are
// packed FFMA2 (fma.rn.f32x2, exact per element):  acc = acc*keep_s + v  (keep_s = 0 at the step where the lane's
// next code starts), done += acc*cap_s (cap_s = 1 at the step where its code completes).
// The index holds, per step, the 16-bit field (c << 7) | (k << 3): the byte offset of the LUT entry inside a tile
// for THIS stream at THIS step (k is known when the index is built), and the tile is 32 KB-aligned in shared
// memory, so a step's address is ONE instruction.  One period = P steps = one completed code per lane; a lane
// reads 2P bytes of index per period.  The 128 KB LUT tile is staged with bulk async copies (cp.async.bulk +
// mbarrier).
// ------------------------------------------------------------------------------------------------------
template <int P>
struct ScanX {
  static constexpr int G = 16 / P;                 // query-pair groups per half-warp (2 or 1)
  static constexpr int NS = 32 / G;                // code streams per warp (16 or 32)
  static constexpr int L = kChunkCodes / NS;       // codes per stream per chunk (64 or 32)
  static constexpr int PERIODS = L + 1;            // + one period for the skew tail
  static constexpr int HALVES = P / 8;             // uint4 words per lane per period
  static constexpr int QB = 8 * G;                 // queries per block (16 or 8)
  static constexpr int ADDS = kScanWarps * NS;    // most keys one query can gain per block-period
};

// F[chunk][t][half][p] (uint4 = 8 fields): stream p = codes chunk*1024 + NS*u + p (u = 0..L-1), delayed by (p mod P)+1
template <int P>
__global__ void skew_fields_kernel(const uint8_t* __restrict__ codes, uint4* __restrict__ F, int64_t n, int m,
                                   int64_t nchunks) {
  using X = ScanX<P>;
  const int64_t total = nchunks * X::PERIODS * X::HALVES * X::NS;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int p = (int)(i % X::NS);
    const int hf = (int)((i / X::NS) % X::HALVES);
    const int t = (int)((i / (X::NS * X::HALVES)) % X::PERIODS);
    const int64_t chunk = i / ((int64_t)X::NS * X::HALVES * X::PERIODS);
    const int delay = (p & (P - 1)) + 1;
    uint32_t f[8];
#pragma unroll
    for (int b = 0; b < 8; b++) {
      const int sb = P * t + hf * 8 + b - delay;       // byte index in the stream's undelayed sequence
      const int k = sb & (P - 1);
      uint32_t field = (uint32_t)k << 3;
      if (sb >= 0 && sb < P * X::L) {
        const int64_t id = chunk * kChunkCodes + (int64_t)X::NS * (sb / P) + p;
        if (id < n && k < m) field |= (uint32_t)codes[id * m + k] << 7;
      }
      f[b] = field;
    }
    F[i] = make_uint4(f[0] | (f[1] << 16), f[2] | (f[3] << 16), f[4] | (f[5] << 16), f[6] | (f[7] << 16));
  }
}

// Fq[chunk][t][half][p] (uint4 = 8 fields), the pre-filter scan's copy: stream p = codes chunk*1024 + NS*t + p
// (t = 0..L-1), one code per period, its codebooks ROTATED by the lane: step S holds codebook (S + p) mod P, so the lanes
// of a half-warp sit on P different bank-pairs at every step while all of them start and finish a code together.
template <int P>
__global__ void rot_fields_kernel(const uint8_t* __restrict__ codes, uint4* __restrict__ F, int64_t n, int m,
                                  int64_t nchunks) {
  using X = ScanX<P>;
  const int64_t total = nchunks * X::L * X::HALVES * X::NS;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int p = (int)(i % X::NS);
    const int hf = (int)((i / X::NS) % X::HALVES);
    const int t = (int)((i / (X::NS * X::HALVES)) % X::L);
    const int64_t chunk = i / ((int64_t)X::NS * X::HALVES * X::L);
    const int64_t id = chunk * kChunkCodes + (int64_t)X::NS * t + p;
    uint32_t f[8];
#pragma unroll
    for (int b = 0; b < 8; b++) {
      const int k = (hf * 8 + b + p) & (P - 1);
      uint32_t field = (uint32_t)k << 3;
      if (id < n && k < m) field |= (uint32_t)codes[id * m + k] << 7;
      f[b] = field;
    }
    F[i] = make_uint4(f[0] | (f[1] << 16), f[2] | (f[3] << 16), f[4] | (f[5] << 16), f[6] | (f[7] << 16));
  }
}

struct ScanXParams {
  const uint4* F;        // skewed offset fields
  const float* norms;    // [n] or nullptr
  const float* lut;      // tiled [qtiles][32768]
  uint64_t* cand;        // [slices][qtiles*QB][cap]
  uint64_t* part;        // [slices][nq][k]
  const uint64_t* lb;    // [nq] or nullptr
  int64_t n, nchunks, chunks_per_slice;
  int nq, k, cap, soft;
  int piggy;             // a service() also compacts every buffer already past this many keys
  int spec_soft;         // soft limit while a speculative threshold (r < k) is in force
  float tau0;            // initial threshold (+inf; a finite value is a measurement aid, RAYUELA_B200_SCAN_TAU0)
  int spec;              // speculative thresholds on (verified at the end; a failed block is redone in pass 1)
  int pass;              // 0: main launch; 1: redo launch -- only blocks whose redo flag is set run, without speculation
  int* redo;             // [slices][qtiles] flags
  // QPF (quantised pre-filter scan, see scanx_kernel): quantised tiles + their parameters, the raw codes for the exact
  // re-evaluation of the survivors, and the per-query lists of survivors awaiting it
  const float* lutq;     // tiled [qtiles][16384]: two queries per word (lut_quant_kernel)
  const uint4* Fq;       // rotated offset fields (rot_fields_kernel)
  const float4* tilep;   // [qtiles] {s, 1/s, norm rounding constant, -}
  const double* qoff;    // [qtiles*QB] sum_k min_c LUT + min norm: the distance that quantises to 0
  const int* qmu;        // [qtiles*QB] threshold margin in units of s
  const uint8_t* codes;  // [n][m] raw codes
  int m;
  unsigned long long* qstats;   // [2] or null (RAYUELA_B200_SCAN_STATS): survivors of the pre-filter, of those accepted
};

// QPF = true: the same scan with a quantised INTEGER pre-filter in front of the exact arithmetic -- results bit-identical.
// The scan is bound by shared-memory bandwidth (4 bytes looked up per byte of code and query), so the tile is stored a
// second time as small integers in units of a per-tile scale s:  v[q][k][c] = rint((LUT[q][k][c] - min_c LUT[q][k][.]) / s)
// (+1 for k = 0), TWO queries per fp32 word as the exact integer v_hi * 4096 + v_lo, i.e. four queries per LDS.64 and per
// FFMA2 (integers below 2^24 add exactly in fp32, so the packed accumulate / restart / capture of the fp32 loop carries
// over unchanged at half the loads and half the math per query).  s is chosen so that the m terms plus the quantised
// norm stay below 2^11 per query: no carry between the two digits.  For every code
//     A = sum_k v + rint((norm - min norm)/s)   satisfies   A <= 1 + (E - off_q)/s + 0.5 m + 1.5 + fp32 slack
// where E is the exact fp32 distance of the reference chain and off_q = sum_k min_c LUT + min norm.  A code is a
// SURVIVOR of query q when A <= T_q = floor((tau_q - off_q)/s) + mu_q (mu_q = that margin, lut_quant_kernel), which every
// code with E <= tau_q is.  The test is one exact subtraction per pair, (T + 2048) - A per base-4096 digit: a digit keeps
// its 2048 bit iff A <= T and never borrows from its neighbour.  Survivors are only noted ((query, id) pushed to the
// warp's own 64-entry ring); every 32 of them the warp re-evaluates them itself, one lane per survivor, with the exact
// fp32 chain from the fp32 tile in L2 (ascending k from 0, + norm last -- the arithmetic of the fp32 hot loop) and hands
// those with E <= tau_q to the unchanged candidate buffers / compaction / speculation.  The window is ~4e-3 of the
// tile's distance range.
template <int P, bool NORMS, bool SPEC, bool QPF = false>
__global__ void __launch_bounds__(kScanWarps * 32, 1) scanx_kernel(ScanXParams p) {
  using X = ScanX<P>;
  constexpr int NT = kScanWarps * 32;
  constexpr int QB = X::QB;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int cnt_s[16];
  __shared__ float tau_s[16];
  __shared__ uint64_t taukey_s[16];  // the threshold as a (dist, id) key: every key seen so far that is <= it is in the buffer
  __shared__ uint64_t lb_s[16];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t hist_s[3 * 128];   // three rotating histograms of 256 16-bit bins (block_radix_select)
  __shared__ int sel_s[4];
  __shared__ int flag_s;     // some query's buffer needs compacting (set by the appending thread)
  __shared__ int nfin_s;     // warps that have finished their chunks
  __shared__ int warm_s;     // some query still has tau = +inf: react to the flag within the same period
  __shared__ int finq_s[16]; // final phase: query already written by its warp
  __shared__ int seen_s;     // codes of this block's slice scanned so far (all warps)
  __shared__ int softq_s[16];// per-query soft limit (lower while a speculative threshold waits for confirmation)
  __shared__ int fail_s;     // a speculative threshold turned out too tight for some query: redo the block
  __shared__ int thr_s[16];  // QPF: integer thresholds T_q (0: nothing passes, 2047: everything does)

  // dynamic shared memory: [sort buffer 64 KB][pad][LUT tile 128 KB, 32 KB-aligned] -- the alignment makes the
  // tile base and the 15-bit offset fields disjoint bit ranges, so a step's address is ONE instruction
  // ((w & 0xFFFF) | base, or (w >> 16) + base)
  uint64_t* sortbuf = reinterpret_cast<uint64_t*>(smem_raw);
  const uint32_t lut_addr = (smem_u32(smem_raw) + kScanSortKeys * 8 + 32767u) & ~32767u;

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int j = lane & (P - 1);
  const int g = (P == 8) ? ((lane >> 3) & 1) : 0;
  const int pidx = (P == 8) ? ((lane >> 4) * 8 + j) : lane;
  const int q0 = blockIdx.x * QB;
  const int slice = blockIdx.y;
  uint64_t* cand = p.cand + ((size_t)slice * gridDim.x * QB + (size_t)blockIdx.x * QB) * p.cap;
  const float inf = __int_as_float(0x7f800000);
  // QPF: T_q from the current exact threshold (double arithmetic: once per query per service())
  auto thr_of = [=](int q) -> int {
    const float tau = tau_s[q];
    if (tau == inf) return 2047;
    if (!(tau > -inf)) return 0;
    const float4 tp = __ldg(p.tilep + blockIdx.x);
    const double x = floor(((double)tau - __ldg(p.qoff + (size_t)blockIdx.x * QB + q)) / (double)tp.x) +
                     (double)__ldg(p.qmu + (size_t)blockIdx.x * QB + q);
    if (!(x >= 1.0)) return 0;                     // A >= 1 for every code (the +1 of codebook 0)
    return (int)fmin(x, 2046.0);
  };

  const uint32_t mbar_addr = smem_u32(&mbar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_addr));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 16) {
    cnt_s[tid] = 0;
    // padded dummy queries of the last tile (all-zero LUT) must never append: their threshold is -inf for good
    tau_s[tid] = (q0 + tid < p.nq) ? p.tau0 : -__int_as_float(0x7f800000);
    taukey_s[tid] = make_key(p.tau0, 0xFFFFFFFFu);
    lb_s[tid] = (p.lb && tid < QB) ? p.lb[min(q0 + tid, p.nq - 1)] : 0ull;
    if (QPF) {
      thr_s[tid] = tid < QB ? thr_of(tid) : 0;
    }
  }
  if (SPEC && p.pass == 1 && p.redo[blockIdx.y * gridDim.x + blockIdx.x] == 0) return;   // redo launch: nothing to redo
  const bool spec = SPEC && p.pass == 0;           // SPEC = false instantiations carry none of the speculation code
  if (SPEC && tid < 16) softq_s[tid] = p.soft;
  if (tid == 0) {
    flag_s = 0;
    nfin_s = 0;
    warm_s = 1;
    seen_s = 0;
    fail_s = 0;
  }
  block_sync();
  if (tid == 0) {
    constexpr int kTileBytes = QPF ? kLutTileBytes / 2 : kLutTileBytes;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_addr), "r"(kTileBytes)
                 : "memory");
    const char* src = QPF ? reinterpret_cast<const char*>(p.lutq) + (size_t)blockIdx.x * kTileBytes
                          : reinterpret_cast<const char*>(p.lut) + (size_t)blockIdx.x * kTileBytes;
#pragma unroll
    for (int i = 0; i < kTileBytes / 32768; i++)
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
              lut_addr + i * 32768),
          "l"(src + i * 32768), "r"(32768), "r"(mbar_addr)
          : "memory");
  }
  {
    uint32_t ok = 0;
    while (!ok) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok)
          : "r"(mbar_addr)
          : "memory");
    }
  }

  // per-lane constants: restart (keep = 0) at step (j+1) mod P, capture at step j
  const uint32_t base = lut_addr + (g << 6);
  float keep[P], capf[P];
#pragma unroll
  for (int s = 0; s < P; s++) {
    keep[s] = (s == ((j + 1) & (P - 1))) ? 0.f : 1.f;
    capf[s] = (s == j) ? 1.f : 0.f;
  }
  float tau[8];
#pragma unroll
  for (int i = 0; i < 8; i++)
    tau[i] = (q0 + (i >> 1) * (2 * X::G) + g * 2 + (i & 1) < p.nq) ? p.tau0 : -__int_as_float(0x7f800000);
  uint64_t acc[4] = {0, 0, 0, 0}, done[4] = {0, 0, 0, 0};
  // QPF: the lane's 8 queries are i = 0..7 <-> block query (i >> 2) * 4G + 4g + (i & 3): LDS.64 number i >> 2 of a step,
  // fp32 word (i >> 1) & 1 of it, base-4096 digit i & 1 of that word.  thr2: per word (T_hi + 2048) * 4096 + (T_lo + 2048).
  uint64_t thr2[2] = {0, 0};
  auto lane_query = [=](int i) -> int { return (i >> 2) * (4 * X::G) + g * 4 + (i & 3); };
  auto thr_word = [=](int i) -> float {      // word holding queries i (low digit) and i + 1
    return (float)((thr_s[lane_query(i + 1)] + 2048) * 4096 + thr_s[lane_query(i)] + 2048);
  };
  float q_invs = 0.f, q_c0 = 0.f, q_ncap = 0.f;
  if (QPF) {
    const float4 tp = __ldg(p.tilep + blockIdx.x);
    q_invs = tp.y;
    q_c0 = tp.z;
    q_ncap = tp.w;
    thr2[0] = pack2(thr_word(0), thr_word(2));
    thr2[1] = pack2(thr_word(4), thr_word(6));
  }
  int warm = 1;
  const int64_t c0 = (int64_t)slice * p.chunks_per_slice;
  const int64_t c1 = min(p.nchunks, c0 + p.chunks_per_slice);
  const float nslice = (float)(min(p.n, c1 * kChunkCodes) - c0 * kChunkCodes);     // codes in this block's slice

  // Speculative threshold.  With a fraction f = seen/nslice of the slice scanned, the final k-th distance is close to
  // the (k*f)-th smallest seen.  A compaction therefore keeps only the r = x + z*sqrt(x) + z*z/2 smallest keys
  // (x = k*f, z = 5.5: the count of keys below the r-th in the whole slice is >= k except with probability ~1e-7 on
  // exchangeable data) and makes the r-th the threshold, which cuts the appends and compactions of large-k searches
  // several-fold.  It is VERIFIED, not trusted: thresholds only ever decrease and every key seen so far that is <= the
  // current threshold KEY (taukey_s) is in the buffer, so the result is exact iff the k-th smallest key of the final
  // buffer is <= the final threshold key; otherwise the block raises its redo flag and is rerun without speculation
  // (r = k: plain "keep the k best") by the second launch.
  bool final_phase = false;                                    // no tightening once the scan is over
  auto spec_rank = [=, &final_phase]() -> int {
    if (!SPEC || !spec || final_phase) return p.k;
    const float x = (float)p.k * fminf(1.0f, (float)seen_s / nslice);
    const int r = (int)(x + 5.5f * sqrtf(x) + 15.125f) + 1;
    return (r * 4 < p.k * 3) ? r : p.k;
  };

  // Small buffers (<= kWarpKeys keys) are handled by ONE warp each, all queries of the block at once: the warp
  // sorts its query's keys in its own 4 KB slice of the sort area (warp-level bitonic network, no block barriers)
  // and leaves them there sorted; returns the number of keys kept (min(c, k)).
  constexpr int kWarpKeys = kScanSortKeys / kScanWarps;   // 512
  uint64_t* wslice = sortbuf + w * kWarpKeys;
  auto warp_bitonic = [=](uint64_t* buf, int np2) {            // np2 >= 64, a power of two; one warp
    for (int size = 2; size <= np2; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int t = lane; t < (np2 >> 1); t += 32) {
          const int lo = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
          const int hi = lo | stride;
          const bool up = (lo & size) == 0;
          const uint64_t a = buf[lo], b = buf[hi];
          if ((a > b) == up) {
            buf[lo] = b;
            buf[hi] = a;
          }
        }
        __syncwarp();
      }
    }
  };
  // Compaction keeps the r = spec_rank() <= k smallest keys and makes the r-th the new threshold (tau_s as a distance for
  // the filter, taukey_s as a full key for the invariant: every key seen so far that is <= taukey_s is in the buffer --
  // later arrivals with key <= taukey_s pass the float filter, compactions keep everything <= the new, smaller key).
  // Without speculation r = k and this is the plain "keep the k best".  ONE selection per compaction.
  auto warp_sort_keep = [=](int q) -> int {
    const int c = cnt_s[q];
    uint64_t* cq = cand + (size_t)q * p.cap;
    const int np2 = max(64, pow2ceil(c));
    for (int t = lane; t < np2; t += 32) wslice[t] = t < c ? cq[t] : ~0ull;
    __syncwarp();
    warp_bitonic(wslice, np2);
    const int r = spec_rank();
    const int keep = c >= p.k ? r : c;
    if (c > keep)
      for (int t = lane; t < keep; t += 32) cq[t] = wslice[t];
    if (lane == 0) {
      cnt_s[q] = keep;
      if (c >= p.k) {
        tau_s[q] = fminf(tau_s[q], ordered_to_f32((uint32_t)(wslice[r - 1] >> 32)));
        taukey_s[q] = min(taukey_s[q], wslice[r - 1]);
        if (SPEC) softq_s[q] = r < p.k ? min(p.soft, p.spec_soft) : p.soft;
      }
    }
    return keep;
  };
  auto needs_compaction = [=](int q) -> bool {
    const int c = cnt_s[q];
    return c > (SPEC ? min(p.piggy, softq_s[q]) : p.piggy) || (c >= p.k && tau_s[q] == __int_as_float(0x7f800000));
  };
  // exactness check of a finished query (see spec_rank): kth = k-th smallest key of the final buffer, c its size.
  // kth <= taukey_s  =>  every key <= kth ever seen is in the buffer  =>  the buffer's k smallest are the true top-k.
  auto verify = [=](int q, int c, uint64_t kth) {
    if (!SPEC) return;
    if (c >= p.k ? kth > taukey_s[q] : tau_s[q] < __int_as_float(0x7f800000)) fail_s = 1;
  };

  auto compact = [=](int q) {
    const int c = cnt_s[q];
    uint64_t* cq = cand + (size_t)q * p.cap;
    if (c <= p.k) {
      if (c == p.k) {
        uint64_t mx = 0;
        for (int t = tid; t < c; t += NT) mx = max(mx, cq[t]);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, off));
        if (lane == 0) sortbuf[w] = mx;
        block_sync();
        if (tid == 0) {
          for (int i = 1; i < kScanWarps; i++) mx = max(mx, sortbuf[i]);
          tau_s[q] = fminf(tau_s[q], ordered_to_f32((uint32_t)(mx >> 32)));
          taukey_s[q] = min(taukey_s[q], mx);
        }
        block_sync();
      }
      return;
    }
    const int r = spec_rank();                                 // block-uniform (seen_s is stable inside service())
    if (c <= 512) {
      const int np2 = pow2ceil(c);
      for (int t = tid; t < np2; t += NT) sortbuf[t] = t < c ? cq[t] : ~0ull;
      block_sync();
      block_bitonic_sort(sortbuf, np2);
      for (int t = tid; t < r; t += NT) cq[t] = sortbuf[t];
      if (tid == 0) {
        cnt_s[q] = r;
        tau_s[q] = fminf(tau_s[q], ordered_to_f32((uint32_t)(sortbuf[r - 1] >> 32)));
        taukey_s[q] = min(taukey_s[q], sortbuf[r - 1]);
        if (SPEC) softq_s[q] = r < p.k ? min(p.soft, p.spec_soft) : p.soft;
      }
      block_sync();
      return;
    }
    for (int t = tid; t < c; t += NT) sortbuf[t] = cq[t];
    block_sync();
    const uint64_t pivot = block_radix_select(sortbuf, c, r, hist_s);
    block_sync();
    if (tid == 0) sel_s[2] = 0;
    block_sync();
    for (int t = tid; t < c; t += NT) {
      const uint64_t key = sortbuf[t];
      if (key <= pivot) cq[atomicAdd(&sel_s[2], 1)] = key;     // exactly r keys (keys are distinct)
    }
    if (tid == 0) {
      cnt_s[q] = r;
      tau_s[q] = fminf(tau_s[q], ordered_to_f32((uint32_t)(pivot >> 32)));
      taukey_s[q] = min(taukey_s[q], pivot);
      if (SPEC) softq_s[q] = r < p.k ? min(p.soft, p.spec_soft) : p.soft;
    }
    block_sync();
  };
  auto finalize = [=](int q) {
    compact(q);
    const int c = cnt_s[q];
    const int np2 = pow2ceil(c);
    uint64_t* cq = cand + (size_t)q * p.cap;
    for (int t = tid; t < np2; t += NT) sortbuf[t] = t < c ? cq[t] : ~0ull;
    block_sync();
    block_bitonic_sort(sortbuf, np2);
    if (tid == 0) verify(q, c, sortbuf[min(c, p.k) - (c > 0)]);
  };

  // Compaction is event-driven: the thread whose append pushes a buffer past the soft limit (or completes the first
  // k candidates while tau is still +inf) raises flag_s; every warp reads the flag at the START of each period
  // (so the load's latency hides behind the period's lookups), acts on it at the END, and then joins service().
  // Between the raise and the last warp's reaction a query gains at most 2*ADDS keys (two periods of every warp),
  // which the capacity soft + 3*ADDS covers.  Finished warps wait in service() until all warps are done, so the
  // block-wide barriers inside always see all 16 warps.
  // service() is deliberately NOT inlined: it has two call sites (working warps inside the period loop, finished
  // warps in their wait loop) and must exist once, so that all 16 warps of the block execute the very same barrier
  // instructions; it hands the refreshed thresholds back through tau_x / warm_x (address-taken locals) so that the
  // hot loop's own copies stay in registers; everything else is captured BY VALUE (also by the helpers it calls) so
  // that taking the closure's address does not push the kernel's locals into local memory.
  // QPF: exact distance of code id for block query q: ((0 + t_0) + t_1) + ... in ascending k from the fp32 tile in L2,
  // + dbnorms[i] last (pairwise_byte.cpp:70-74) -- the fp32 hot loop's arithmetic.  All code bytes first, then all table
  // entries: two load levels per survivor.
  auto exact_dist = [=](int q, uint32_t id) -> float {
    const float* lt = p.lut + (size_t)blockIdx.x * (kLutTileBytes / 4);
    const float* lq = (P == 8) ? lt + (q >> 2) * 8192 + (((q & 3) >> 1) << 4) + (q & 1) : lt + (q >> 1) * 8192 + (q & 1);
    const uint8_t* cb = p.codes + (size_t)id * p.m;
    float nrm = 0.f;
    if (NORMS) nrm = __ldg(p.norms + id);
    uint32_t cw[P / 4];
    if (P == 8 && p.m == 8) {
      const uint2 c8 = __ldg(reinterpret_cast<const uint2*>(cb));
      cw[0] = c8.x;
      cw[1] = c8.y;
    } else if (P == 16 && p.m == 16) {
      const uint4 c16 = __ldg(reinterpret_cast<const uint4*>(cb));
      cw[0] = c16.x; cw[1] = c16.y; cw[2] = c16.z; cw[3] = c16.w;
    } else {
#pragma unroll
      for (int w4 = 0; w4 < P / 4; w4++) {
        uint32_t x = 0;
#pragma unroll
        for (int b = 0; b < 4; b++)
          if (w4 * 4 + b < p.m) x |= (uint32_t)__ldg(cb + w4 * 4 + b) << (8 * b);
        cw[w4] = x;
      }
    }
    float tv[P];
#pragma unroll
    for (int k = 0; k < P; k++) tv[k] = k < p.m ? __ldg(lq + (int)((cw[k >> 2] >> (8 * (k & 3))) & 255u) * 32 + (k << 1)) : 0.f;
    float d = 0.f;
#pragma unroll
    for (int k = 0; k < P; k++)
      if (k < p.m) d = __fadd_rn(d, tv[k]);
    if (NORMS) d = __fadd_rn(d, nrm);
    return d;
  };

  float tau_x[8];
  uint64_t thr_x[2];
  int warm_x = 1;
  auto service = [=, &tau_x, &thr_x, &warm_x](int progress) __attribute__((noinline)) -> bool {   // progress: codes this warp has completed
    block_sync();                                   // nobody is appending past this point
    const int nf = nfin_s;
    const int fl = *(volatile int*)&flag_s;
    if (SPEC && fl && spec) {                          // block-wide progress for spec_rank()
      if (lane == 0) atomicAdd(&seen_s, progress);
      block_sync();
    }
    if (fl) {
      if (w < QB && cnt_s[w] <= kWarpKeys && needs_compaction(w)) warp_sort_keep(w);   // warp w <-> query w
      block_sync();
      for (int q = 0; q < QB; q++)
        if (needs_compaction(q)) compact(q);                                           // the large ones, block-wide
      if (tid == 0) {
        flag_s = 0;
        seen_s = 0;
        int warm = 0;
        for (int q = 0; q < QB; q++) warm |= tau_s[q] == __int_as_float(0x7f800000);
        warm_s = warm;
      }
      if (QPF && tid < QB) thr_s[tid] = thr_of(tid);
    }
    block_sync();
    warm_x = warm_s;
    if (QPF) {
      thr_x[0] = pack2(thr_word(0), thr_word(2));
      thr_x[1] = pack2(thr_word(4), thr_word(6));
    } else {
#pragma unroll
      for (int tt = 0; tt < 4; tt++) {
        tau_x[2 * tt] = tau_s[tt * 2 * X::G + g * 2];
        tau_x[2 * tt + 1] = tau_s[tt * 2 * X::G + g * 2 + 1];
      }
    }
    return nf == kScanWarps;
  };

  const uint32_t n32 = (uint32_t)p.n;
  const uint32_t flag_addr = smem_u32(&flag_s);

  if constexpr (QPF) {
    // Pre-filter loop.  Integer sums are exact in any order, so a lane's code needs no fixed codebook order here: the index
    // holds a second, ROTATED copy of the offset fields (rot_fields_kernel) in which lane j = lane mod P visits codebook
    // (S + j) mod P at step S -- the lanes of a half-warp are still at P different bank-pairs at every instant, but every
    // lane starts and completes its code together with the period: no restart multiply, no capture, no skew tail; a step is
    // 2 LDS.64 + 2 packed adds for 8 queries.  The fields come from L2 (latency ~ a period and a half of this loop), so they
    // are fetched TWO periods ahead into two alternating register sets, reloaded as soon as the period's shared-memory
    // addresses have been formed from them; same for the norms.
    constexpr int LQ = X::L;                                  // periods per chunk = codes per stream
    int wdone = 0;
    const uint4* fq0 = nullptr;
    const float* np0 = nullptr;
    uint32_t id0 = 0;
    // Survivors go to a 64-entry ring of (query, id) pairs PER WARP (in the half of the tile area the 2-byte tables leave
    // free); whenever it holds 32 the warp itself re-evaluates them exactly, one lane per survivor, and appends those with
    // E <= tau to the candidate buffers exactly as the fp32 loop does -- no block barrier, the other 15 warps keep scanning
    // under the two dependent L2 round trips.  Ring state is warp-uniform (registers).
    uint64_t* ring = reinterpret_cast<uint64_t*>(smem_raw + (lut_addr - smem_u32(smem_raw)) + 65536) + w * 64;
    uint32_t rhead = 0, rfill = 0;
    auto flush_batch = [&]() {
      __syncwarp();
      const uint32_t cnt = min(rfill, 32u);
      bool ok = false;
      if (lane < cnt) {
        const uint64_t e = ring[(rhead + lane) & 63u];
        const int q = (int)(e >> 32);
        const uint32_t id = (uint32_t)e;
        const float d = exact_dist(q, id);
        const float tq = tau_s[q];
        if (d <= tq) {
          const uint64_t key = make_key(d, id);
          if (!p.lb || key > lb_s[q]) {
            ok = true;
            const int pos = atomicAdd(&cnt_s[q], 1);
            cand[(size_t)q * p.cap + pos] = key;
            if (pos >= (SPEC ? softq_s[q] : p.soft) || (tq == inf && pos + 1 >= p.k)) atomicExch(&flag_s, 1);
          }
        }
      }
      if (p.qstats) {
        const uint32_t acc_n = __popc(__ballot_sync(0xffffffffu, ok));
        if (lane == 0) {
          atomicAdd(p.qstats, (unsigned long long)cnt);
          atomicAdd(p.qstats + 1, (unsigned long long)acc_n);
        }
      }
      rhead = (rhead + cnt) & 63u;
      rfill -= cnt;
      __syncwarp();
    };
    auto flush_all = [&]() {
      while (rfill) flush_batch();
    };
    auto period_q = [&](uint4 (&Wc)[X::HALVES], float& nc, const int t) {
      const int raised = lds_volatile(flag_addr);
      uint32_t ad[P];
#pragma unroll
      for (int hf = 0; hf < X::HALVES; hf++) {
        const uint32_t wr[4] = {Wc[hf].x, Wc[hf].y, Wc[hf].z, Wc[hf].w};
#pragma unroll
        for (int b8 = 0; b8 < 8; b8++)
          ad[hf * 8 + b8] = (b8 & 1) ? (wr[b8 >> 1] >> 16) + base : ((wr[b8 >> 1] & 0xFFFFu) | base);
      }
      const uint32_t id = id0 + (uint32_t)(t * X::NS);         // the code this period evaluates
      const float nrm = nc;
#ifndef QPF_VARIANT
#define QPF_VARIANT 3
#endif
      auto prefetch = [&]() {
        if (t + 2 < LQ) {
#pragma unroll
          for (int hf = 0; hf < X::HALVES; hf++) Wc[hf] = __ldg(fq0 + (t + 2) * (X::HALVES * X::NS) + hf * X::NS);
          if (NORMS) {
            if (QPF_VARIANT & 2) {
              nc = __ldg(p.norms + min(id + 2 * X::NS, n32 - 1));
            } else {
              nc = 0.f;
              if (id + 2 * X::NS < n32) nc = __ldg(np0 + (t + 2) * X::NS);
            }
          }
        }
      };
      if (!(QPF_VARIANT & 1)) prefetch();
      uint64_t a0 = lds64<0>(ad[0]), a1 = lds64<32768>(ad[0]);
#pragma unroll
      for (int S = 1; S < P; S++) {
        a0 = fadd2(a0, lds64<0>(ad[S]));
        a1 = fadd2(a1, lds64<32768>(ad[S]));
      }
      if (QPF_VARIANT & 1) prefetch();
      {
        // (T + 2048) - A per digit, exact in fp32 (integers < 2^24); + 2^24 aligns the integer with the mantissa (the
        // one rounding, to even, can only turn a digit of 2047 into 2048: a spurious survivor, never a lost one)
        const uint64_t neg1 = pack2(-1.0f, -1.0f);
        uint64_t u0 = ffma2(a0, neg1, thr2[0]), u1 = ffma2(a1, neg1, thr2[1]);
        if (NORMS) {
          const float nqf = fminf(__fadd_rn(fmaf(nrm, q_invs, q_c0), -12582912.0f), q_ncap);   // min(rint((norm - norm0)/s), cap)
          const uint64_t n2 = pack2(nqf, nqf), m4097 = pack2(-4097.0f, -4097.0f);
          u0 = ffma2(n2, m4097, u0);
          u1 = ffma2(n2, m4097, u1);
        }
        const uint64_t two24 = pack2(16777216.0f, 16777216.0f);
        u0 = fadd2(u0, two24);
        u1 = fadd2(u1, two24);
        const uint32_t uw[4] = {(uint32_t)u0, (uint32_t)(u0 >> 32), (uint32_t)u1, (uint32_t)(u1 >> 32)};
        const bool has = id < n32 && ((uw[0] | uw[1] | uw[2] | uw[3]) & 0x00400400u) != 0u;
        if (__any_sync(0xffffffffu, has)) {                       // warp-uniform from here on
          uint32_t bits = 0;                                      // bit i: query i of this lane survived
          if (has) {
#pragma unroll
            for (int i = 0; i < 8; i++) bits |= ((uw[i >> 1] >> ((i & 1) ? 22 : 10)) & 1u) << i;
          }
          uint32_t present = __reduce_or_sync(0xffffffffu, bits);
          while (present) {                                        // one query slot at a time: <= 32 entries per round
            const int b = __ffs(present) - 1;
            present &= present - 1;
            const bool mine = (bits >> b) & 1u;
            const uint32_t bal = __ballot_sync(0xffffffffu, mine);
            const uint32_t nb = __popc(bal);
            if (rfill + nb > 64u) flush_batch();
            if (mine)
              ring[(rhead + rfill + __popc(bal & ((1u << lane) - 1u))) & 63u] = ((uint64_t)(uint32_t)lane_query(b) << 32) | id;
            rfill += nb;
          }
          if (rfill >= 32u) flush_batch();
        }
      }
      // while some tau is still +inf every code of every warp is a candidate: do not wait a period to react
      if (raised || (warm && lds_volatile(flag_addr))) {
        flush_all();                                              // every survivor noted so far is in the buffers
        service(wdone + (t + 1) * X::NS);
        // The refreshed thresholds are rebuilt HERE from shared memory (conversions = real consumers on this cold path),
        // not handed back through service()'s address-taken locals: those come back as local-memory loads whose
        // scoreboard ptxas shares with the field prefetches, and the next period's threshold test then waits on it --
        // every other period stalled until its own prefetch had landed (ncu: 25 % of the kernel on that one FFMA2).
        thr2[0] = pack2(thr_word(0), thr_word(2));
        thr2[1] = pack2(thr_word(4), thr_word(6));
        warm = lds_volatile(smem_u32(&warm_s));
      }
    };
    for (int64_t chunk = c0 + w; chunk < c1; chunk += kScanWarps, wdone += kChunkCodes) {
      fq0 = p.Fq + chunk * (LQ * X::HALVES * X::NS) + pidx;
      np0 = p.norms + chunk * kChunkCodes + pidx;
      id0 = (uint32_t)(chunk * kChunkCodes) + pidx;
      uint4 WA[X::HALVES], WB[X::HALVES];
#pragma unroll
      for (int hf = 0; hf < X::HALVES; hf++) {
        WA[hf] = __ldg(fq0 + hf * X::NS);
        WB[hf] = __ldg(fq0 + X::HALVES * X::NS + hf * X::NS);
      }
      float nA = 0.f, nB = 0.f;
      if (NORMS && id0 < n32) nA = __ldg(np0);
      if (NORMS && id0 + X::NS < n32) nB = __ldg(np0 + X::NS);
      for (int t = 0; t < LQ; t += 2) {
        period_q(WA, nA, t);
        period_q(WB, nB, t + 1);
      }
    }
    flush_all();
    __syncwarp();
    if (lane == 0) atomicAdd(&nfin_s, 1);
    __syncwarp();
    while (!service(wdone)) {
    }
  } else {
  int wdone = 0;                                         // codes of finished chunks of this warp
  for (int64_t chunk = c0 + w; chunk < c1; chunk += kScanWarps, wdone += kChunkCodes) {
    const uint4* fp = p.F + chunk * (X::PERIODS * X::HALVES * X::NS) + pidx;
    const float* np = p.norms + chunk * kChunkCodes + pidx;            // norm of the code completed in period t+1
    uint32_t id = (uint32_t)(chunk * kChunkCodes) + pidx - X::NS;      // code completed in period t (t >= 1)
    uint4 W[X::HALVES], Wn[X::HALVES];
#pragma unroll
    for (int hf = 0; hf < X::HALVES; hf++) W[hf] = __ldg(fp + hf * X::NS);
    fp += X::HALVES * X::NS;
    float nrm0 = 0.f;
    for (int t = 0; t < X::PERIODS; t++) {
      const int raised = lds_volatile(flag_addr);
#pragma unroll
      for (int hf = 0; hf < X::HALVES; hf++) {
        Wn[hf] = make_uint4(0, 0, 0, 0);
        if (t + 1 < X::PERIODS) Wn[hf] = __ldg(fp + hf * X::NS);
      }
      float nrm1 = 0.f;
      if (NORMS && t + 1 < X::PERIODS && id + X::NS < n32) nrm1 = __ldg(np);
#pragma unroll
      for (int hf = 0; hf < X::HALVES; hf++) {
        const uint32_t wr[4] = {W[hf].x, W[hf].y, W[hf].z, W[hf].w};
#pragma unroll
        for (int b8 = 0; b8 < 8; b8++) {
          const int S = hf * 8 + b8;
          const uint32_t a = (b8 & 1) ? (wr[b8 >> 1] >> 16) + base : ((wr[b8 >> 1] & 0xFFFFu) | base);
          const uint64_t v0 = lds64<0>(a), v1 = lds64<32768>(a), v2 = lds64<65536>(a), v3 = lds64<98304>(a);
          const uint64_t kp2 = pack2(keep[S], keep[S]), cp2 = pack2(capf[S], capf[S]);
          acc[0] = ffma2(acc[0], kp2, v0);
          acc[1] = ffma2(acc[1], kp2, v1);
          acc[2] = ffma2(acc[2], kp2, v2);
          acc[3] = ffma2(acc[3], kp2, v3);
          done[0] = ffma2(acc[0], cp2, done[0]);
          done[1] = ffma2(acc[1], cp2, done[1]);
          done[2] = ffma2(acc[2], cp2, done[2]);
          done[3] = ffma2(acc[3], cp2, done[3]);
        }
      }
      if (t >= 1 && id < n32) {
        float dv[8];
        const uint64_t n2 = pack2(nrm0, nrm0);
        bool anyp = false;
#pragma unroll
        for (int tt = 0; tt < 4; tt++) {
          uint64_t dd = done[tt];
          if (NORMS) dd = fadd2(dd, n2);                   // + dbnorms[i] last, pairwise_byte.cpp:74
          dv[2 * tt] = __uint_as_float((uint32_t)dd);
          dv[2 * tt + 1] = __uint_as_float((uint32_t)(dd >> 32));
          anyp |= (dv[2 * tt] <= tau[2 * tt]) | (dv[2 * tt + 1] <= tau[2 * tt + 1]);
        }
        if (anyp) {
#pragma unroll
          for (int i = 0; i < 8; i++) {
            if (dv[i] <= tau[i]) {
              const int q = (i >> 1) * (2 * X::G) + g * 2 + (i & 1);
              const uint64_t key = make_key(dv[i], id);
              if (!p.lb || key > lb_s[q]) {
                const int pos = atomicAdd(&cnt_s[q], 1);
                cand[(size_t)q * p.cap + pos] = key;
                if (pos >= (SPEC ? softq_s[q] : p.soft) || (tau[i] == inf && pos + 1 >= p.k)) atomicExch(&flag_s, 1);
              }
            }
          }
        }
      }
#pragma unroll
      for (int tt = 0; tt < 4; tt++) done[tt] = 0ull;
#pragma unroll
      for (int hf = 0; hf < X::HALVES; hf++) W[hf] = Wn[hf];
      nrm0 = nrm1;
      fp += X::HALVES * X::NS;
      np += X::NS;
      id += X::NS;
      // while some tau is still +inf every code of every warp is a candidate: do not wait a period to react
      if (raised || (warm && lds_volatile(flag_addr))) {
        service(wdone + t * X::NS);
#pragma unroll
        for (int i = 0; i < 8; i++) tau[i] = tau_x[i];
        warm = warm_x;
      }
    }
  }
  __syncwarp();
  if (lane == 0) atomicAdd(&nfin_s, 1);
  __syncwarp();
  while (!service(wdone)) {
  }

  }

  // final phase: small buffers by their own warp (all at once), the rest block-wide
  final_phase = true;
  if (w < QB) {
    const bool mine = q0 + w < p.nq && cnt_s[w] <= kWarpKeys;
    if (mine) {
      const int c = cnt_s[w];
      const int keep = warp_sort_keep(w);
      __syncwarp();
      if (lane == 0) verify(w, c, wslice[max(keep, 1) - 1]);
      uint64_t* out = p.part + ((size_t)slice * p.nq + q0 + w) * p.k;
      for (int i = lane; i < p.k; i += 32) out[i] = i < keep ? wslice[i] : ~0ull;
    }
    if (lane == 0) finq_s[w] = mine;
  }
  block_sync();
  if (p.k <= 1024) {
    // large buffers, k <= 1024: block-wide selection of the k smallest of each (one query after the other), then
    // ALL queries are sorted at once, one warp each, in the (now dead) LUT tile: 16 x 8 KB
    for (int q = 0; q < QB; q++)
      if (q0 + q < p.nq && !finq_s[q]) compact(q);
    block_sync();
    uint64_t* wsort = reinterpret_cast<uint64_t*>(smem_raw + (lut_addr - smem_u32(smem_raw))) + w * 1024;
    if (w < QB && q0 + w < p.nq && !finq_s[w]) {
      const int c = cnt_s[w];                                  // <= k
      const uint64_t* cq = cand + (size_t)w * p.cap;
      const int np2 = max(64, pow2ceil(c));
      for (int t = lane; t < np2; t += 32) wsort[t] = t < c ? cq[t] : ~0ull;
      __syncwarp();
      warp_bitonic(wsort, np2);
      if (lane == 0) verify(w, c, wsort[max(c, 1) - 1]);
      uint64_t* out = p.part + ((size_t)slice * p.nq + q0 + w) * p.k;
      for (int i = lane; i < p.k; i += 32) out[i] = i < c ? wsort[i] : ~0ull;
    }
  } else {
    for (int q = 0; q < QB; q++) {
      if (q0 + q >= p.nq) break;
      if (finq_s[q]) continue;
      finalize(q);
      const int c = cnt_s[q];
      uint64_t* out = p.part + ((size_t)slice * p.nq + q0 + q) * p.k;
      for (int i = tid; i < p.k; i += NT) out[i] = i < c ? sortbuf[i] : ~0ull;
      block_sync();
    }
  }
  if (SPEC) {
    block_sync();
    if (tid == 0) p.redo[blockIdx.y * gridDim.x + blockIdx.x] = (spec && fail_s) ? 1 : 0;
  }
}

// ------------------------------------------------------------------------------------------------------
// K6: merge of S sorted lists per query into the global top-k, by the (dist, id) total order.
//   keys != null : lists are 64-bit keys [S][nq][k] from scanx_kernel (ids local; id_add makes them final)
//   else         : lists are (dists, idx) [S][nq][k] (already-final ids; the multi-GPU exchange format)
// One block per query; S*k keys sorted in shared memory.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) merge_kernel(const uint64_t* __restrict__ keys, const float* __restrict__ din,
                                                    const int32_t* __restrict__ iin, int S, int nq, int k,
                                                    float* __restrict__ dout, int32_t* __restrict__ iout,
                                                    int64_t id_add, int ldo) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* s = reinterpret_cast<uint64_t*>(smem_raw);
  const int q = blockIdx.x;
  const int total = S * k;
  const int np2 = pow2ceil(total);
  for (int t = threadIdx.x; t < np2; t += blockDim.x) {
    uint64_t key = ~0ull;
    if (t < total) {
      size_t src = ((size_t)(t / k) * nq + q) * k + (t % k);
      key = keys ? keys[src] : make_key(din[src], (uint32_t)iin[src]);
    }
    s[t] = key;
  }
  __syncthreads();
  if (S > 1) block_bitonic_sort(s, np2);
  for (int t = threadIdx.x; t < k; t += blockDim.x) {
    uint64_t key = s[t];
    dout[(size_t)q * ldo + t] = ordered_to_f32((uint32_t)(key >> 32));
    iout[(size_t)q * ldo + t] = (int32_t)((int64_t)(uint32_t)key + id_add);
  }
}

// The same merge for already-SORTED lists (which is what both callers hand over) without sorting anything: a tree of
// pairwise rank merges inside shared memory.  Level by level, lists 2i and 2i+1 are merged into their common top-k: the
// element at index t of one list goes to position t + (number of smaller elements in the other list, by binary search);
// positions >= k are dropped.  S*k*10 shared loads instead of the bitonic network's ~S*k*log^2 compare-exchanges: the
// 8-GPU k = 1000 merge (8000 keys per query) drops from 4.6 ms to well under 1 ms per 10k queries.
// Shared memory: A = S*k keys, B = ceil(S/2)*k keys.
__global__ void __launch_bounds__(256) merge_tree_kernel(const uint64_t* __restrict__ keys, const float* __restrict__ din,
                                                         const int32_t* __restrict__ iin, int S, int nq, int k,
                                                         float* __restrict__ dout, int32_t* __restrict__ iout,
                                                         int64_t id_add, int ldo) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* A = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* B = A + (size_t)S * k;
  const int q = blockIdx.x;
  for (int t = threadIdx.x; t < S * k; t += blockDim.x) {
    const size_t src = ((size_t)(t / k) * nq + q) * k + (t % k);
    A[t] = keys ? keys[src] : make_key(din[src], (uint32_t)iin[src]);
  }
  __syncthreads();
  int L = S;
  while (L > 1) {
    const int pairs = L >> 1;
    for (int t = threadIdx.x; t < pairs * k; t += blockDim.x) {
      const int pi = t / k, i = t - pi * k;
      const uint64_t* a = A + (size_t)(2 * pi) * k;
      const uint64_t* b = a + k;
      uint64_t* o = B + (size_t)pi * k;
      {
        const uint64_t key = a[i];
        int lo = 0, hi = k - i;                          // more than k-i-1 smaller keys in b: position >= k anyway
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (b[mid] < key) lo = mid + 1; else hi = mid;
        }
        if (i + lo < k) o[i + lo] = key;
      }
      {
        const uint64_t key = b[i];
        int lo = 0, hi = k - i;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (a[mid] <= key) lo = mid + 1; else hi = mid;   // a wins ties: equal keys keep list order
        }
        if (i + lo < k) o[i + lo] = key;
      }
    }
    if (L & 1)
      for (int t = threadIdx.x; t < k; t += blockDim.x) B[(size_t)pairs * k + t] = A[(size_t)(L - 1) * k + t];
    __syncthreads();
    L = (L + 1) >> 1;
    for (int t = threadIdx.x; t < L * k; t += blockDim.x) A[t] = B[t];      // next level reads A again
    __syncthreads();
  }
  for (int t = threadIdx.x; t < k; t += blockDim.x) {
    const uint64_t key = A[t];
    dout[(size_t)q * ldo + t] = ordered_to_f32((uint32_t)(key >> 32));
    iout[(size_t)q * ldo + t] = (int32_t)((int64_t)(uint32_t)key + id_add);
  }
}

// ---- general merge (any S*k): pairwise rank merges of sorted key lists in global memory --------------------------
// (dists, ids) -> 64-bit keys of the (dist, id) total order
__global__ void pairs_to_keys_kernel(const float* __restrict__ d, const int32_t* __restrict__ i,
                                     uint64_t* __restrict__ keys, size_t total) {
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x)
    keys[t] = make_key(d[t], (uint32_t)i[t]);
}
// out[q] = the k smallest of A[q] U B[q] (both sorted ascending, k keys each), sorted.  Element t of A goes to position
// t + #{b < a}; element t of B to t + #{a <= b} (A wins ties, so equal keys from two lists keep list order); positions
// >= k are dropped.  grid (ceil(k/256), nq).
__global__ void __launch_bounds__(256) merge2_kernel(const uint64_t* __restrict__ A, const uint64_t* __restrict__ B,
                                                     uint64_t* __restrict__ out, int k) {
  const size_t q = blockIdx.y;
  const uint64_t* a = A + q * k;
  const uint64_t* b = B + q * k;
  uint64_t* o = out + q * k;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= k) return;
  {
    const uint64_t key = a[t];
    int lo = 0, hi = min(k, k - t);                   // more than k-t-1 smaller keys in B: position >= k anyway
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (b[mid] < key) lo = mid + 1; else hi = mid;
    }
    if (t + lo < k) o[t + lo] = key;
  }
  {
    const uint64_t key = b[t];
    int lo = 0, hi = min(k, k - t);
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (a[mid] <= key) lo = mid + 1; else hi = mid;
    }
    if (t + lo < k) o[t + lo] = key;
  }
}
__global__ void keys_to_pairs_kernel(const uint64_t* __restrict__ keys, float* __restrict__ dout,
                                     int32_t* __restrict__ iout, int64_t id_add, int k, int ldo) {
  const size_t q = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= k) return;
  const uint64_t key = keys[q * k + t];
  dout[q * ldo + t] = ordered_to_f32((uint32_t)(key >> 32));
  iout[q * ldo + t] = (int32_t)((int64_t)(uint32_t)key + id_add);
}

// a search whose LUT held a non-finite / overflowing entry returns NaN distances and id -1 everywhere (device-pointer
// calls cannot return a status without synchronising; host-pointer calls also fail with RAYUELA_ERR_ARG)
__global__ void poison_kernel(const int* __restrict__ bad, float* __restrict__ d, int32_t* __restrict__ i, size_t total) {
  if (*bad == 0) return;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    d[t] = __int_as_float(0x7fc00000);
    i[t] = -1;
  }
}

// lower bound for the next pass of a k > one-pass search: the last key the previous pass returned
__global__ void lower_bound_kernel(const float* __restrict__ d, const int32_t* __restrict__ i, int nq, int ldo,
                                   int col, int64_t id_add, uint64_t* __restrict__ lb) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < nq) lb[q] = make_key(d[(size_t)q * ldo + col], (uint32_t)((int64_t)i[(size_t)q * ldo + col] - id_add));
}

}  // namespace ryl

// ======================================================================================================
// host side
// ======================================================================================================
using namespace ryl;

struct rayuela_index {
  int kind = 0, m = 0, h = 0, period = 0, device = 0;   // period: 8 (m <= 8) or 16 codebooks per lane cycle
  int64_t n = 0, id_offset = 0;
  int64_t nchunks = 0;   // 1024-code warp chunks of the skewed layout
  DevBuf norms, skew;    // fp32 norms (LSQ); skewed offset fields (skew_fields_kernel)
  DevBuf codes;          // raw codes [n][m]: the pre-filter scan re-evaluates its survivors from them
  DevBuf rot;            // rotated offset fields of the pre-filter scan (rot_fields_kernel)
  float nmin = 0.f, nmax = 0.f, ncap = 0.f;   // range of the norms (the pre-filter scan quantises them on the fly) and
                                              // the value above which their quantised image saturates
  bool q16_ok = false;   // norms finite: the 16-bit pre-filter scan may be used
  // multi-device parent (rayuela_init / RAYUELA_B200_DEVICES, host arrays): one shard per device slot, no own buffers
  std::vector<rayuela_index*> shards;
  std::vector<DeviceSlot> slots;
};

static int host_pow2ceil(int x) {
  int p = 2;
  while (p < x) p <<= 1;
  return p;
}

static int index_create_single(rayuela_index** out, int kind, const uint8_t* codes, const float* dbnorms, int64_t n,
                               int m, int h, int64_t id_offset, unsigned flags, cudaStream_t s) {
  RYL_ARG(n >= 1 && n < (1ll << 32) - 4096, "index_create: n must be in 1..2^32-4097 per index shard");
  const bool dev = flags & RAYUELA_DEVICE_PTRS;
  rayuela_index* ix = new rayuela_index();
  ix->kind = kind;
  ix->m = m;
  ix->h = h;
  ix->period = m <= 8 ? 8 : 16;
  ix->n = n;
  ix->id_offset = id_offset;
  cudaGetDevice(&ix->device);
  auto body = [&]() -> int {
    RYL_TRY(ix->codes.alloc((size_t)n * m, s));
    RYL_CUDA(cudaMemcpyAsync(ix->codes.p, codes, (size_t)n * m, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
    struct { const uint8_t* d; } raw{ix->codes.as<uint8_t>()};
    ix->nchunks = (n + kChunkCodes - 1) / kChunkCodes;
    if (ix->period == 8) {
      const int64_t words = ix->nchunks * ScanX<8>::PERIODS * ScanX<8>::HALVES * ScanX<8>::NS;
      RYL_TRY(ix->skew.alloc((size_t)words * sizeof(uint4), s));
      int blocks = (int)std::min<int64_t>((words + 255) / 256, 148 * 16);
      RYL_LAUNCH(skew_fields_kernel<8>, blocks, 256, 0, s, raw.d, ix->skew.as<uint4>(), n, m, ix->nchunks);
      const int64_t rwords = ix->nchunks * ScanX<8>::L * ScanX<8>::HALVES * ScanX<8>::NS;
      RYL_TRY(ix->rot.alloc((size_t)rwords * sizeof(uint4), s));
      RYL_LAUNCH(rot_fields_kernel<8>, blocks, 256, 0, s, raw.d, ix->rot.as<uint4>(), n, m, ix->nchunks);
    } else {
      const int64_t words = ix->nchunks * ScanX<16>::PERIODS * ScanX<16>::HALVES * ScanX<16>::NS;
      RYL_TRY(ix->skew.alloc((size_t)words * sizeof(uint4), s));
      int blocks = (int)std::min<int64_t>((words + 255) / 256, 148 * 16);
      RYL_LAUNCH(skew_fields_kernel<16>, blocks, 256, 0, s, raw.d, ix->skew.as<uint4>(), n, m, ix->nchunks);
      const int64_t rwords = ix->nchunks * ScanX<16>::L * ScanX<16>::HALVES * ScanX<16>::NS;
      RYL_TRY(ix->rot.alloc((size_t)rwords * sizeof(uint4), s));
      RYL_LAUNCH(rot_fields_kernel<16>, blocks, 256, 0, s, raw.d, ix->rot.as<uint4>(), n, m, ix->nchunks);
    }
    if (kind == RAYUELA_SCAN_LSQ) {
      RYL_TRY(ix->norms.alloc((size_t)n * sizeof(float), s));
      RYL_CUDA(cudaMemcpyAsync(ix->norms.p, dbnorms, (size_t)n * sizeof(float),
                               dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
      DevBuf rng, sums;
      RYL_TRY(rng.alloc(2 * sizeof(unsigned int), s));
      RYL_TRY(sums.alloc(3 * sizeof(double), s));
      const unsigned int init[2] = {0xFFFFFFFFu, 0u};
      RYL_CUDA(cudaMemcpyAsync(rng.p, init, sizeof init, cudaMemcpyHostToDevice, s));
      RYL_CUDA(cudaMemsetAsync(sums.p, 0, 3 * sizeof(double), s));
      RYL_LAUNCH(norm_range_kernel, (int)std::min<int64_t>((n + 255) / 256, sm_count() * 8), 256, 0, s,
                 ix->norms.as<float>(), n, rng.as<unsigned int>(), sums.as<double>());
      unsigned int got[2];
      double sm[3];
      RYL_CUDA(cudaMemcpyAsync(got, rng.p, sizeof got, cudaMemcpyDeviceToHost, s));
      RYL_CUDA(cudaMemcpyAsync(sm, sums.p, sizeof sm, cudaMemcpyDeviceToHost, s));
      RYL_CUDA(cudaStreamSynchronize(s));
      ix->nmin = ordered_to_f32(got[0]);
      ix->nmax = ordered_to_f32(got[1]);
      ix->q16_ok = got[0] <= got[1] && std::isfinite(ix->nmin) && std::isfinite(ix->nmax);
      ix->ncap = ix->nmax;
      if (ix->q16_ok && sm[2] >= 2) {
        const double mean = sm[0] / sm[2], var = std::max(0.0, sm[1] / sm[2] - mean * mean);
        const double cap = mean + 4.0 * std::sqrt(var);
        if (cap < (double)ix->nmax && cap > (double)ix->nmin) ix->ncap = (float)cap;
      }
    } else {
      ix->q16_ok = true;
    }
    RYL_CUDA(cudaStreamSynchronize(s));
    return RAYUELA_OK;
  };
  int rc = body();
  if (rc != RAYUELA_OK) {
    delete ix;
    return rc;
  }
  *out = ix;
  return RAYUELA_OK;
}

extern "C" int rayuela_index_free(rayuela_index* ix) {
  if (ix) {
    int cur = 0;
    cudaGetDevice(&cur);
    for (rayuela_index* sh : ix->shards) {
      if (!sh) continue;
      cudaSetDevice(sh->device);
      rayuela_index_free(sh);
    }
    if (!ix->shards.empty()) cudaSetDevice(cur);
    ix->norms.release();
    ix->skew.release();
    ix->codes.release();
    ix->rot.release();
    delete ix;
  }
  return RAYUELA_OK;
}

extern "C" int rayuela_index_create(rayuela_index** out, int kind, const uint8_t* codes, const float* dbnorms,
                                    int64_t n, int m, int h, int64_t id_offset, unsigned flags, void* stream) {
  RYL_ARG(out != nullptr, "index_create: out is null");
  RYL_ARG(kind >= 0 && kind <= 2, "index_create: unknown kind");
  RYL_ARG(h >= 1 && h <= kH, "index_create: h must be in 1..256 (one byte per codebook)");
  RYL_ARG(m >= 1 && m <= 16, "index_create: m must be in 1..16");
  RYL_ARG(n >= 1, "index_create: n must be positive");
  RYL_ARG(codes != nullptr, "index_create: codes is null");
  RYL_ARG(kind != RAYUELA_SCAN_LSQ || dbnorms != nullptr, "index_create: LSQ scan needs dbnorms");
  if (!(flags & RAYUELA_DEVICE_PTRS)) {
    const std::vector<DeviceSlot> slots = device_slots();
    const int D = (int)slots.size();
    if (D > 1 && n >= (int64_t)D * 1024) {
      // base-sharded index: slot i holds rows splitarray(n, D)[i] and returns global ids (id_offset + slice start)
      rayuela_index* ix = new rayuela_index();
      ix->kind = kind; ix->m = m; ix->h = h; ix->period = m <= 8 ? 8 : 16; ix->n = n; ix->id_offset = id_offset;
      cudaGetDevice(&ix->device);
      ix->slots = slots;
      ix->shards.assign(D, nullptr);
      int rc = for_each_slot(slots, [&](int i) -> int {
        int64_t a, b;
        split_range(n, D, i, &a, &b);
        return index_create_single(&ix->shards[i], kind, codes + (size_t)a * m, dbnorms ? dbnorms + a : nullptr, b - a,
                                   m, h, id_offset + a, 0, slots[i].stream);
      });
      if (rc != RAYUELA_OK) {
        rayuela_index_free(ix);
        return rc;
      }
      *out = ix;
      return RAYUELA_OK;
    }
  }
  return index_create_single(out, kind, codes, dbnorms, n, m, h, id_offset, flags, (cudaStream_t)stream);
}

// Merge of S sorted lists per query.  Up to 16384 keys per query: one block per query sorts them in shared memory
// (merge_kernel).  Beyond that (e.g. 2 shards at the reference's default k = 10000, 8 shards at k > 2048): a tree of
// pairwise rank merges in global memory (merge2_kernel), any S and k.
static int merge_lists(const uint64_t* keys, const float* din, const int32_t* iin, int S, int nq, int k, float* dout,
                       int32_t* iout, int64_t id_add, cudaStream_t s, int ldo = 0) {
  if (ldo == 0) ldo = k;
  // sorted lists, more than a handful of keys: rank-merge tree in shared memory (A = S*k keys + B = ceil(S/2)*k keys)
  const size_t tree_smem = ((size_t)S * k + (size_t)((S + 1) / 2) * k) * sizeof(uint64_t);
  if (S >= 2 && (int64_t)S * k > 1024 && tree_smem <= 200 * 1024) {
    RYL_CUDA(cudaFuncSetAttribute(merge_tree_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tree_smem));
    RYL_LAUNCH(merge_tree_kernel, nq, 256, tree_smem, s, keys, din, iin, S, nq, k, dout, iout, id_add, ldo);
    return RAYUELA_OK;
  }
  const size_t smem = (size_t)host_pow2ceil(S * k) * sizeof(uint64_t);
  if ((int64_t)S * k <= 16384 && smem <= 200 * 1024) {
    RYL_CUDA(cudaFuncSetAttribute(merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RYL_LAUNCH(merge_kernel, nq, 256, smem, s, keys, din, iin, S, nq, k, dout, iout, id_add, ldo);
    return RAYUELA_OK;
  }
  const size_t per = (size_t)nq * k;                       // keys per list
  DevBuf k0, ping, pong;
  const uint64_t* cur = keys;
  if (!cur) {
    RYL_TRY(k0.alloc((size_t)S * per * sizeof(uint64_t), s));
    RYL_LAUNCH(pairs_to_keys_kernel, sm_count() * 8, 256, 0, s, din, iin, k0.as<uint64_t>(), (size_t)S * per);
    cur = k0.as<uint64_t>();
  }
  int L = S;
  RYL_TRY(ping.alloc((size_t)((L + 1) / 2) * per * sizeof(uint64_t), s));
  if (L > 2) RYL_TRY(pong.alloc((size_t)((L + 3) / 4) * per * sizeof(uint64_t), s));
  uint64_t* dst = ping.as<uint64_t>();
  uint64_t* other = pong.as<uint64_t>();
  const dim3 grid((k + 255) / 256, nq);
  while (L > 1) {
    const int pairs = L / 2;
    for (int i = 0; i < pairs; i++)
      RYL_LAUNCH(merge2_kernel, grid, 256, 0, s, cur + (size_t)(2 * i) * per, cur + (size_t)(2 * i + 1) * per,
                 dst + (size_t)i * per, k);
    if (L & 1)
      RYL_CUDA(cudaMemcpyAsync(dst + (size_t)pairs * per, cur + (size_t)(L - 1) * per, per * sizeof(uint64_t),
                               cudaMemcpyDeviceToDevice, s));
    L = (L + 1) / 2;
    cur = dst;
    std::swap(dst, other);
  }
  RYL_LAUNCH(keys_to_pairs_kernel, grid, 256, 0, s, cur, dout, iout, id_add, k, ldo);
  return RAYUELA_OK;
}

// all arrays on the current device; bad: device int, set when a LUT entry is unusable (see lut_kernel)
static int index_search_dev(rayuela_index* ix, const float* q_dev, const float* cb_dev, int nq, int d, int k,
                            float* d_dev, int32_t* i_dev, int* bad, cudaStream_t s, bool fast_lut = false) {
  const int m = ix->m, h = ix->h, mh = m * h;
  // opt-in: tensor-core LUT for the LSQ scan (codebooks split / packed once per search)
  fast_lut = fast_lut && ix->kind == RAYUELA_SCAN_LSQ && h == kH && unary_tc_supported(d, mh);
  DevBuf Cp_d, zero_d;
  if (fast_lut) {
    RYL_TRY(unary_tc_pack_codebooks(cb_dev, d, mh, &Cp_d, s));
    RYL_TRY(zero_d.alloc((size_t)mh * sizeof(float), s));
    RYL_CUDA(cudaMemsetAsync(zero_d.p, 0, zero_d.bytes, s));
  }
  const bool pq = ix->kind == RAYUELA_SCAN_PQ;
  const int len = pq ? d / m : d;
  const int period = ix->period;
  const int QT = period == 16 ? ScanX<16>::QB : ScanX<8>::QB;               // queries per block
  const int adds = period == 16 ? ScanX<16>::ADDS : ScanX<8>::ADDS;         // keys one query can gain per block-period
  // One pass returns at most kmax results per query (shared-memory selection buffer).  Larger k -- the reference's
  // default is k = 10000 (src/Linscan.jl:10) -- takes ceil(k / kmax) passes: pass p keeps only keys strictly
  // greater than the last key of pass p-1, which is exact because (dist, id) keys are a total order.
  const int kmax = 4096;
  const int64_t id_add = (pq ? 0 : 1) + ix->id_offset;  // linscan_aqd.cpp:88 vs pairwise_byte.cpp:76
  const float* norms = ix->kind == RAYUELA_SCAN_LSQ ? ix->norms.as<float>() : nullptr;

  // Query chunks: whole waves of query tiles first (one block per SM per wave, base unsliced), then the
  // remainder, which is sliced along the base so the last partial wave still fills the machine.
  const int sms = sm_count();
  int nqc = 0;
  for (int qb = 0; qb < nq; qb += nqc) {
    const int tiles_left = (nq - qb + QT - 1) / QT;
    const int max_tiles = std::max(sms, (16384 / QT) / sms * sms);
    nqc = nq - qb;
    if (tiles_left > sms) nqc = std::min(nqc, std::min(tiles_left / sms * sms, max_tiles) * QT);
    const int qtiles = (nqc + QT - 1) / QT;
    DevBuf lut, lb;
    RYL_TRY(lut.alloc((size_t)qtiles * kLutTileBytes, s));
    if (m < period || nqc % QT || h < kH) RYL_CUDA(cudaMemsetAsync(lut.p, 0, lut.bytes, s));   // zero rows for k >= m, c >= h
    dim3 lg(m * ((h + 31) / 32), (nqc + 31) / 32);
    const float* qptr = q_dev + (size_t)qb * d;
    const int tiled = period == 16 ? 2 : 1;
    if (fast_lut) {
      DevBuf rowmajor;
      RYL_TRY(rowmajor.alloc((size_t)nqc * mh * sizeof(float), s));
      RYL_TRY(unary_tc_launch(qptr, Cp_d, zero_d.as<float>(), rowmajor.as<float>(), nullptr, nqc, d, mh, s));
      RYL_LAUNCH(lut_relayout_kernel, sm_count() * 8, 256, 0, s, rowmajor.as<float>(), lut.as<float>(), nqc, mh, tiled, bad);
    } else if (ix->kind == RAYUELA_SCAN_LSQ)
      RYL_LAUNCH(lut_kernel<RAYUELA_SCAN_LSQ>, lg, 256, 0, s, qptr, cb_dev, lut.as<float>(), nqc, d, len, mh, tiled, bad, h);
    else if (ix->kind == RAYUELA_SCAN_CQ)
      RYL_LAUNCH(lut_kernel<RAYUELA_SCAN_CQ>, lg, 256, 0, s, qptr, cb_dev, lut.as<float>(), nqc, d, len, mh, tiled, bad, h);
    else
      RYL_LAUNCH(lut_kernel<RAYUELA_SCAN_PQ>, lg, 256, 0, s, qptr, cb_dev, lut.as<float>(), nqc, d, len, mh, tiled, bad, h);
    if (k > kmax) RYL_TRY(lb.alloc((size_t)nqc * sizeof(uint64_t), s));
    // QPF: quantised twin of the tiles (exact results, half the shared-memory bytes per lookup; see scanx_kernel)
    const char* q16_env = getenv("RAYUELA_B200_SCAN_PREFILTER");                        // tuning knob: 0 disables
    // (large k: the survivors' exact re-evaluation outgrows what the narrower loop saves once the quantisation window holds
    // about as many codes as the result list -- measured crossover between k = 256 and 1000 on an LSQ-encoded base, beyond
    // 1000 on isotropic random codes)
    int q16_maxk = 256;
    if (const char* e = getenv("RAYUELA_B200_SCAN_PREFILTER_MAXK")) q16_maxk = atoi(e);     // tuning knob
    const bool q16 = ix->q16_ok && !(q16_env && atoi(q16_env) == 0) && std::min(k, kmax) <= q16_maxk;
    DevBuf lutq, tilep, qoff, qmu;
    if (q16) {
      RYL_TRY(lutq.alloc((size_t)qtiles * 16384 * sizeof(float), s));
      RYL_TRY(tilep.alloc((size_t)qtiles * sizeof(float4), s));
      RYL_TRY(qoff.alloc((size_t)qtiles * QT * sizeof(double), s));
      RYL_TRY(qmu.alloc((size_t)qtiles * QT * sizeof(int), s));
      if (period == 16) {
        RYL_CUDA(cudaFuncSetAttribute(lut_quant_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
        RYL_LAUNCH(lut_quant_kernel<16>, qtiles, 512, 131072, s, lut.as<float>(), lutq.as<float>(), tilep.as<float4>(),
                   qoff.as<double>(), qmu.as<int>(), m, h, ix->nmin, ix->nmax, ix->ncap, norms ? 1 : 0);
      } else {
        RYL_CUDA(cudaFuncSetAttribute(lut_quant_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
        RYL_LAUNCH(lut_quant_kernel<8>, qtiles, 512, 131072, s, lut.as<float>(), lutq.as<float>(), tilep.as<float4>(),
                   qoff.as<double>(), qmu.as<int>(), m, h, ix->nmin, ix->nmax, ix->ncap, norms ? 1 : 0);
      }
    }

    for (int koff = 0; koff < k; koff += kmax) {
      const int kp = std::min(kmax, k - koff);
      const uint64_t* lbp = koff > 0 ? lb.as<uint64_t>() : nullptr;
      // Event-driven compaction: a buffer is compacted (k smallest kept, tau tightened) once it holds more than
      // `soft` keys -- up to 4k: fewer, relatively cheaper selections; measured optimum, gpurun r2_soft.log -- and
      // the capacity leaves room for the periods of appends that can land before every warp has reacted to the flag.
      // (small k: 384, so that a buffer past the limit still fits the 512-key slice one warp can compact alone)
      // (pre-filter loop: + the 16 x 64 survivors that may sit in the warps' rings when the flag goes up)
      const int slack = 3 * adds + (q16 ? kScanWarps * 64 : 0);
      int soft = std::max(384, std::min(4 * kp, kScanSortKeys - slack));
      if (const char* e = getenv("RAYUELA_B200_SCAN_SOFT"))   // tuning knob
        soft = std::max(kp, std::min(atoi(e), kScanSortKeys - slack));
      const int cap = soft + slack;
      // sort buffer + up to 32 KB of padding so the LUT tile starts on a 32 KB boundary + the tile
      const size_t smem = (size_t)kScanSortKeys * sizeof(uint64_t) + 32768 + (size_t)kLutTileBytes;
      // DB slices.  A launch of qtiles x S blocks takes ceil(qtiles*S / SMs) waves of (1/S + c) base passes each,
      // c = per-block fixed cost (LUT staging, threshold warm-up, final selection) ~ 0.065 + 0.0005 k of a pass
      // (fit to tools/slices_probe.py): whole waves of tiles stay unsliced, a partial wave is cut so that it fills
      // the machine once (33 tiles -> 4 slices, 100 tiles -> 4 slices = 2.7 waves of quarter passes, 1 tile -> 61)
      const int64_t unit = (int64_t)kChunkCodes * kScanWarps;   // codes per block round
      const int smax = (int)std::min<int64_t>(std::min<int64_t>(std::max<int64_t>(1, ix->n / unit), 4 * sms),
                                              std::max(1, 16384 / kp));
      int S = 1;
      {
        const double c = 0.065 + 0.0005 * kp;
        double best = 1e30;
        for (int cand_s = 1; cand_s <= smax; cand_s++) {
          const double waves = (double)(((int64_t)qtiles * cand_s + sms - 1) / sms);
          const double t = waves * (1.0 / cand_s + c);
          if (t < best * 0.999) {
            best = t;
            S = cand_s;
          }
        }
      }
      if (const char* e = getenv("RAYUELA_B200_SCAN_SLICES")) S = std::max(1, atoi(e));   // tuning knob
      S = std::min(S, std::max(1, std::min(smax, 16384 / kp)));
      int64_t slice_len = (ix->n + S - 1) / S;
      slice_len = (slice_len + unit - 1) / unit * unit;
      S = (int)((ix->n + slice_len - 1) / slice_len);

      DevBuf cand, part;
      RYL_TRY(cand.alloc((size_t)S * qtiles * QT * cap * sizeof(uint64_t), s));
      RYL_TRY(part.alloc((size_t)S * nqc * kp * sizeof(uint64_t), s));
      ScanXParams p;
      p.F = ix->skew.as<uint4>();
      p.norms = norms;
      p.lut = lut.as<float>();
      p.cand = cand.as<uint64_t>();
      p.part = part.as<uint64_t>();
      p.lb = lbp;
      p.n = ix->n;
      p.nchunks = ix->nchunks;
      p.chunks_per_slice = slice_len / kChunkCodes;
      p.nq = nqc;
      p.k = kp;
      p.cap = cap;
      p.soft = soft;
      p.piggy = kp + (soft - kp) / 6;   // a service() also compacts every buffer already a sixth of the way to the soft limit
                                        // (sweep at k = 16 / 100 / 1000 / 4000, profiles/r2_scan_knob_sweeps.txt: +2..5 % over 1/2)
      p.spec_soft = 2 * kp + adds;
      if (const char* e = getenv("RAYUELA_B200_SCAN_SPECSOFT")) p.spec_soft = std::max(kp, std::min(atoi(e), soft));   // tuning knob
      if (const char* e = getenv("RAYUELA_B200_SCAN_PIGGY")) p.piggy = std::max(kp, std::min(atoi(e), soft));          // tuning knob
      p.tau0 = std::numeric_limits<float>::infinity();
      if (const char* e = getenv("RAYUELA_B200_SCAN_TAU0")) p.tau0 = (float)atof(e);   // measurement aid only
      // speculative thresholds (verified; see scanx_kernel): worth it once a compaction is more than a warp's work
      const char* spec_env = getenv("RAYUELA_B200_SCAN_SPEC");                          // tuning knob: 0 disables
      p.spec = (kp >= 16 && lbp == nullptr && !(spec_env && atoi(spec_env) == 0)) ? 1 : 0;
      p.pass = 0;
      DevBuf redo;
      RYL_TRY(redo.alloc((size_t)S * qtiles * sizeof(int), s));
      p.redo = redo.as<int>();
      p.lutq = lutq.as<float>();
      p.Fq = ix->rot.as<uint4>();
      p.tilep = tilep.as<float4>();
      p.qoff = qoff.as<double>();
      p.qmu = qmu.as<int>();
      p.codes = ix->codes.as<uint8_t>();
      p.m = m;
      DevBuf qstats;
      p.qstats = nullptr;
      if (q16 && getenv("RAYUELA_B200_SCAN_STATS")) {
        RYL_TRY(qstats.alloc(2 * sizeof(unsigned long long), s));
        RYL_CUDA(cudaMemsetAsync(qstats.p, 0, 2 * sizeof(unsigned long long), s));
        p.qstats = qstats.as<unsigned long long>();
      }
#define RYL_SCANX(PP, NN, SS, QQ)                                                                                    \
  {                                                                                                                  \
    RYL_CUDA(cudaFuncSetAttribute(scanx_kernel<PP, NN, SS, QQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    RYL_LAUNCH((scanx_kernel<PP, NN, SS, QQ>), dim3(qtiles, S), kScanWarps * 32, smem, s, p);                       \
    if (SS) { /* redo launch: blocks whose speculation failed rerun exactly, all others exit at once */             \
      p.pass = 1;                                                                                                    \
      RYL_LAUNCH((scanx_kernel<PP, NN, SS, QQ>), dim3(qtiles, S), kScanWarps * 32, smem, s, p);                     \
    }                                                                                                                \
  }
#define RYL_SCANX_Q(PP, NN, SS) { if (q16) RYL_SCANX(PP, NN, SS, true) else RYL_SCANX(PP, NN, SS, false) }
      if (period == 16) {
        if (norms) { if (p.spec) RYL_SCANX_Q(16, true, true) else RYL_SCANX_Q(16, true, false) }
        else { if (p.spec) RYL_SCANX_Q(16, false, true) else RYL_SCANX_Q(16, false, false) }
      } else {
        if (norms) { if (p.spec) RYL_SCANX_Q(8, true, true) else RYL_SCANX_Q(8, true, false) }
        else { if (p.spec) RYL_SCANX_Q(8, false, true) else RYL_SCANX_Q(8, false, false) }
      }
#undef RYL_SCANX_Q
#undef RYL_SCANX
      if (p.qstats) {
        unsigned long long h2[2];
        RYL_CUDA(cudaMemcpyAsync(h2, qstats.p, sizeof h2, cudaMemcpyDeviceToHost, s));
        RYL_CUDA(cudaStreamSynchronize(s));
        fprintf(stderr, "[scan stats] nq=%d k=%d S=%d: pre-filter survivors %.1f per query, accepted %.1f per query\n", nqc, kp, S,
                (double)h2[0] / nqc, (double)h2[1] / nqc);
      }
      float* dq = d_dev + (size_t)qb * k + koff;
      int32_t* iq = i_dev + (size_t)qb * k + koff;
      RYL_TRY(merge_lists(part.as<uint64_t>(), nullptr, nullptr, S, nqc, kp, dq, iq, id_add, s, k));
      if (koff + kp < k)
        RYL_LAUNCH(lower_bound_kernel, (nqc + 255) / 256, 256, 0, s, dq, iq, nqc, k, kp - 1, id_add, lb.as<uint64_t>());
    }
  }
  RYL_LAUNCH(poison_kernel, 64, 256, 0, s, bad, d_dev, i_dev, (size_t)nq * k);
  return RAYUELA_OK;
}

static const char* kBadLutMsg =
    "index_search: a lookup-table entry is not finite (or exceeds 1e37): queries / codebooks contain inf or NaN, or "
    "their products overflow";

// multi-device parent: every slot scans its shard for ALL queries (LUTs replicated); the per-shard top-k lists land in the
// first slot's gather buffer -- stored there directly by the shard's last kernel over NVLink peer access (or copied
// peer-to-peer when that is unavailable) -- and are merged there by the (dist, id) total order: the single exchange
// step of SURVEY 8e, fused into the search inside one process instead of an NCCL all-gather between processes.
static int index_search_multi(rayuela_index* ix, const float* queries, const float* codebooks, int nq, int d, int k,
                              float* dists, int32_t* idx, bool fast_lut) {
  const int D = (int)ix->shards.size();
  const int len = ix->kind == RAYUELA_SCAN_PQ ? d / ix->m : d;
  const size_t per = (size_t)nq * k;
  for (int i = 0; i < D; i++)
    RYL_ARG((int64_t)k <= ix->shards[i]->n, "index_search: k exceeds the size of a base shard");
  int cur = 0;
  RYL_CUDA(cudaGetDevice(&cur));
  const DeviceSlot root = ix->slots[0];
  RYL_CUDA(cudaSetDevice(root.device));
  auto body = [&]() -> int {
    DevBuf gd, gi, bad_all;
    RYL_TRY(gd.alloc((size_t)D * per * sizeof(float), root.stream));
    RYL_TRY(gi.alloc((size_t)D * per * sizeof(int32_t), root.stream));
    RYL_CUDA(cudaStreamSynchronize(root.stream));            // the gather buffers exist before any peer writes to them
    std::vector<int> bad(D, 0);
    RYL_TRY(for_each_slot(ix->slots, [&](int i) -> int {
      cudaStream_t s = ix->slots[i].stream;
      InArg<float> q_in, cb_in;
      RYL_TRY(q_in.bind(queries, (size_t)nq * d, false, s));
      RYL_TRY(cb_in.bind(codebooks, (size_t)ix->m * ix->h * len, false, s));
      DevBuf dl, il, bd;
      RYL_TRY(bd.alloc(sizeof(int), s));
      RYL_CUDA(cudaMemsetAsync(bd.p, 0, sizeof(int), s));
      float* dst_d = gd.as<float>() + (size_t)i * per;
      int32_t* dst_i = gi.as<int32_t>() + (size_t)i * per;
      if (ix->slots[i].direct_to_first) {
        // the shard's final merge kernel stores its (dist, id) lists over NVLink straight into slot 0's gather buffer:
        // search and exchange are one step, there is no separate copy
        RYL_TRY(index_search_dev(ix->shards[i], q_in.d, cb_in.d, nq, d, k, dst_d, dst_i, bd.as<int>(), s, fast_lut));
      } else {
        RYL_TRY(dl.alloc(per * sizeof(float), s));
        RYL_TRY(il.alloc(per * sizeof(int32_t), s));
        RYL_TRY(index_search_dev(ix->shards[i], q_in.d, cb_in.d, nq, d, k, dl.as<float>(), il.as<int32_t>(), bd.as<int>(), s,
                                 fast_lut));
        RYL_CUDA(cudaMemcpyPeerAsync(dst_d, root.device, dl.p, ix->slots[i].device, per * sizeof(float), s));
        RYL_CUDA(cudaMemcpyPeerAsync(dst_i, root.device, il.p, ix->slots[i].device, per * sizeof(int32_t), s));
      }
      RYL_CUDA(cudaMemcpyAsync(&bad[i], bd.p, sizeof(int), cudaMemcpyDeviceToHost, s));
      RYL_CUDA(cudaStreamSynchronize(s));
      return RAYUELA_OK;
    }));
    for (int i = 0; i < D; i++)
      if (bad[i]) return fail(RAYUELA_ERR_ARG, kBadLutMsg);
    OutArg<float> d_out;
    OutArg<int32_t> i_out;
    RYL_TRY(d_out.bind(dists, per, false, root.stream));
    RYL_TRY(i_out.bind(idx, per, false, root.stream));
    RYL_TRY(merge_lists(nullptr, gd.as<float>(), gi.as<int32_t>(), D, nq, k, d_out.d, i_out.d, 0, root.stream));
    RYL_TRY(d_out.flush(root.stream));
    RYL_TRY(i_out.flush(root.stream));
    RYL_CUDA(cudaStreamSynchronize(root.stream));
    return RAYUELA_OK;
  };
  const int rc = body();
  cudaSetDevice(cur);
  return rc;
}

extern "C" int rayuela_index_search(rayuela_index* ix, const float* queries, const float* codebooks, int nq, int d,
                                    int k, float* dists, int32_t* idx, unsigned flags, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  RYL_ARG(ix != nullptr, "index_search: null index");
  RYL_ARG(nq >= 1 && d >= 1, "index_search: nq and d must be positive");
  RYL_ARG(k >= 1 && (int64_t)k <= ix->n, "index_search: k must be in 1..n");
  const int m = ix->m, mh = m * ix->h;
  const bool pq = ix->kind == RAYUELA_SCAN_PQ;
  RYL_ARG(!pq || d % m == 0, "index_search: PQ scan needs d divisible by m");
  const int len = pq ? d / m : d;
  const bool dev = flags & RAYUELA_DEVICE_PTRS;
  const char* fast_env = getenv("RAYUELA_B200_FAST_LUT");
  const bool fast_lut = (flags & RAYUELA_FAST_LUT) || (fast_env && atoi(fast_env) != 0);
  if (!ix->shards.empty()) {
    RYL_ARG(!dev, "index_search: a multi-device index takes host arrays");
    return index_search_multi(ix, queries, codebooks, nq, d, k, dists, idx, fast_lut);
  }

  InArg<float> q_in, cb_in;
  RYL_TRY(q_in.bind(queries, (size_t)nq * d, dev, s));
  RYL_TRY(cb_in.bind(codebooks, (size_t)mh * len, dev, s));
  OutArg<float> d_out;
  OutArg<int32_t> i_out;
  RYL_TRY(d_out.bind(dists, (size_t)nq * k, dev, s));
  RYL_TRY(i_out.bind(idx, (size_t)nq * k, dev, s));
  DevBuf bad;
  RYL_TRY(bad.alloc(sizeof(int), s));
  RYL_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), s));
  RYL_TRY(index_search_dev(ix, q_in.d, cb_in.d, nq, d, k, d_out.d, i_out.d, bad.as<int>(), s, fast_lut));
  RYL_TRY(d_out.flush(s));
  RYL_TRY(i_out.flush(s));
  if (!dev) {
    int bad_h = 0;
    RYL_CUDA(cudaMemcpyAsync(&bad_h, bad.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    RYL_CUDA(cudaStreamSynchronize(s));
    if (bad_h) return fail(RAYUELA_ERR_ARG, kBadLutMsg);
  }
  return RAYUELA_OK;
}

extern "C" int rayuela_topk_merge(const float* dists_in, const int32_t* idx_in, int S, int nq, int k,
                                  float* dists_out, int32_t* idx_out, unsigned flags, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  RYL_ARG(S >= 1 && nq >= 1 && k >= 1, "topk_merge: S, nq, k must be positive");
  const bool dev = flags & RAYUELA_DEVICE_PTRS;
  InArg<float> din;
  InArg<int32_t> iin;
  RYL_TRY(din.bind(dists_in, (size_t)S * nq * k, dev, s));
  RYL_TRY(iin.bind(idx_in, (size_t)S * nq * k, dev, s));
  OutArg<float> dout;
  OutArg<int32_t> iout;
  RYL_TRY(dout.bind(dists_out, (size_t)nq * k, dev, s));
  RYL_TRY(iout.bind(idx_out, (size_t)nq * k, dev, s));
  RYL_TRY(merge_lists(nullptr, din.d, iin.d, S, nq, k, dout.d, iout.d, 0, s));
  RYL_TRY(dout.flush(s));
  RYL_TRY(iout.flush(s));
  if (!dev) RYL_CUDA(cudaStreamSynchronize(s));
  return RAYUELA_OK;
}

// ---- exact-signature compat symbols (host pointers, synchronous, abort on failure like UB would) -------
static void compat_scan(int kind, float* dists, int32_t* idx, const uint8_t* codes, const float* queries,
                        const float* codebooks, const float* dbnorms, int nq, int64_t n, int m, int h, int d, int k) {
  rayuela_index* ix = nullptr;
  int rc = rayuela_index_create(&ix, kind, codes, dbnorms, n, m, h, 0, 0, nullptr);
  if (rc == RAYUELA_OK) rc = rayuela_index_search(ix, queries, codebooks, nq, d, k, dists, idx, 0, nullptr);
  rayuela_index_free(ix);
  if (rc != RAYUELA_OK) {
    fprintf(stderr, "librayuela_b200: linscan failed (%d): %s\n", rc, rayuela_last_error());
    abort();
  }
}

extern "C" void linscan_aqd_query(float* dists, unsigned int* res, unsigned char* codes, float* centers,
                                  float* queries, int N, unsigned int NQ, int B, int K, int dim1codes,
                                  int dim1queries, int subdim) {
  // m is derived as B/8 (linscan_aqd.cpp:40); dim1codes is the code stride (== m at src/Linscan.jl:22-23)
  int m = B / 8;
  if (dim1codes != m || dim1queries != m * subdim) {
    fprintf(stderr, "librayuela_b200: linscan_aqd_query needs dim1codes == B/8 and dim1queries == (B/8)*subdim\n");
    abort();
  }
  compat_scan(RAYUELA_SCAN_PQ, dists, reinterpret_cast<int32_t*>(res), codes, queries, centers, nullptr, (int)NQ, N,
              m, 256, dim1queries, K);
}

extern "C" void linscan_aqd_query_extra_byte(float* dists, int* idx, unsigned char* codes, float* queries,
                                             float* codebooks, float* dbnorms, int nqueries, int ncodes, int m,
                                             int h, int d, int nn) {
  compat_scan(RAYUELA_SCAN_LSQ, dists, idx, codes, queries, codebooks, dbnorms, nqueries, ncodes, m, h, d, nn);
}

extern "C" void linscan_aqd_cq_query_extra_byte(float* dists, int* idx, unsigned char* codes, float* queries,
                                                float* codebooks, int nqueries, int ncodes, int m, int h, int d,
                                                int nn) {
  compat_scan(RAYUELA_SCAN_CQ, dists, idx, codes, queries, codebooks, nullptr, nqueries, ncodes, m, h, d, nn);
}
