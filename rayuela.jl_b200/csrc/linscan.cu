// linscan.cu -- path (2): asymmetric-distance linear scan on B200.
//   K4 lut_kernel     per-query m*256 lookup table, exact reference arithmetic order
//   K5 scan_kernel    byte-code scan from a shared-memory LUT tile + streaming top-k (threshold filter,
//                     candidate buffer, block bitonic compaction)
//   K6 merge_kernel   k-way merge of sorted (dist,id) lists (DB slices of one GPU, or per-GPU shards)
// Replaces deps/src/linscan_aqd.cpp:37-102 and deps/src/linscan_aqd_pairwise_byte.cpp:14-176.
#include <algorithm>
#include <cstring>

#include "common.cuh"

namespace ryl {

static constexpr int kH = 256;
static constexpr int kScanThreads = 256;
static constexpr int kCodesPerThread = 4;
static constexpr int kRound = kScanThreads * kCodesPerThread;  // codes per block round

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
                                                  int tiled) {
  constexpr int CH = 64;
  __shared__ float cs[32][CH + 1];
  __shared__ float qs[32][CH];
  const int e0 = blockIdx.x * 32, q0 = blockIdx.y * 32;
  const int e = threadIdx.x & 31, qg = threadIdx.x >> 5;
  const int qoff = (KIND == RAYUELA_SCAN_PQ) ? (e0 / kH) * len : 0;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int base = 0; base < len; base += CH) {
    const int chunk = min(CH, len - base);
    for (int i = threadIdx.x; i < 32 * CH; i += 256) {
      int r = i / CH, t = i % CH;
      if (t < chunk) {
        cs[r][t] = cb[(size_t)(e0 + r) * len + base + t];
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
    if (q < nq) {
      if (tiled) {
        // layout of scan8_kernel's shared-memory tile (see there): [q/16][(q%16)/4][c][((q%4)/2)*8 + k][q%2]
        const int ent = e0 + e, k = ent >> 8, c = ent & 255;
        lut[(size_t)(q >> 4) * 32768 + ((q & 15) >> 2) * 8192 + c * 32 + ((((q & 3) >> 1) * 8 + k) << 1) + (q & 1)] =
            acc[i];
      } else {
        lut[(size_t)q * mh + e0 + e] = acc[i];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// Index re-layout: codes m-by-n (vector-major, m bytes each) -> padded to MP = 8 or 16 bytes per vector so
// the scan reads each code with one aligned 8/16-byte load.
// ------------------------------------------------------------------------------------------------------
__global__ void pad_codes_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int64_t n, int m, int mp) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = n * mp;
  for (; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t v = i / mp;
    int k = (int)(i % mp);
    out[i] = k < m ? in[v * m + k] : (uint8_t)0;
  }
}

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
      __syncthreads();
    }
  }
}

__device__ __forceinline__ int pow2ceil(int x) {
  int p = 2;
  while (p < x) p <<= 1;
  return p;
}

struct ScanParams {
  const uint8_t* codes;  // [n][MP]
  const float* norms;    // [n] or nullptr
  const float* lut;      // [nq][m*256]
  uint64_t* cand;        // [slices][nqtiles*QT][cap]
  uint64_t* part;        // [slices][nq][k] sorted keys (low word = local id)
  const uint64_t* lb;    // [nq] or nullptr: only keys strictly greater than lb[q] qualify (k > one pass)
  int64_t n, slice_len;
  int nq, k, cap;
};

// ------------------------------------------------------------------------------------------------------
// K5: scan.  grid = (query tiles, DB slices).  The block keeps the LUTs of QT queries in shared memory,
// streams its slice of codes, and for every (code, query) sums the M lookups in ascending-k order starting
// from +0 (pairwise_byte.cpp:70-73), adds the norm last (:74), and keeps (dist, id) if dist <= tau_q, the
// k-th best distance known so far for that query.  Candidates go to a per-query buffer (global, L2-resident);
// when a buffer could overflow in the next round the block sorts it (bitonic, shared memory), keeps the k
// best and tightens tau_q.  The final buffers are sorted and written as keys.
// ------------------------------------------------------------------------------------------------------
template <int M, int QT, bool NORMS>
__global__ void __launch_bounds__(kScanThreads) scan_kernel(ScanParams p) {
  constexpr int MP = (M <= 8) ? 8 : 16;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* lut_s = reinterpret_cast<float*>(smem_raw);                       // [QT][M][256]
  uint64_t* sortbuf = reinterpret_cast<uint64_t*>(lut_s + QT * M * kH);    // [cap]
  __shared__ int cnt_s[QT];
  __shared__ float tau_s[QT];
  __shared__ uint64_t lb_s[QT];

  const int q0 = blockIdx.x * QT;
  const int slice = blockIdx.y;
  if (threadIdx.x < QT) lb_s[threadIdx.x] = p.lb ? p.lb[min(q0 + (int)threadIdx.x, p.nq - 1)] : 0ull;
  const int64_t begin = (int64_t)slice * p.slice_len;
  const int64_t end = min(p.n, begin + p.slice_len);
  uint64_t* cand = p.cand + ((size_t)slice * gridDim.x * QT + (size_t)blockIdx.x * QT) * p.cap;

  for (int i = threadIdx.x; i < QT * M * kH / 4; i += kScanThreads) {
    int q = (i * 4) / (M * kH);
    int r = (i * 4) % (M * kH);
    int qq = min(q0 + q, p.nq - 1);
    reinterpret_cast<float4*>(lut_s)[i] = *reinterpret_cast<const float4*>(p.lut + (size_t)qq * M * kH + r);
  }
  if (threadIdx.x < QT) {
    cnt_s[threadIdx.x] = 0;
    tau_s[threadIdx.x] = __int_as_float(0x7f800000);
  }
  __syncthreads();

  float tau[QT];
#pragma unroll
  for (int q = 0; q < QT; q++) tau[q] = __int_as_float(0x7f800000);

  auto compact = [&](int q) {
    const int c = cnt_s[q];
    const int np2 = pow2ceil(c);
    uint64_t* cq = cand + (size_t)q * p.cap;
    for (int t = threadIdx.x; t < np2; t += kScanThreads) sortbuf[t] = t < c ? cq[t] : ~0ull;
    __syncthreads();
    block_bitonic_sort(sortbuf, np2);
    const int keep = min(c, p.k);
    for (int t = threadIdx.x; t < keep; t += kScanThreads) cq[t] = sortbuf[t];
    if (threadIdx.x == 0) {
      cnt_s[q] = keep;
      if (c >= p.k) tau_s[q] = ordered_to_f32((uint32_t)(sortbuf[p.k - 1] >> 32));
    }
    __syncthreads();
  };

  for (int64_t base = begin; base < end; base += kRound) {
#pragma unroll
    for (int cc = 0; cc < kCodesPerThread; cc++) {
      const int64_t i = base + cc * kScanThreads + threadIdx.x;
      if (i < end) {
        uint32_t w[MP / 4];
        if (MP == 8) {
          uint2 v = *reinterpret_cast<const uint2*>(p.codes + i * MP);
          w[0] = v.x;
          w[1] = v.y;
        } else {
          uint4 v = *reinterpret_cast<const uint4*>(p.codes + i * MP);
          w[0] = v.x;
          w[1] = v.y;
          w[MP / 4 - 2] = v.z;
          w[MP / 4 - 1] = v.w;
        }
        float nrm = 0.f;
        if (NORMS) nrm = p.norms[i];
#pragma unroll
        for (int q = 0; q < QT; q++) {
          float s = 0.0f;
#pragma unroll
          for (int k = 0; k < M; k++) {
            uint32_t b = (w[k >> 2] >> (8 * (k & 3))) & 0xFFu;
            s = __fadd_rn(s, lut_s[(q * M + k) * kH + b]);
          }
          if (NORMS) s = __fadd_rn(s, nrm);
          if (s <= tau[q]) {
            const uint64_t key = make_key(s, (uint32_t)i);
            if (!p.lb || key > lb_s[q]) {
              int pos = atomicAdd(&cnt_s[q], 1);
              cand[(size_t)q * p.cap + pos] = key;
            }
          }
        }
      }
    }
    __syncthreads();
    bool any = false;
    for (int q = 0; q < QT; q++) {
      if (cnt_s[q] > p.cap - kRound) {  // block-uniform
        compact(q);
        any = true;
      }
    }
    if (any) {
#pragma unroll
      for (int q = 0; q < QT; q++) tau[q] = tau_s[q];
    }
  }

  for (int q = 0; q < QT; q++) {
    if (q0 + q >= p.nq) break;
    compact(q);
    const int c = cnt_s[q];
    uint64_t* out = p.part + ((size_t)slice * p.nq + q0 + q) * p.k;
    for (int t = threadIdx.x; t < p.k; t += kScanThreads) out[t] = t < c ? sortbuf[t] : ~0ull;
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------------
// K5 (m <= 8): bank-conflict-free scan.
//
// A 4-byte LUT lookup per byte of code makes the scan shared-memory-gather bound (32 lookups/clk/SM), and
// 32 lanes looking up random entries of the same 256-entry row collide ~3.6x (measured, profiles/r1_v1).
// Here every lane of a half-warp is at a DIFFERENT codebook k at any instant, and the LUT tile is laid out
// so that codebook k of query-pair g lives in bank-pair g*8 + k:
//      tile[tt][c][bp = g*8 + k][e]   (float; 4 queries per 32 KB tile: q = tt*4 + g*2 + e)
// lane = hw*16 + g*8 + j handles code stream p = hw*8 + j of its warp's chunk and the query pair g of every
// tile; at step s it is at codebook k = (s - j - 1) mod 8, so the 16 lanes of a half-warp hit 16 distinct
// bank-pairs with one LDS.64 each -> no conflicts, 2 queries per load.  The sum for one code must still be
// ((0 + t_0) + t_1) + ... in ascending k (pairwise_byte.cpp:70-73), so a lane's code simply starts j+1 steps
// "late": the index stores each lane's byte stream pre-skewed (skew_codes_kernel) and the lane reads one
// aligned 8-byte word per 8 steps.  Accumulate / restart / capture are done with packed FFMA2
// (fma.rn.f32x2, exact per element):  acc = acc*keep_s + v   (keep_s = 0 at the step where the lane's next
// code starts), done += acc*cap_s (cap_s = 1 at the step where its code completes).
// The 128 KB LUT tile is staged with bulk async copies (cp.async.bulk + mbarrier).
// ------------------------------------------------------------------------------------------------------
static constexpr int kChunkL = 64;                    // codes per lane stream per chunk
static constexpr int kChunkCodes = 16 * kChunkL;      // 1024 codes per warp chunk
static constexpr int kChunkBlocks = kChunkL + 1;      // 8-step blocks per chunk (one extra for the skew tail)
static constexpr int kScan8Warps = 16;
static constexpr int kScan8RBMax = 16;                // longest round (blocks per warp between overflow checks)
static constexpr int kScan8SortKeys = 8192;           // shared-memory sort buffer (64 KB)
static constexpr int kLutTileBytes = 131072;          // 16 queries * 8 codebooks * 256 * 4 B

// W[chunk][t][p] (uint64): bytes of stream p = codes chunk*1024 + 16u + p (u = 0..63), delayed by (p&7)+1 bytes.
__global__ void skew_codes_kernel(const uint8_t* __restrict__ codes, uint64_t* __restrict__ W, int64_t n, int m,
                                  int64_t nchunks) {
  const int64_t total = nchunks * kChunkBlocks * 16;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int p = (int)(i & 15);
    const int t = (int)((i >> 4) % kChunkBlocks);
    const int64_t chunk = (i >> 4) / kChunkBlocks;
    const int delay = (p & 7) + 1;
    uint64_t w = 0;
    for (int b = 0; b < 8; b++) {
      const int sb = 8 * t + b - delay;              // byte index in the lane's undelayed stream
      if (sb < 0 || sb >= 8 * kChunkL) continue;
      const int u = sb >> 3, k = sb & 7;
      const int64_t id = chunk * kChunkCodes + 16 * u + p;
      if (id < n && k < m) w |= (uint64_t)codes[id * m + k] << (8 * b);
    }
    W[i] = w;
  }
}

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
template <int IMM>
__device__ __forceinline__ uint64_t lds64(uint32_t addr) {
  uint64_t v;
  asm("ld.shared.b64 %0, [%1+%2];" : "=l"(v) : "r"(addr), "n"(IMM));
  return v;
}

// Block-wide radix select over c distinct 64-bit keys in shared memory: returns the k-th smallest key
// (1 <= k <= c).  Eight 8-bit passes from the most significant byte; per pass a 256-bin histogram of the keys
// that match the prefix decided so far (lanes with equal digits are aggregated with match.any so the all-equal
// leading bytes of similar distances cost one shared atomic per warp), then warp 0 picks the bin holding rank k.
__device__ __forceinline__ uint64_t block_radix_select(const uint64_t* buf, int c, int k, int* hist, int* sel) {
  const int tid = threadIdx.x, lane = tid & 31, nt = blockDim.x;
  // Leading bytes shared by all keys (same exponent / high mantissa for similar distances) need no pass:
  // OR of (key ^ buf[0]) over the block tells the first byte where any two keys differ.
  uint64_t diff = 0;
  const uint64_t k0 = buf[0];
  for (int t = tid; t < c; t += nt) diff |= buf[t] ^ k0;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) diff |= __shfl_xor_sync(0xffffffffu, diff, off);
  if (tid == 0) {
    sel[2] = 0;
    sel[3] = 0;
  }
  __syncthreads();
  if (lane == 0 && diff) {
    atomicOr(&sel[2], (int)(uint32_t)diff);
    atomicOr(&sel[3], (int)(uint32_t)(diff >> 32));
  }
  __syncthreads();
  diff = ((uint64_t)(uint32_t)sel[3] << 32) | (uint32_t)sel[2];
  const int first = diff ? (__clzll((long long)diff) >> 3) : 8;       // first pass that can discriminate
  uint64_t prefix = first ? (k0 >> (64 - 8 * first)) : 0;
  int krem = k;
  for (int pass = first; pass < 8; pass++) {
    const int shift = 56 - 8 * pass;
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    if (pass == first) {
      // the first discriminating byte usually takes few distinct values: aggregate equal digits per warp
      for (int t0 = 0; t0 < c; t0 += nt) {        // uniform trip count: match.any needs converged warps
        const int t = t0 + tid;
        const bool act = t < c;
        const uint32_t digit = act ? (uint32_t)(buf[t] >> shift) & 255u : 256u + lane;
        const uint32_t peers = __match_any_sync(0xffffffffu, digit);
        if (act && lane == __ffs(peers) - 1) atomicAdd(&hist[digit], __popc(peers));
      }
    } else {
      for (int t = tid; t < c; t += nt) {
        const uint64_t key = buf[t];
        if ((key >> (shift + 8)) == prefix) atomicAdd(&hist[(uint32_t)(key >> shift) & 255u], 1);
      }
    }
    __syncthreads();
    if (tid < 32) {
      int loc[8], sum = 0;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        loc[i] = hist[lane * 8 + i];
        sum += loc[i];
      }
      int inc = sum;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += v;
      }
      int run = inc - sum;
      if (run < krem && krem <= inc) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
          if (run < krem && krem <= run + loc[i]) {
            sel[0] = lane * 8 + i;
            sel[1] = krem - run;
          }
          run += loc[i];
        }
      }
    }
    __syncthreads();
    prefix = (prefix << 8) | (uint32_t)sel[0];
    krem = sel[1];
  }
  return prefix;
}

struct Scan8Params {
  const uint64_t* W;     // skewed codes [nchunks][65][16]
  const float* norms;    // [n] or nullptr
  const float* lut;      // tiled [qtiles][32768]
  uint64_t* cand;        // [slices][qtiles*16][cap]
  uint64_t* part;        // [slices][nq][k]
  const uint64_t* lb;    // [nq] or nullptr: only keys strictly greater than lb[q] qualify (k > one pass)
  int64_t n, nchunks, chunks_per_slice;
  int nq, k, cap, soft, rbmax;   // soft: compact a query's buffer once it holds more than this many keys
};

template <bool NORMS>
__global__ void __launch_bounds__(kScan8Warps * 32, 1) scan8_kernel(Scan8Params p) {
  constexpr int NT = kScan8Warps * 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int cnt_s[16];
  __shared__ float tau_s[16];
  __shared__ __align__(8) uint64_t mbar;

  const uint32_t lut_addr = smem_u32(smem_raw);
  uint64_t* sortbuf = reinterpret_cast<uint64_t*>(smem_raw + kLutTileBytes);

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int j = lane & 7, g = (lane >> 3) & 1, pidx = (lane >> 4) * 8 + j;
  const int q0 = blockIdx.x * 16;
  const int slice = blockIdx.y;
  uint64_t* cand = p.cand + ((size_t)slice * gridDim.x * 16 + (size_t)blockIdx.x * 16) * p.cap;

  // ---- stage the LUT tile: 4 bulk async copies of 32 KB, completion on one mbarrier -----------------------
  const uint32_t mbar_addr = smem_u32(&mbar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_addr));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __shared__ uint64_t lb_s[16];
  if (tid < 16) {
    cnt_s[tid] = 0;
    tau_s[tid] = __int_as_float(0x7f800000);
    lb_s[tid] = p.lb ? p.lb[min(q0 + tid, p.nq - 1)] : 0ull;
  }
  __syncthreads();
  if (tid == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_addr), "r"(kLutTileBytes)
                 : "memory");
    const char* src = reinterpret_cast<const char*>(p.lut) + (size_t)blockIdx.x * kLutTileBytes;
#pragma unroll
    for (int i = 0; i < 4; i++)
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

  // ---- per-lane constants -----------------------------------------------------------------------------------
  uint32_t off[8];
  uint64_t keep2[8], cap2[8];
#pragma unroll
  for (int s = 0; s < 8; s++) {
    off[s] = lut_addr + ((g * 8 + ((s - j - 1) & 7)) << 3);
    const float kp = (s == ((j + 1) & 7)) ? 0.f : 1.f;
    const float cp = (s == j) ? 1.f : 0.f;
    keep2[s] = pack2(kp, kp);
    cap2[s] = pack2(cp, cp);
  }
  float tau[8];
#pragma unroll
  for (int i = 0; i < 8; i++) tau[i] = __int_as_float(0x7f800000);
  uint64_t acc[4] = {0, 0, 0, 0}, done[4] = {0, 0, 0, 0};

  __shared__ int hist_s[256];
  __shared__ int sel_s[4];
  // Intermediate compaction: keep the k smallest keys (unordered) and set tau to the k-th distance.
  auto compact = [&](int q) {
    const int c = cnt_s[q];
    uint64_t* cq = cand + (size_t)q * p.cap;
    if (c <= p.k) {            // nothing to drop; tau only if the buffer holds exactly k keys
      if (c == p.k) {          // tau = the largest of the k keys
        uint64_t mx = 0;
        for (int t = tid; t < c; t += NT) mx = max(mx, cq[t]);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, off));
        if (lane == 0) sortbuf[w] = mx;
        __syncthreads();
        if (tid == 0) {
          for (int i = 1; i < kScan8Warps; i++) mx = max(mx, sortbuf[i]);
          tau_s[q] = ordered_to_f32((uint32_t)(mx >> 32));
        }
        __syncthreads();
      }
      return;
    }
    if (c <= 512) {            // small buffers: a bitonic sort is cheaper than eight radix passes
      const int np2 = pow2ceil(c);
      for (int t = tid; t < np2; t += NT) sortbuf[t] = t < c ? cq[t] : ~0ull;
      __syncthreads();
      block_bitonic_sort(sortbuf, np2);
      for (int t = tid; t < p.k; t += NT) cq[t] = sortbuf[t];
      if (tid == 0) {
        cnt_s[q] = p.k;
        tau_s[q] = ordered_to_f32((uint32_t)(sortbuf[p.k - 1] >> 32));
      }
      __syncthreads();
      return;
    }
    for (int t = tid; t < c; t += NT) sortbuf[t] = cq[t];
    __syncthreads();
    const uint64_t pivot = block_radix_select(sortbuf, c, p.k, hist_s, sel_s);
    if (tid == 0) sel_s[2] = 0;
    __syncthreads();
    for (int t = tid; t < c; t += NT) {
      const uint64_t key = sortbuf[t];
      if (key <= pivot) cq[atomicAdd(&sel_s[2], 1)] = key;     // exactly k keys (keys are distinct)
    }
    if (tid == 0) {
      cnt_s[q] = p.k;
      tau_s[q] = ordered_to_f32((uint32_t)(pivot >> 32));
    }
    __syncthreads();
  };
  // Final: the k smallest, sorted, left in sortbuf[0 .. min(c,k)).
  auto finalize = [&](int q) {
    compact(q);
    const int c = cnt_s[q];
    const int np2 = pow2ceil(c);
    uint64_t* cq = cand + (size_t)q * p.cap;
    for (int t = tid; t < np2; t += NT) sortbuf[t] = t < c ? cq[t] : ~0ull;
    __syncthreads();
    block_bitonic_sort(sortbuf, np2);
  };

  // ---- this block's chunk range; warp w takes chunks c0 + w, c0 + w + 16, ... ---------------------------------
  const int64_t c0 = (int64_t)slice * p.chunks_per_slice;
  const int64_t c1 = min(p.nchunks, c0 + p.chunks_per_slice);
  const int nci = (int)((c1 - c0 + kScan8Warps - 1) / kScan8Warps);

  const uint32_t n32 = (uint32_t)p.n;
  const int hard = p.cap - kScan8Warps * 16 * p.rbmax;   // a round adds at most 256 keys per block-step
  int sched = 1;                                          // warm-up: rounds of 1,1,2,4,... blocks so tau
  bool first = true;                                      // stops being +inf as early as possible

  for (int ci = 0; ci < nci; ci++) {
    const int64_t chunk = c0 + (int64_t)ci * kScan8Warps + w;
    const bool active = chunk < c1;                                    // warp-uniform
    const uint64_t* wp = p.W + (active ? chunk : c0) * (kChunkBlocks * 16) + pidx;
    const float* np = p.norms + chunk * kChunkCodes + pidx;            // norm of the code completed in block t+1
    uint32_t id = (uint32_t)(chunk * kChunkCodes) + pidx - 16;         // code completed in block t (t >= 1)
    uint64_t W0 = 0, W1 = 0;
    if (active) {
      W0 = __ldg(wp);
      W1 = __ldg(wp + 16);
    }
    wp += 32;
    float nrm0 = 0.f;
    int t = 0;
    while (t < kChunkBlocks) {
      const int len = min(sched, kChunkBlocks - t);
      for (int b = 0; b < len; b++, t++) {
        if (active) {
          uint64_t W2 = 0;
          if (t + 2 < kChunkBlocks) W2 = __ldg(wp);
          float nrm1 = 0.f;
          if (NORMS && t + 1 < kChunkBlocks && id + 16 < n32) nrm1 = __ldg(np);
          const uint32_t lo = (uint32_t)W0, hi = (uint32_t)(W0 >> 32);
#define RYL_STEP(S, WREG, SHL, SHR)                                                      \
  {                                                                                      \
    const uint32_t a = ((SHL ? (WREG << 7) : (WREG >> SHR)) & 0x7F80u) + off[S];         \
    uint64_t v0 = lds64<0>(a), v1 = lds64<32768>(a), v2 = lds64<65536>(a), v3 = lds64<98304>(a); \
    acc[0] = ffma2(acc[0], keep2[S], v0);                                                \
    acc[1] = ffma2(acc[1], keep2[S], v1);                                                \
    acc[2] = ffma2(acc[2], keep2[S], v2);                                                \
    acc[3] = ffma2(acc[3], keep2[S], v3);                                                \
    done[0] = ffma2(acc[0], cap2[S], done[0]);                                           \
    done[1] = ffma2(acc[1], cap2[S], done[1]);                                           \
    done[2] = ffma2(acc[2], cap2[S], done[2]);                                           \
    done[3] = ffma2(acc[3], cap2[S], done[3]);                                           \
  }
          RYL_STEP(0, lo, 1, 0)
          RYL_STEP(1, lo, 0, 1)
          RYL_STEP(2, lo, 0, 9)
          RYL_STEP(3, lo, 0, 17)
          RYL_STEP(4, hi, 1, 0)
          RYL_STEP(5, hi, 0, 1)
          RYL_STEP(6, hi, 0, 9)
          RYL_STEP(7, hi, 0, 17)
#undef RYL_STEP
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
                  const int q = (i >> 1) * 4 + g * 2 + (i & 1);
                  const uint64_t key = make_key(dv[i], id);
                  if (!p.lb || key > lb_s[q]) {
                    int pos = atomicAdd(&cnt_s[q], 1);
                    cand[(size_t)q * p.cap + pos] = key;
                  }
                }
              }
            }
          }
#pragma unroll
          for (int tt = 0; tt < 4; tt++) done[tt] = 0ull;
          W0 = W1;
          W1 = W2;
          nrm0 = nrm1;
          wp += 16;
          np += 16;
          id += 16;
        }
      }
      sched = min(p.rbmax, first ? 1 : sched * 2);
      first = false;
      // a buffer is compacted when it exceeds the soft limit, could overflow in the next round, or holds its
      // first k candidates (tau still +inf)
      bool mine = false;
      if (tid < 16) {
        const int c = cnt_s[tid];
        mine = c > p.soft || c > hard || (c >= p.k && tau_s[tid] == __int_as_float(0x7f800000));
      }
      if (__syncthreads_or(mine)) {
        for (int q = 0; q < 16; q++) {
          const int c = cnt_s[q];
          if (c > p.soft || c > hard || (c >= p.k && tau_s[q] == __int_as_float(0x7f800000))) compact(q);
        }
#pragma unroll
        for (int tt = 0; tt < 4; tt++) {
          tau[2 * tt] = tau_s[tt * 4 + g * 2];
          tau[2 * tt + 1] = tau_s[tt * 4 + g * 2 + 1];
        }
      }
    }
  }

  for (int q = 0; q < 16; q++) {
    if (q0 + q >= p.nq) break;
    finalize(q);
    const int c = cnt_s[q];
    uint64_t* out = p.part + ((size_t)slice * p.nq + q0 + q) * p.k;
    for (int i = tid; i < p.k; i += NT) out[i] = i < c ? sortbuf[i] : ~0ull;
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------------
// K6: merge of S sorted lists per query into the global top-k, by the (dist, id) total order.
//   keys != null : lists are 64-bit keys [S][nq][k] from scan_kernel (ids local; id_add makes them final)
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
  int kind = 0, m = 0, h = 0, mp = 0, device = 0;
  int64_t n = 0, id_offset = 0;
  int64_t nchunks = 0;   // m <= 8: skewed layout for scan8_kernel
  DevBuf codes, norms, skew;
};

template <int M, int QT>
static int launch_scan_mq(const ScanParams& p, bool norms, dim3 grid, size_t smem, cudaStream_t s) {
  if (norms) {
    RYL_CUDA(cudaFuncSetAttribute(scan_kernel<M, QT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RYL_LAUNCH((scan_kernel<M, QT, true>), grid, kScanThreads, smem, s, p);
  } else {
    RYL_CUDA(cudaFuncSetAttribute(scan_kernel<M, QT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RYL_LAUNCH((scan_kernel<M, QT, false>), grid, kScanThreads, smem, s, p);
  }
  return RAYUELA_OK;
}

static constexpr int scan_qt(int m) { return m <= 8 ? 16 : 8; }

static int launch_scan(int m, const ScanParams& p, bool norms, dim3 grid, size_t smem, cudaStream_t s) {
  switch (m) {
#define RYL_CASE(M) \
  case M:           \
    return launch_scan_mq<M, scan_qt(M)>(p, norms, grid, smem, s);
    RYL_CASE(1) RYL_CASE(2) RYL_CASE(3) RYL_CASE(4) RYL_CASE(5) RYL_CASE(6) RYL_CASE(7) RYL_CASE(8)
    RYL_CASE(9) RYL_CASE(10) RYL_CASE(11) RYL_CASE(12) RYL_CASE(13) RYL_CASE(14) RYL_CASE(15) RYL_CASE(16)
#undef RYL_CASE
  }
  return fail(RAYUELA_ERR_ARG, "linscan: m must be in 1..16");
}

static int host_pow2ceil(int x) {
  int p = 2;
  while (p < x) p <<= 1;
  return p;
}

extern "C" int rayuela_index_create(rayuela_index** out, int kind, const uint8_t* codes, const float* dbnorms,
                                    int64_t n, int m, int h, int64_t id_offset, unsigned flags, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  RYL_ARG(out != nullptr, "index_create: out is null");
  RYL_ARG(kind >= 0 && kind <= 2, "index_create: unknown kind");
  RYL_ARG(h == kH, "index_create: only h = 256 is supported (one byte per codebook)");
  RYL_ARG(m >= 1 && m <= 16, "index_create: m must be in 1..16");
  RYL_ARG(n >= 1 && n < (1ll << 32) - 4096, "index_create: n must be in 1..2^32-4097 per index shard");
  RYL_ARG(codes != nullptr, "index_create: codes is null");
  RYL_ARG(kind != RAYUELA_SCAN_LSQ || dbnorms != nullptr, "index_create: LSQ scan needs dbnorms");
  const bool dev = flags & RAYUELA_DEVICE_PTRS;
  rayuela_index* ix = new rayuela_index();
  ix->kind = kind;
  ix->m = m;
  ix->h = h;
  ix->mp = m <= 8 ? 8 : 16;
  ix->n = n;
  ix->id_offset = id_offset;
  cudaGetDevice(&ix->device);
  auto body = [&]() -> int {
    InArg<uint8_t> raw;
    RYL_TRY(raw.bind(codes, (size_t)n * m, dev, s));
    if (m <= 8) {   // skewed streams for the conflict-free scan
      ix->nchunks = (n + kChunkCodes - 1) / kChunkCodes;
      const int64_t words = ix->nchunks * kChunkBlocks * 16;
      RYL_TRY(ix->skew.alloc((size_t)words * sizeof(uint64_t), s));
      int blocks = (int)std::min<int64_t>((words + 255) / 256, 148 * 16);
      RYL_LAUNCH(skew_codes_kernel, blocks, 256, 0, s, raw.d, ix->skew.as<uint64_t>(), n, m, ix->nchunks);
    } else {
      RYL_TRY(ix->codes.alloc((size_t)n * ix->mp, s));
      int blocks = (int)std::min<int64_t>((n * ix->mp + 255) / 256, 148 * 16);
      RYL_LAUNCH(pad_codes_kernel, blocks, 256, 0, s, raw.d, ix->codes.as<uint8_t>(), n, m, ix->mp);
    }
    if (kind == RAYUELA_SCAN_LSQ) {
      RYL_TRY(ix->norms.alloc((size_t)n * sizeof(float), s));
      RYL_CUDA(cudaMemcpyAsync(ix->norms.p, dbnorms, (size_t)n * sizeof(float),
                               dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
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
    ix->codes.release();
    ix->norms.release();
    ix->skew.release();
    delete ix;
  }
  return RAYUELA_OK;
}

static int merge_lists(const uint64_t* keys, const float* din, const int32_t* iin, int S, int nq, int k, float* dout,
                       int32_t* iout, int64_t id_add, cudaStream_t s, int ldo = 0) {
  if (ldo == 0) ldo = k;
  size_t smem = (size_t)host_pow2ceil(S * k) * sizeof(uint64_t);
  RYL_ARG(smem <= 200 * 1024, "topk merge: S*k too large for a single pass (max 16384 keys... 25600)");
  RYL_CUDA(cudaFuncSetAttribute(merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  RYL_LAUNCH(merge_kernel, nq, 256, smem, s, keys, din, iin, S, nq, k, dout, iout, id_add, ldo);
  return RAYUELA_OK;
}

extern "C" int rayuela_index_search(rayuela_index* ix, const float* queries, const float* codebooks, int nq, int d,
                                    int k, float* dists, int32_t* idx, unsigned flags, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  RYL_ARG(ix != nullptr, "index_search: null index");
  RYL_ARG(nq >= 1 && d >= 1, "index_search: nq and d must be positive");
  RYL_ARG(k >= 1 && (int64_t)k <= ix->n, "index_search: k must be in 1..n");
  const int m = ix->m, mh = m * kH;
  const bool pq = ix->kind == RAYUELA_SCAN_PQ;
  RYL_ARG(!pq || d % m == 0, "index_search: PQ scan needs d divisible by m");
  const int len = pq ? d / m : d;
  const bool dev = flags & RAYUELA_DEVICE_PTRS;

  InArg<float> q_in, cb_in;
  RYL_TRY(q_in.bind(queries, (size_t)nq * d, dev, s));
  RYL_TRY(cb_in.bind(codebooks, (size_t)mh * len, dev, s));
  OutArg<float> d_out;
  OutArg<int32_t> i_out;
  RYL_TRY(d_out.bind(dists, (size_t)nq * k, dev, s));
  RYL_TRY(i_out.bind(idx, (size_t)nq * k, dev, s));

  const bool v2 = m <= 8;                                   // conflict-free scan8_kernel
  const int QT = v2 ? 16 : scan_qt(m);
  // One pass returns at most kmax results per query (shared-memory selection buffer).  Larger k -- the reference's
  // default is k = 10000 (src/Linscan.jl:10) -- takes ceil(k / kmax) passes: pass p keeps only keys strictly
  // greater than the last key of pass p-1, which is exact because (dist, id) keys are a total order.
  const int kmax = v2 ? 4096 : 3584;
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
    const size_t lut_floats = v2 ? (size_t)qtiles * (kLutTileBytes / 4) : (size_t)nqc * mh;
    RYL_TRY(lut.alloc(lut_floats * sizeof(float), s));
    if (v2 && (m < 8 || nqc % 16)) RYL_CUDA(cudaMemsetAsync(lut.p, 0, lut.bytes, s));   // zero rows for k >= m
    dim3 lg(mh / 32, (nqc + 31) / 32);
    const float* qptr = q_in.d + (size_t)qb * d;
    const int tiled = v2 ? 1 : 0;
    if (ix->kind == RAYUELA_SCAN_LSQ)
      RYL_LAUNCH(lut_kernel<RAYUELA_SCAN_LSQ>, lg, 256, 0, s, qptr, cb_in.d, lut.as<float>(), nqc, d, len, mh, tiled);
    else if (ix->kind == RAYUELA_SCAN_CQ)
      RYL_LAUNCH(lut_kernel<RAYUELA_SCAN_CQ>, lg, 256, 0, s, qptr, cb_in.d, lut.as<float>(), nqc, d, len, mh, tiled);
    else
      RYL_LAUNCH(lut_kernel<RAYUELA_SCAN_PQ>, lg, 256, 0, s, qptr, cb_in.d, lut.as<float>(), nqc, d, len, mh, tiled);
    if (k > kmax) RYL_TRY(lb.alloc((size_t)nqc * sizeof(uint64_t), s));

    for (int koff = 0; koff < k; koff += kmax) {
      const int kp = std::min(kmax, k - koff);
      const uint64_t* lbp = koff > 0 ? lb.as<uint64_t>() : nullptr;
      // v2: soft compaction limit 2k, hard limit = capacity minus what one round can add (256 codes per block-step)
      // soft compaction limit: up to 4k (fewer, relatively cheaper selections) while a full round still fits
      int soft = std::max(512, std::min(4 * kp, std::max(2 * kp, kScan8SortKeys - 256 * kScan8RBMax))), rbmax = kScan8RBMax;
      if (soft + 256 * kScan8RBMax > kScan8SortKeys) {
        soft = kp + (kScan8SortKeys - kp) / 2;
        rbmax = std::max(1, (kScan8SortKeys - soft) / 256);
      }
      const int cap = v2 ? soft + 256 * rbmax : host_pow2ceil(std::max(2 * kRound, 2 * kp + kRound));
      const size_t smem = v2 ? (size_t)kLutTileBytes + (size_t)kScan8SortKeys * sizeof(uint64_t)
                             : (size_t)QT * mh * sizeof(float) + (size_t)cap * sizeof(uint64_t);
      RYL_ARG(smem <= 227 * 1024, "index_search: shared-memory budget exceeded");
      // DB slices: enough blocks for several waves, slices no shorter than 8 rounds, S*k within one merge pass
      const int64_t unit = v2 ? (int64_t)kChunkCodes * kScan8Warps : kRound;   // codes per block round
      // whole waves of query tiles run unsliced; fewer tiles than SMs -> slice the base to fill one wave (two for
      // small k, where the per-slice warm-up is cheap)
      int S = 1;
      if (qtiles % sms != 0) S = std::max(1, (v2 && kp <= 64 && qtiles * 2 <= sms ? 2 : 1) * sms / qtiles);
      if (const char* e = getenv("RAYUELA_B200_SCAN_SLICES")) S = std::max(1, atoi(e));   // tuning knob
      S = (int)std::min<int64_t>(S, std::max<int64_t>(1, ix->n / (v2 ? unit : 8 * unit)));
      S = std::min(S, std::max(1, 16384 / kp));
      int64_t slice_len = (ix->n + S - 1) / S;
      slice_len = (slice_len + unit - 1) / unit * unit;
      S = (int)((ix->n + slice_len - 1) / slice_len);

      DevBuf cand, part;
      RYL_TRY(cand.alloc((size_t)S * qtiles * QT * cap * sizeof(uint64_t), s));
      RYL_TRY(part.alloc((size_t)S * nqc * kp * sizeof(uint64_t), s));
      if (v2) {
        Scan8Params p;
        p.W = ix->skew.as<uint64_t>();
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
        p.rbmax = rbmax;
        if (norms) {
          RYL_CUDA(cudaFuncSetAttribute(scan8_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          RYL_LAUNCH(scan8_kernel<true>, dim3(qtiles, S), kScan8Warps * 32, smem, s, p);
        } else {
          RYL_CUDA(cudaFuncSetAttribute(scan8_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          RYL_LAUNCH(scan8_kernel<false>, dim3(qtiles, S), kScan8Warps * 32, smem, s, p);
        }
      } else {
        ScanParams p;
        p.codes = ix->codes.as<uint8_t>();
        p.norms = norms;
        p.lut = lut.as<float>();
        p.cand = cand.as<uint64_t>();
        p.part = part.as<uint64_t>();
        p.lb = lbp;
        p.n = ix->n;
        p.slice_len = slice_len;
        p.nq = nqc;
        p.k = kp;
        p.cap = cap;
        RYL_TRY(launch_scan(m, p, norms != nullptr, dim3(qtiles, S), smem, s));
      }
      float* dq = d_out.d + (size_t)qb * k + koff;
      int32_t* iq = i_out.d + (size_t)qb * k + koff;
      RYL_TRY(merge_lists(part.as<uint64_t>(), nullptr, nullptr, S, nqc, kp, dq, iq, id_add, s, k));
      if (koff + kp < k)
        RYL_LAUNCH(lower_bound_kernel, (nqc + 255) / 256, 256, 0, s, dq, iq, nqc, k, kp - 1, id_add, lb.as<uint64_t>());
    }
  }
  RYL_TRY(d_out.flush(s));
  RYL_TRY(i_out.flush(s));
  if (!dev) RYL_CUDA(cudaStreamSynchronize(s));
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
