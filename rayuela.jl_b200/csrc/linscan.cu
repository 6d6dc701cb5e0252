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
                                                  float* __restrict__ lut, int nq, int d, int len, int mh) {
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
    if (q < nq) lut[(size_t)q * mh + e0 + e] = acc[i];
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

  const int q0 = blockIdx.x * QT;
  const int slice = blockIdx.y;
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
            int pos = atomicAdd(&cnt_s[q], 1);
            cand[(size_t)q * p.cap + pos] = make_key(s, (uint32_t)i);
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
// K6: merge of S sorted lists per query into the global top-k, by the (dist, id) total order.
//   keys != null : lists are 64-bit keys [S][nq][k] from scan_kernel (ids local; id_add makes them final)
//   else         : lists are (dists, idx) [S][nq][k] (already-final ids; the multi-GPU exchange format)
// One block per query; S*k keys sorted in shared memory.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) merge_kernel(const uint64_t* __restrict__ keys, const float* __restrict__ din,
                                                    const int32_t* __restrict__ iin, int S, int nq, int k,
                                                    float* __restrict__ dout, int32_t* __restrict__ iout,
                                                    int64_t id_add) {
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
    dout[(size_t)q * k + t] = ordered_to_f32((uint32_t)(key >> 32));
    iout[(size_t)q * k + t] = (int32_t)((int64_t)(uint32_t)key + id_add);
  }
}

}  // namespace ryl

// ======================================================================================================
// host side
// ======================================================================================================
using namespace ryl;

struct rayuela_index {
  int kind = 0, m = 0, h = 0, mp = 0, device = 0;
  int64_t n = 0, id_offset = 0;
  DevBuf codes, norms;
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
  RYL_ARG(n >= 1 && n < (1ll << 32), "index_create: n must be in 1..2^32-1");
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
    RYL_TRY(ix->codes.alloc((size_t)n * ix->mp, s));
    int blocks = (int)std::min<int64_t>((n * ix->mp + 255) / 256, 148 * 16);
    RYL_LAUNCH(pad_codes_kernel, blocks, 256, 0, s, raw.d, ix->codes.as<uint8_t>(), n, m, ix->mp);
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
    delete ix;
  }
  return RAYUELA_OK;
}

static int merge_lists(const uint64_t* keys, const float* din, const int32_t* iin, int S, int nq, int k, float* dout,
                       int32_t* iout, int64_t id_add, cudaStream_t s) {
  size_t smem = (size_t)host_pow2ceil(S * k) * sizeof(uint64_t);
  RYL_ARG(smem <= 200 * 1024, "topk merge: S*k too large for a single pass (max 16384 keys... 25600)");
  RYL_CUDA(cudaFuncSetAttribute(merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  RYL_LAUNCH(merge_kernel, nq, 256, smem, s, keys, din, iin, S, nq, k, dout, iout, id_add);
  return RAYUELA_OK;
}

extern "C" int rayuela_index_search(rayuela_index* ix, const float* queries, const float* codebooks, int nq, int d,
                                    int k, float* dists, int32_t* idx, unsigned flags, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  RYL_ARG(ix != nullptr, "index_search: null index");
  RYL_ARG(nq >= 1 && d >= 1, "index_search: nq and d must be positive");
  RYL_ARG(k >= 1 && (int64_t)k <= ix->n, "index_search: k must be in 1..n");
  RYL_ARG(k <= 4096, "index_search: k > 4096 is not supported yet");
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

  const int QT = scan_qt(m);
  const int cap = host_pow2ceil(std::max(2 * kRound, 2 * k + kRound));
  const size_t smem = (size_t)QT * mh * sizeof(float) + (size_t)cap * sizeof(uint64_t);
  RYL_ARG(smem <= 227 * 1024, "index_search: shared-memory budget exceeded");
  const int64_t id_add = (pq ? 0 : 1) + ix->id_offset;  // linscan_aqd.cpp:88 vs pairwise_byte.cpp:76

  const int chunk_q = 16384;
  for (int qb = 0; qb < nq; qb += chunk_q) {
    const int nqc = std::min(chunk_q, nq - qb);
    const int qtiles = (nqc + QT - 1) / QT;
    // DB slices: enough blocks for >= 2 waves, slices no shorter than 8 rounds, S*k within one merge pass
    int S = std::max(1, (2 * sm_count() + qtiles - 1) / qtiles);
    S = (int)std::min<int64_t>(S, std::max<int64_t>(1, ix->n / (8 * kRound)));
    S = std::min(S, std::max(1, 16384 / k));
    int64_t slice_len = (ix->n + S - 1) / S;
    slice_len = (slice_len + kRound - 1) / kRound * kRound;
    S = (int)((ix->n + slice_len - 1) / slice_len);

    DevBuf lut, cand, part;
    RYL_TRY(lut.alloc((size_t)nqc * mh * sizeof(float), s));
    RYL_TRY(cand.alloc((size_t)S * qtiles * QT * cap * sizeof(uint64_t), s));
    RYL_TRY(part.alloc((size_t)S * nqc * k * sizeof(uint64_t), s));

    dim3 lg(mh / 32, (nqc + 31) / 32);
    const float* qptr = q_in.d + (size_t)qb * d;
    if (ix->kind == RAYUELA_SCAN_LSQ)
      RYL_LAUNCH(lut_kernel<RAYUELA_SCAN_LSQ>, lg, 256, 0, s, qptr, cb_in.d, lut.as<float>(), nqc, d, len, mh);
    else if (ix->kind == RAYUELA_SCAN_CQ)
      RYL_LAUNCH(lut_kernel<RAYUELA_SCAN_CQ>, lg, 256, 0, s, qptr, cb_in.d, lut.as<float>(), nqc, d, len, mh);
    else
      RYL_LAUNCH(lut_kernel<RAYUELA_SCAN_PQ>, lg, 256, 0, s, qptr, cb_in.d, lut.as<float>(), nqc, d, len, mh);

    ScanParams p;
    p.codes = ix->codes.as<uint8_t>();
    p.norms = ix->kind == RAYUELA_SCAN_LSQ ? ix->norms.as<float>() : nullptr;
    p.lut = lut.as<float>();
    p.cand = cand.as<uint64_t>();
    p.part = part.as<uint64_t>();
    p.n = ix->n;
    p.slice_len = slice_len;
    p.nq = nqc;
    p.k = k;
    p.cap = cap;
    RYL_TRY(launch_scan(m, p, p.norms != nullptr, dim3(qtiles, S), smem, s));
    RYL_TRY(merge_lists(part.as<uint64_t>(), nullptr, nullptr, S, nqc, k, d_out.d + (size_t)qb * k,
                        i_out.d + (size_t)qb * k, id_add, s));
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
