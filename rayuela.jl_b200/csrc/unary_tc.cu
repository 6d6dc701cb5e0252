// unary_tc.cu -- OPT-IN fast mode for the one dense contraction of path (1): the unaries
//     U[l][e] = -2 <C[e], x_l> + ||C[e]||^2            (get_unaries, src/utils.jl:121-149; CUBLAS sgemm + vec_add in the
//                                                        reference's GPU flavour, src/LSQ_GPU.jl:74-78)
// as a tcgen05 (5th-generation tensor core) GEMM with TMEM accumulators.  The DEFAULT stays the exact fp32 SIMT kernel
// (icm.cu, unary_kernel): its per-output fmaf chain is what makes the encode bit-identical to the oracle.  This mode is
// selected per call with the RAYUELA_FAST_UNARIES flag (or RAYUELA_B200_FAST_UNARIES=1) and is validated by tolerance
// (|U_fast - U_exact| <= 2^-15 * sum_t |x_t c_t|), qerror and code-agreement statistics, never by bit identity.
//
// Precision: fp32 inputs are split x = hi + lo into two bf16 (lo = bf16(x - hi)); the product is formed as
// hi*hi + hi*lo + lo*hi in three kind::f16 MMAs accumulating in fp32 ("bf16x3", relative error ~2^-16 per product).
//
// Data movement: a pre-pass (split_pack_kernel) writes both operands as ready-made shared-memory images -- one 128-byte
// row per 64 bf16 of K, 16-byte chunks XOR-swizzled by (row & 7), i.e. exactly the canonical K-major SWIZZLE_128B layout
// tcgen05.mma reads -- so a tile is fetched with plain 1-D bulk async copies (cp.async.bulk + mbarrier complete_tx), no
// tensor maps.  Kernel: one persistent CTA per SM, 10 warps: warp 0 = bulk-copy producer, warp 1 = TMEM allocator + MMA
// issuer (one elected lane), warps 2..9 = epilogue (tcgen05.ld 32x32b -> fma(-2, acc, ||c||^2) -> XOR-swizzled 4 KB staging
// tile per warp -> transposed read-back -> full 128-byte-line global stores; per-vector max |U| for K3's pre-filter slack).  Tile = 128 vectors x 256 entries (one codebook) x K = d <= 128; the A images of an
// M-tile stay in shared memory for all m codebooks; two 256-column TMEM accumulators let the epilogue of codebook j
// overlap the MMAs of codebook j+1.
#include <cuda_bf16.h>

#include "common.cuh"

namespace ryl {

static constexpr int kTcM = 128;       // vectors per tile (TMEM lanes)
static constexpr int kTcN = 256;       // entries per tile (one codebook; TMEM columns per accumulator)
static constexpr int kTcKB = 64;       // bf16 per 128-byte swizzle row
static constexpr int kTcMaxKB = 2;     // d <= 128

// ---- pre-pass: fp32 rows -> bf16 hi / lo, in SWIZZLE_128B tile images --------------------------------------------------
// image of (tile t, k-block kb, part p): TR rows x 128 bytes at ((t*KB + kb)*2 + p) * TR*128; chunk c of row r at
// r*128 + ((c ^ (r & 7)) * 16).  Rows >= R and columns >= d are zero.
__global__ void __launch_bounds__(256) split_pack_kernel(const float* __restrict__ src, int64_t R, int d, int TR, int KB,
                                                         uint4* __restrict__ dst, int64_t ntiles) {
  const int64_t total = ntiles * KB * TR * 8;      // 16-byte chunks per part
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i & 7);
    const int r = (int)((i >> 3) % TR);
    const int kb = (int)((i / (8 * (int64_t)TR)) % KB);
    const int64_t t = i / (8 * (int64_t)TR * KB);
    const int64_t row = t * TR + r;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; e++) {
      const int col = kb * kTcKB + c * 8 + e;
      v[e] = (row < R && col < d) ? src[(size_t)row * d + col] : 0.f;
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * e]), h1 = __float2bfloat16_rn(v[2 * e + 1]);
      const __nv_bfloat16 l0 = __float2bfloat16_rn(v[2 * e] - __bfloat162float(h0));
      const __nv_bfloat16 l1 = __float2bfloat16_rn(v[2 * e + 1] - __bfloat162float(h1));
      hi[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
      lo[e] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    const size_t img = ((size_t)(t * KB + kb) * 2) * TR * 8;        // in uint4 units
    const size_t off = (size_t)r * 8 + (c ^ (r & 7));
    dst[img + off] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    dst[img + (size_t)TR * 8 + off] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// ---- PTX wrappers -----------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// bounded wait: a protocol bug traps (the launch fails with an error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  long long t0 = 0;
  for (uint32_t spin = 0; !ok; spin++) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!ok && (spin & 1023u) == 1023u) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ll) __trap();                   // ~2 s at 2 GHz
    }
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32, issued by one thread
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns of TMEM -> 32 registers per thread (thread = lane = row)
__device__ __forceinline__ void tmem_ld32(uint32_t (&r)[32], uint32_t taddr) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
// K-major SWIZZLE_128B shared-memory matrix descriptor: start address >> 4, LBO (ignored for swizzled K-major) = 1,
// SBO = 1024 bytes (8 rows x 128 B) >> 4, descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t tc_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D = f32 (bits 4-5 = 1), A = B = bf16 (bits 7-9, 10-12 = 1), both K-major, N >> 3 at 17, M >> 4 at 24
static constexpr uint32_t kTcIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTcN >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);

struct UnaryTcParams {
  const uint4* A;      // packed X images: [mtiles][KB][2][128 rows][8 chunks]
  const uint4* Bp;     // packed C images: [ntiles][KB][2][256 rows][8 chunks]
  const float* nrm;    // [mh]
  float* U;            // [n][mh]
  unsigned int* umax;  // [n] float bits, or null
  int64_t n;
  int mh, KB;
  int64_t mtiles;
  int ntiles;
};

__global__ void __launch_bounds__(320, 1) unary_tc_kernel(UnaryTcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // [A: KB x 2 x 16 KB][B: KB x 2 x 32 KB] (1024-byte aligned images), then barriers
  const uint32_t smem0 = (tc_smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_img = 128 * 128, b_img = 256 * 128;              // bytes per image
  const uint32_t a_base = smem0, b_base = smem0 + (uint32_t)p.KB * 2 * a_img;
  __shared__ __align__(8) uint64_t bars[2 + 2 * kTcMaxKB + 4];
  __shared__ uint32_t tmem_slot;
  const uint32_t bar_a_full = tc_smem_u32(&bars[0]), bar_a_empty = tc_smem_u32(&bars[1]);
  auto bar_b_full = [&](int kb) { return tc_smem_u32(&bars[2 + kb]); };
  auto bar_b_empty = [&](int kb) { return tc_smem_u32(&bars[2 + kTcMaxKB + kb]); };
  auto bar_t_full = [&](int b) { return tc_smem_u32(&bars[2 + 2 * kTcMaxKB + b]); };
  auto bar_t_empty = [&](int b) { return tc_smem_u32(&bars[4 + 2 * kTcMaxKB + b]); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar_a_full, 1);
    mbar_init(bar_a_empty, 1);
    for (int kb = 0; kb < kTcMaxKB; kb++) {
      mbar_init(bar_b_full(kb), 1);
      mbar_init(bar_b_empty(kb), 1);
    }
    for (int b = 0; b < 2; b++) {
      mbar_init(bar_t_full(b), 1);
      mbar_init(bar_t_empty(b), 8);                                  // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {                                                   // TMEM: 2 accumulators x 256 columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    // ===== producer: A images once per M-tile, B images per (M-tile, codebook), stage = k-block =====
    if (lane == 0) {
      uint32_t a_phase = 0, b_phase[kTcMaxKB] = {0, 0};
      for (int64_t mt = blockIdx.x; mt < p.mtiles; mt += gridDim.x) {
        mbar_wait(bar_a_empty, a_phase ^ 1);                         // the MMAs of the previous M-tile are done with A
        a_phase ^= 1;
        mbar_expect_tx(bar_a_full, (uint32_t)p.KB * 2 * a_img);
        const char* asrc = reinterpret_cast<const char*>(p.A) + (size_t)mt * p.KB * 2 * a_img;
        for (int i = 0; i < p.KB * 2; i++) bulk_g2s(a_base + i * a_img, asrc + (size_t)i * a_img, a_img, bar_a_full);
        for (int nt = 0; nt < p.ntiles; nt++) {
          for (int kb = 0; kb < p.KB; kb++) {
            mbar_wait(bar_b_empty(kb), b_phase[kb] ^ 1);
            b_phase[kb] ^= 1;
            mbar_expect_tx(bar_b_full(kb), 2 * b_img);
            const char* bsrc = reinterpret_cast<const char*>(p.Bp) + ((size_t)nt * p.KB + kb) * 2 * b_img;
            bulk_g2s(b_base + (kb * 2 + 0) * b_img, bsrc, b_img, bar_b_full(kb));
            bulk_g2s(b_base + (kb * 2 + 1) * b_img, bsrc + b_img, b_img, bar_b_full(kb));
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      uint32_t a_phase = 0, b_phase[kTcMaxKB] = {0, 0}, t_phase[2] = {0, 0};
      uint32_t tile = 0;
      for (int64_t mt = blockIdx.x; mt < p.mtiles; mt += gridDim.x) {
        mbar_wait(bar_a_full, a_phase);
        a_phase ^= 1;
        for (int nt = 0; nt < p.ntiles; nt++, tile++) {
          const int buf = tile & 1;
          mbar_wait(bar_t_empty(buf), t_phase[buf] ^ 1);             // the epilogue has drained this accumulator
          t_phase[buf] ^= 1;
          tc_fence_after();
          const uint32_t dcol = tmem + buf * kTcN;
          uint32_t acc = 0;
          for (int kb = 0; kb < p.KB; kb++) {
            mbar_wait(bar_b_full(kb), b_phase[kb]);
            b_phase[kb] ^= 1;
            tc_fence_after();
            // hi*hi, hi*lo, lo*hi over the 4 UMMA_K = 16 steps of this 64-wide k-block (+32 bytes per step)
#pragma unroll
            for (int pr = 0; pr < 3; pr++) {
              const int pa = pr == 2 ? 1 : 0, pb = pr == 1 ? 1 : 0;
              const uint32_t aaddr = a_base + (kb * 2 + pa) * a_img, baddr = b_base + (kb * 2 + pb) * b_img;
#pragma unroll
              for (int k = 0; k < 4; k++) {
                tc_mma(dcol, tc_desc(aaddr + k * 32), tc_desc(baddr + k * 32), kTcIdesc, acc);
                acc = 1;
              }
            }
            tc_commit(bar_b_empty(kb));                              // stage kb may be refilled once these MMAs retire
          }
          tc_commit(bar_t_full(buf));
          if (nt == p.ntiles - 1) tc_commit(bar_a_empty);
        }
      }
    }
  } else {
    // ===== epilogue: 8 warps; warp w may only read TMEM lanes 32*(w%4).. (= 32 rows of the tile), so two warps share a
    // lane quarter and split the 256 columns.  The tcgen05.ld of the next 32-column chunk is in flight while the current
    // one is scaled, biased and stored =====
    const int q = warp & 3, half = (warp - 2) >> 2;
    float4* stage = reinterpret_cast<float4*>(smem_raw + (smem0 - tc_smem_u32(smem_raw)) + (size_t)p.KB * 2 * (a_img + b_img)) +
                    (warp - 2) * 256;                       // 32 rows x 8 quads of 16 bytes = 4 KB per epilogue warp
    uint32_t t_phase[2] = {0, 0};
    uint32_t tile = 0;
    for (int64_t mt = blockIdx.x; mt < p.mtiles; mt += gridDim.x) {
      const int64_t l = mt * kTcM + q * 32 + lane;
      float mx = 0.f;
      bool bad = false;
      for (int nt = 0; nt < p.ntiles; nt++, tile++) {
        const int buf = tile & 1;
        mbar_wait(bar_t_full(buf), t_phase[buf]);
        t_phase[buf] ^= 1;
        tc_fence_after();
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + buf * kTcN + half * (kTcN / 2);
        const float* nr = p.nrm + (size_t)nt * kTcN + half * (kTcN / 2);
        uint32_t r[2][32];
        tmem_ld32(r[0], taddr);
#pragma unroll
        for (int ch = 0; ch < kTcN / 2 / 32; ch++) {
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (ch + 1 < kTcN / 2 / 32) tmem_ld32(r[(ch + 1) & 1], taddr + (ch + 1) * 32);
          // thread = row: scale, bias, track max |U|, and park the 32 x 32 chunk in this warp's 4 KB staging tile with
          // the 16-byte column quads XOR-swizzled by the row, so that both the row-wise writes and the transposed reads
          // below are bank-conflict free
          __syncwarp();                                    // the previous chunk's reads of the staging tile are done
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            const float4 nv = __ldg(reinterpret_cast<const float4*>(nr + ch * 32 + e));
            float4 o;
            o.x = fmaf(-2.0f, __uint_as_float(r[ch & 1][e]), nv.x);
            o.y = fmaf(-2.0f, __uint_as_float(r[ch & 1][e + 1]), nv.y);
            o.z = fmaf(-2.0f, __uint_as_float(r[ch & 1][e + 2]), nv.z);
            o.w = fmaf(-2.0f, __uint_as_float(r[ch & 1][e + 3]), nv.w);
            stage[lane * 8 + ((e >> 2) ^ (lane & 7))] = o;
            mx = fmaxf(fmaxf(mx, fmaxf(fabsf(o.x), fabsf(o.y))), fmaxf(fabsf(o.z), fabsf(o.w)));
            bad |= (o.x != o.x) | (o.y != o.y) | (o.z != o.z) | (o.w != o.w);
          }
          __syncwarp();
          // transposed read-back: one instruction stores 4 rows x 128 contiguous bytes (full lines) instead of 16 bytes
          // into each of 32 different lines
#pragma unroll
          for (int i = 0; i < 8; i++) {
            const int row = 4 * i + (lane >> 3), quad = lane & 7;
            const float4 o = stage[row * 8 + (quad ^ (row & 7))];
            const int64_t lr = mt * kTcM + q * 32 + row;
            if (lr < p.n)
              *reinterpret_cast<float4*>(p.U + (size_t)lr * p.mh + (size_t)nt * kTcN + half * (kTcN / 2) + ch * 32 + quad * 4) = o;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_t_empty(buf));
      }
      // two warps hold parts of a row: combine with atomicMax on the float bits (non-negative floats order like integers)
      if (p.umax && l < p.n) atomicMax(p.umax + l, __float_as_uint(bad ? __int_as_float(0x7f800000) : mx));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

}  // namespace ryl

using namespace ryl;

namespace ryl {
bool unary_tc_supported(int d, int mh) { return d >= 1 && d <= kTcKB * kTcMaxKB && mh % kTcN == 0; }

// U[l][e] = -2 <C[e], X[l]> + nrm[e] for nc vectors (device pointers); umax as unary_kernel's epilogue (may be null).
// Cp = packed codebooks from unary_tc_pack_codebooks (packed once per encode call); scratch for X is per chunk.
int unary_tc_pack_codebooks(const float* C, int d, int mh, DevBuf* Cp, cudaStream_t s) {
  const int KB = (d + kTcKB - 1) / kTcKB, ntiles = mh / kTcN;
  RYL_TRY(Cp->alloc((size_t)ntiles * KB * 2 * kTcN * 128, s));
  const int64_t chunks = (int64_t)ntiles * KB * kTcN * 8;
  RYL_LAUNCH(split_pack_kernel, (int)std::min<int64_t>((chunks + 255) / 256, (int64_t)sm_count() * 8), 256, 0, s, C,
             (int64_t)mh, d, kTcN, KB, Cp->as<uint4>(), (int64_t)ntiles);
  return RAYUELA_OK;
}

int unary_tc_launch(const float* X, const DevBuf& Cp, const float* nrm, float* U, unsigned int* umax, int64_t nc, int d,
                    int mh, cudaStream_t s) {
  const int KB = (d + kTcKB - 1) / kTcKB, ntiles = mh / kTcN;
  const int64_t mtiles = (nc + kTcM - 1) / kTcM;
  DevBuf Ap;
  RYL_TRY(Ap.alloc((size_t)mtiles * KB * 2 * kTcM * 128, s));
  const int64_t chunks = mtiles * KB * kTcM * 8;
  RYL_LAUNCH(split_pack_kernel, (int)std::min<int64_t>((chunks + 255) / 256, (int64_t)sm_count() * 16), 256, 0, s, X, nc,
             d, kTcM, KB, Ap.as<uint4>(), mtiles);
  UnaryTcParams p;
  p.A = Ap.as<uint4>();
  p.Bp = Cp.as<uint4>();
  p.nrm = nrm;
  p.U = U;
  p.umax = umax;
  p.n = nc;
  p.mh = mh;
  p.KB = KB;
  p.mtiles = mtiles;
  p.ntiles = ntiles;
  const size_t smem = (size_t)KB * 2 * (kTcM * 128 + kTcN * 128) + 8 * 4096 + 1024;   // images + staging tiles + alignment
  RYL_CUDA(cudaFuncSetAttribute(unary_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)std::min<int64_t>(mtiles, sm_count());
  RYL_LAUNCH(unary_tc_kernel, grid, 320, smem, s, p);
  return RAYUELA_OK;
}
}  // namespace ryl
