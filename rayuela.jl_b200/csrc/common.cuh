// common.cuh -- shared host/device helpers for librayuela_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include "../../include/rayuela_b200.h"

namespace ryl {

// ---- error channel ---------------------------------------------------------------------------------
extern thread_local std::string g_err;
extern std::atomic<uint64_t> g_launches;

inline int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define RYL_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      char _b[512];                                                                                 \
      snprintf(_b, sizeof _b, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return ryl::fail(_e == cudaErrorMemoryAllocation ? RAYUELA_ERR_OOM : RAYUELA_ERR_CUDA, _b);    \
    }                                                                                               \
  } while (0)

#define RYL_ARG(cond, msg)                                  \
  do {                                                      \
    if (!(cond)) return ryl::fail(RAYUELA_ERR_ARG, (msg));  \
  } while (0)

#define RYL_TRY(expr)            \
  do {                           \
    int _rc = (expr);            \
    if (_rc != RAYUELA_OK) return _rc; \
  } while (0)

// every kernel launch in the library goes through this (counts launches, checks launch errors)
#define RYL_LAUNCH(kernel, grid, block, smem, stream, ...)                 \
  do {                                                                     \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);            \
    ryl::g_launches.fetch_add(1, std::memory_order_relaxed);               \
    RYL_CUDA(cudaGetLastError());                                          \
  } while (0)

// ---- device buffer (stream-ordered) -----------------------------------------------------------------
// The default memory pool hands freed blocks back to the driver at every synchronisation (release threshold 0),
// which turns each call's scratch allocations into real cudaMalloc/cudaFree.  Keep them cached instead.
inline void keep_pool_memory() {
  static thread_local int done_dev = -1;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev == done_dev) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    uint64_t thr = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  done_dev = dev;
}

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  cudaStream_t s = nullptr;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  int alloc(size_t nbytes, cudaStream_t stream) {
    release();
    s = stream;
    bytes = nbytes;
    if (nbytes == 0) return RAYUELA_OK;
    keep_pool_memory();
    RYL_CUDA(cudaMallocAsync(&p, nbytes, stream));
    return RAYUELA_OK;
  }
  void release() {
    if (p) cudaFreeAsync(p, s);
    p = nullptr;
    bytes = 0;
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

// Input view: device pointer as-is, or a stream-ordered upload of a host array.
template <class T>
struct InArg {
  const T* d = nullptr;
  DevBuf own;
  int bind(const T* src, size_t count, bool is_device, cudaStream_t s) {
    if (is_device || src == nullptr || count == 0) {
      d = src;
      return RAYUELA_OK;
    }
    RYL_TRY(own.alloc(count * sizeof(T), s));
    RYL_CUDA(cudaMemcpyAsync(own.p, src, count * sizeof(T), cudaMemcpyHostToDevice, s));
    d = own.as<T>();
    return RAYUELA_OK;
  }
};

// Output view: device pointer as-is, or a device scratch copied back to the host array by flush().
template <class T>
struct OutArg {
  T* d = nullptr;
  T* host = nullptr;
  size_t count = 0;
  DevBuf own;
  int bind(T* dst, size_t cnt, bool is_device, cudaStream_t s, bool copy_in = false) {
    count = cnt;
    if (is_device || dst == nullptr || cnt == 0) {
      d = dst;
      return RAYUELA_OK;
    }
    host = dst;
    RYL_TRY(own.alloc(cnt * sizeof(T), s));
    d = own.as<T>();
    if (copy_in) RYL_CUDA(cudaMemcpyAsync(d, dst, cnt * sizeof(T), cudaMemcpyHostToDevice, s));
    return RAYUELA_OK;
  }
  int flush(cudaStream_t s) {
    if (host && count) RYL_CUDA(cudaMemcpyAsync(host, d, count * sizeof(T), cudaMemcpyDeviceToHost, s));
    return RAYUELA_OK;
  }
};

inline int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// ---- device set (rayuela_init / RAYUELA_B200_DEVICES) -------------------------------------------------------
// Slots of the configured device list; a device may be listed more than once (two shards on one GPU -- how the
// single-GPU tests exercise the sharded paths).  Empty list = single-device mode on the caller's current device.
struct DeviceSlot {
  int device = 0;
  cudaStream_t stream = nullptr;   // non-blocking stream owned by the library, one per slot
  bool direct_to_first = false;    // kernels on this device may store straight into slot 0's memory (same device, or peer
                                   // access over NVLink enabled): the sharded search writes its results there, no copy
};
const std::vector<DeviceSlot>& device_slots();   // lazily initialised from RAYUELA_B200_DEVICES when rayuela_init was not called

// reference rule for splitting n items in nparts (src/utils.jl:179-203): the first n % nparts parts get one extra
inline void split_range(int64_t n, int nparts, int part, int64_t* a, int64_t* b) {
  const int64_t per = n / nparts, xtra = n % nparts;
  *a = part * per + std::min<int64_t>(part, xtra);
  *b = *a + per + (part < xtra ? 1 : 0);
}

// Runs fn(slot_index) on one host thread per device slot (each thread makes its slot's device current) and returns
// the first non-OK status, with that thread's error message copied to the caller's error channel.
int for_each_slot(const std::vector<DeviceSlot>& slots, const std::function<int(int)>& fn);

// ---- unary_tc.cu: opt-in tcgen05 GEMM  out[l][e] = -2 <C[e], X[l]> + nrm[e]  (bf16x3, fp32 accumulation in TMEM) ----
struct DevBuf;
bool unary_tc_supported(int d, int mh);
int unary_tc_pack_codebooks(const float* C, int d, int mh, DevBuf* Cp, cudaStream_t s);
int unary_tc_launch(const float* X, const DevBuf& Cp, const float* nrm, float* U, unsigned int* umax, int64_t nc, int d,
                    int mh, cudaStream_t s);

// ---- Philox4x32-10 (host + device): the RNG contract shared with the oracle (DESIGN.md "RNG") --------
__host__ __device__ inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    if (r > 0) {
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0;
    c[1] = n1;
    c[2] = n2;
    c[3] = n3;
  }
}
__host__ __device__ inline uint32_t mulhi32(uint32_t a, uint32_t b) {
  return (uint32_t)(((uint64_t)a * b) >> 32);
}

// randperm(m) for ILS iteration `it` (src/LSQ.jl:219): Fisher-Yates over Philox words,
// counter {it,0,0,0xFFFFFFFF-block}.
inline void philox_randperm(uint64_t seed, int it, int m, int* perm) {
  uint32_t w[4];
  int have = 0, blk = 0;
  for (int i = 0; i < m; i++) perm[i] = i;
  for (int i = m - 1; i >= 1; i--) {
    if (have == 0) {
      w[0] = (uint32_t)it;
      w[1] = 0;
      w[2] = 0;
      w[3] = 0xFFFFFFFFu - (uint32_t)blk;
      philox4x32_10(w, (uint32_t)seed, (uint32_t)(seed >> 32));
      have = 4;
      blk++;
    }
    uint32_t r = w[4 - have];
    have--;
    int j = (int)mulhi32(r, (uint32_t)(i + 1));
    int t = perm[i];
    perm[i] = perm[j];
    perm[j] = t;
  }
}

// ---- (dist, id) total order as one 64-bit key --------------------------------------------------------
// std::partial_sort over pair<float,int> orders by dist then id (pairwise_byte.cpp:82).  Map the float to
// an order-preserving uint32 (sign-aware: LSQ distances can be negative), put the id in the low word.
__host__ __device__ inline uint32_t f32_to_ordered(float f) {
  f = f + 0.0f;  // canonicalise -0.0 to +0.0 (equal as floats, must be equal as keys)
  uint32_t u;
#ifdef __CUDA_ARCH__
  u = __float_as_uint(f);
#else
  memcpy(&u, &f, 4);
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ inline float ordered_to_f32(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}
__host__ __device__ inline uint64_t make_key(float dist, uint32_t id) {
  return ((uint64_t)f32_to_ordered(dist) << 32) | id;
}

}  // namespace ryl
