// api.cu -- library-wide state of librayuela_b200.so: error channel, device selection / device set, launch counter.
#include <mutex>
#include <thread>

#include "common.cuh"

namespace ryl {
thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};

static std::mutex g_dev_mu;
static std::vector<DeviceSlot> g_slots;
static bool g_slots_ready = false;

static void destroy_slots_locked() {
  int cur = 0;
  cudaGetDevice(&cur);
  for (auto& sl : g_slots)
    if (sl.stream) {
      cudaSetDevice(sl.device);
      cudaStreamSynchronize(sl.stream);
      cudaStreamDestroy(sl.stream);
    }
  g_slots.clear();
  cudaSetDevice(cur);
}

static int set_slots_locked(const int* devices, int n) {
  int ndev = 0, cur = 0;
  RYL_CUDA(cudaGetDeviceCount(&ndev));
  RYL_CUDA(cudaGetDevice(&cur));
  for (int i = 0; i < n; i++)
    RYL_ARG(devices[i] >= 0 && devices[i] < ndev, "rayuela_init: device index out of range");
  destroy_slots_locked();
  for (int i = 0; i < n; i++) {
    DeviceSlot sl;
    sl.device = devices[i];
    RYL_CUDA(cudaSetDevice(sl.device));
    RYL_CUDA(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
    g_slots.push_back(sl);
  }
  // peer access (NVLink): a shard's search kernels store their results straight into slot 0's gather buffer; without
  // it (or on a failure) the results are staged locally and copied with cudaMemcpyPeerAsync
  for (int i = 0; i < n; i++) {
    g_slots[i].direct_to_first = devices[i] == devices[0];
    for (int j = 0; j < n; j++)
      if (devices[i] != devices[j]) {
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, devices[i], devices[j]) == cudaSuccess && can) {
          cudaSetDevice(devices[i]);
          cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
          if (e != cudaSuccess) cudaGetLastError();   // cudaErrorPeerAccessAlreadyEnabled: fine
          if (j == 0 && (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled)) {
            // the gather buffer comes from slot 0's stream-ordered pool, which is mapped on peers only on request
            cudaMemPool_t pool;
            cudaMemAccessDesc acc = {};
            acc.location.type = cudaMemLocationTypeDevice;
            acc.location.id = devices[i];
            acc.flags = cudaMemAccessFlagsProtReadWrite;
            if (cudaDeviceGetDefaultMemPool(&pool, devices[0]) == cudaSuccess &&
                cudaMemPoolSetAccess(pool, &acc, 1) == cudaSuccess)
              g_slots[i].direct_to_first = true;
            else
              cudaGetLastError();
          }
        }
      }
  }
  RYL_CUDA(cudaSetDevice(cur));
  g_slots_ready = true;
  return RAYUELA_OK;
}

const std::vector<DeviceSlot>& device_slots() {
  std::lock_guard<std::mutex> lk(g_dev_mu);
  if (!g_slots_ready) {
    g_slots_ready = true;
    if (const char* e = getenv("RAYUELA_B200_DEVICES")) {    // "0,1,2,3" (device list from the environment)
      std::vector<int> devs;
      const char* p = e;
      while (*p) {
        char* end = nullptr;
        long v = strtol(p, &end, 10);
        if (end == p) break;
        devs.push_back((int)v);
        p = (*end == ',') ? end + 1 : end;
      }
      if (devs.size() > 1 && set_slots_locked(devs.data(), (int)devs.size()) != RAYUELA_OK) {
        fprintf(stderr, "librayuela_b200: RAYUELA_B200_DEVICES=%s ignored: %s\n", e, g_err.c_str());
        g_slots.clear();
      }
    }
  }
  return g_slots;
}

int for_each_slot(const std::vector<DeviceSlot>& slots, const std::function<int(int)>& fn) {
  const int D = (int)slots.size();
  std::vector<int> rc(D, RAYUELA_OK);
  std::vector<std::string> msg(D);
  std::vector<std::thread> th;
  th.reserve(D);
  for (int i = 0; i < D; i++)
    th.emplace_back([&, i]() {
      if (cudaSetDevice(slots[i].device) != cudaSuccess) {
        rc[i] = RAYUELA_ERR_CUDA;
        msg[i] = "cudaSetDevice failed for a configured device";
        return;
      }
      rc[i] = fn(i);
      if (rc[i] != RAYUELA_OK) msg[i] = g_err;   // the worker thread's own error channel
    });
  for (auto& t : th) t.join();
  for (int i = 0; i < D; i++)
    if (rc[i] != RAYUELA_OK) return fail(rc[i], "device slot " + std::to_string(i) + ": " + msg[i]);
  return RAYUELA_OK;
}
}  // namespace ryl

extern "C" const char* rayuela_last_error(void) { return ryl::g_err.c_str(); }

extern "C" int rayuela_set_device(int device) {
  RYL_CUDA(cudaSetDevice(device));
  return RAYUELA_OK;
}

extern "C" uint64_t rayuela_launch_count(void) { return ryl::g_launches.load(); }

extern "C" int rayuela_init(const int* devices, int n_devices) {
  std::lock_guard<std::mutex> lk(ryl::g_dev_mu);
  RYL_ARG(n_devices >= 0 && (n_devices == 0 || devices), "rayuela_init: bad device list");
  if (n_devices <= 1) {               // single-device mode: calls run on the caller's current device
    ryl::destroy_slots_locked();
    ryl::g_slots_ready = true;
    if (n_devices == 1) RYL_CUDA(cudaSetDevice(devices[0]));
    return RAYUELA_OK;
  }
  return ryl::set_slots_locked(devices, n_devices);
}

extern "C" int rayuela_shutdown(void) {
  std::lock_guard<std::mutex> lk(ryl::g_dev_mu);
  ryl::destroy_slots_locked();
  ryl::g_slots_ready = true;
  return RAYUELA_OK;
}

extern "C" int rayuela_device_count(void) {
  const auto& s = ryl::device_slots();
  return s.empty() ? 1 : (int)s.size();
}
