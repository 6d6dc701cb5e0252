// api.cu -- library-wide state of librayuela_b200.so: error channel, device selection, launch counter.
#include "common.cuh"

namespace ryl {
thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};
}  // namespace ryl

extern "C" const char* rayuela_last_error(void) { return ryl::g_err.c_str(); }

extern "C" int rayuela_set_device(int device) {
  RYL_CUDA(cudaSetDevice(device));
  return RAYUELA_OK;
}

extern "C" uint64_t rayuela_launch_count(void) { return ryl::g_launches.load(); }
