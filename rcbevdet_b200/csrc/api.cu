// Library-level entry points of the C ABI (include/rcbevdet_b200.h).
#include <mutex>

#include "common.cuh"

namespace rcb {

int sm_count_cached(int device) {
  static std::mutex mu;
  static int cache[64];
  int dev = device;
  if (dev < 0 && cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev >= 64) {
    int v = 0;
    return cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0 ? v : 148;
  }
  std::lock_guard<std::mutex> lock(mu);
  if (cache[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    cache[dev] = v;
  }
  return cache[dev];
}

static thread_local const int *t_launch_gate = nullptr;
const int *launch_gate() { return t_launch_gate; }

}  // namespace rcb

extern "C" int rcb_set_launch_gate(const int *gate) {
  rcb::t_launch_gate = gate;
  return RCB_OK;
}

extern "C" int rcb_version(void) { return 100; }

extern "C" const char *rcb_error_string(int code) {
  switch (code) {
    case RCB_OK: return "ok";
    case RCB_ERR_ARG: return "invalid argument (null pointer, negative size or inconsistent shape)";
    case RCB_ERR_WORKSPACE: return "workspace too small";
    case RCB_ERR_UNSUPPORTED: return "size outside the supported envelope";
    case RCB_ERR_ALIGN: return "pointer not sufficiently aligned";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
  }
}

extern "C" int rcb_device_info(int device, int *sm_count, int *l2_bytes, int *cc_major, int *cc_minor) {
  int v = 0;
  if (sm_count) {
    RCB_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device));
    *sm_count = v;
  }
  if (l2_bytes) {
    RCB_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, device));
    *l2_bytes = v;
  }
  if (cc_major) {
    RCB_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, device));
    *cc_major = v;
  }
  if (cc_minor) {
    RCB_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, device));
    *cc_minor = v;
  }
  return RCB_OK;
}
