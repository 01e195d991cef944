// Shared helpers for the sm_100a kernels of the BEV pooling path.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rcbevdet_b200.h"

#define RCB_CUDA_TRY(expr)                 \
  do {                                     \
    cudaError_t _e = (expr);               \
    if (_e != cudaSuccess) return (int)_e; \
  } while (0)

#define RCB_LAUNCH_CHECK()                 \
  do {                                     \
    cudaError_t _e = cudaGetLastError();   \
    if (_e != cudaSuccess) return (int)_e; \
  } while (0)

namespace rcb {

// Programmatic dependent launch: a kernel launched with launch_pdl() may be scheduled while the
// previous kernel of the stream is still draining; pdl_prologue() -- the first statement of every
// such kernel -- blocks until that kernel has completed and its writes are visible, then lets the
// next launch do the same.  It hides launch latency and the ramp-up of each grid, nothing else;
// under a normal launch both instructions are no-ops.
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args &&...args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Launch gate (rcb_set_launch_gate, api.cu): while the calling host thread has a gate set, the kernels
// of the general entry points (prepare, cell-stationary forward, pixel-stationary backward and its
// out_grad transpose) are launched as usual but exit at once when *gate == 0.  It is how a chain
// enqueues its fallback behind the strip kernels without reading their plan's status word back: the
// gate IS that status word (0 = the strip kernels ran, non-zero = they refused and exited).
const int *launch_gate();
__device__ __forceinline__ bool gate_closed(const int *gate) { return gate != nullptr && *gate == 0; }

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

__host__ __device__ inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// 128-bit read-only loads.  feat rows / out_grad rows are re-read by many warps: keep them in L1.
__device__ __forceinline__ float4 ldg_f4(const float4 *p) { return __ldg(p); }

// streaming (touch-once) loads and stores: do not pollute L1
__device__ __forceinline__ float ld_stream_f32(const float *p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int ld_stream_s32(const int *p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ld_stream_f4(const float4 *p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream_f32(float *p, float v) {
  asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v));
}
__device__ __forceinline__ void st_stream_s32(int *p, int v) {
  asm volatile("st.global.L1::no_allocate.s32 [%0], %1;" ::"l"(p), "r"(v));
}
__device__ __forceinline__ void st_stream_f4(float4 *p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w));
}

// Four consecutive channels of a channels-last row, widened to fp32.
template <typename T>
struct Row4;
template <>
struct Row4<float> {
  static __device__ __forceinline__ float4 load(const float *base, size_t quad) {
    return __ldg(reinterpret_cast<const float4 *>(base) + quad);
  }
  static __device__ __forceinline__ float4 load_bytes(const char *p) {
    return __ldg(reinterpret_cast<const float4 *>(p));
  }
};
template <>
struct Row4<__nv_bfloat16> {
  static __device__ __forceinline__ float4 load_bytes(const char *p) {
    uint2 raw = __ldg(reinterpret_cast<const uint2 *>(p));
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162 *>(&raw.x);
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162 *>(&raw.y);
    float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
  }
  static __device__ __forceinline__ float4 load(const __nv_bfloat16 *base, size_t quad) {
    uint2 raw = __ldg(reinterpret_cast<const uint2 *>(base) + quad);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162 *>(&raw.x);
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162 *>(&raw.y);
    float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
  }
};
template <>
struct Row4<__half> {
  static __device__ __forceinline__ float4 load_bytes(const char *p) {
    uint2 raw = __ldg(reinterpret_cast<const uint2 *>(p));
    __half2 a = *reinterpret_cast<__half2 *>(&raw.x);
    __half2 b = *reinterpret_cast<__half2 *>(&raw.y);
    float2 fa = __half22float2(a), fb = __half22float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
  }
  static __device__ __forceinline__ float4 load(const __half *base, size_t quad) {
    uint2 raw = __ldg(reinterpret_cast<const uint2 *>(base) + quad);
    __half2 a = *reinterpret_cast<__half2 *>(&raw.x);
    __half2 b = *reinterpret_cast<__half2 *>(&raw.y);
    float2 fa = __half22float2(a), fb = __half22float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
  }
};

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }

// Exact unsigned 32-bit division by a run-time constant (Granlund-Montgomery round-up method):
// q = (t + ((n - t) >> s1)) >> s2 with t = umulhi(m, n).  Valid for every n < 2^32, d >= 1.
struct FastDiv {
  unsigned m, s1, s2, d;
  __host__ static FastDiv make(unsigned d) {
    FastDiv f;
    unsigned l = 0;
    while ((1ull << l) < d) ++l;
    f.m = (unsigned)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
    f.s1 = l < 1 ? l : 1;
    f.s2 = l > 0 ? l - 1 : 0;
    f.d = d;
    return f;
  }
  __device__ __forceinline__ unsigned div(unsigned n) const {
    const unsigned t = __umulhi(m, n);
    return (t + ((n - t) >> s1)) >> s2;
  }
};

struct DeviceGuard {
  int prev = -1;
  int err = 0;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (device >= 0 && device != prev) err = (int)cudaSetDevice(device);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

int sm_count_cached(int device);
int check_pool_desc(const rcb_pool_desc *d);

}  // namespace rcb
