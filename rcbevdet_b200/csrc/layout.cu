// (n_img, C, HW) planes -> (n_img, HW, C) rows.
// Replaces the transpose copies the reference makes around its kernels:
//   feat.contiguous()      mmdet3d/ops/bev_pool_v2/bev_pool.py:21  (context arrives as a permuted
//                          (B,N,C,H,W) -> (B,N,H,W,C) view, view_transformer.py:195)
//   out_grad.contiguous()  bev_pool.py:69 (gradient of the permute at :91)
// 32x32 shared-memory tiles, both sides coalesced; optional fp32 -> bf16 narrowing on the way.
#include "common.cuh"

namespace rcb {

template <typename TIn, typename TOut>
__device__ __forceinline__ TOut convert_elem(TIn v);
template <>
__device__ __forceinline__ float convert_elem<float, float>(float v) { return v; }
template <>
__device__ __forceinline__ unsigned short convert_elem<unsigned short, unsigned short>(unsigned short v) {
  return v;
}
template <>
__device__ __forceinline__ unsigned short convert_elem<float, unsigned short>(float v) {
  return __bfloat16_as_ushort(__float2bfloat16_rn(v));
}

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) k_planes_to_rows(const TIn *__restrict__ src, TOut *__restrict__ dst,
                                                        int C, int HW, long long src_img_stride, const int *gate) {
  pdl_prologue();
  if (gate_closed(gate)) return;
  __shared__ TOut tile[32][33];
  const int img = blockIdx.z;
  const int hw0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const TIn *s = src + (size_t)img * src_img_stride;
  TOut *d = dst + (size_t)img * HW * C;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int c = c0 + ty + k, hw = hw0 + tx;
    if (c < C && hw < HW) tile[ty + k][tx] = convert_elem<TIn, TOut>(s[(size_t)c * HW + hw]);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int hw = hw0 + ty + k, c = c0 + tx;
    if (c < C && hw < HW) d[(size_t)hw * C + c] = tile[tx][ty + k];
  }
}

// fp32 fast path (C % 4 == 0, HW % 4 == 0, 16-byte aligned): one CTA moves 64 consecutive pixels x
// all C channels with 128-bit global accesses on both sides.  Loads: every channel row contributes
// 256 contiguous bytes; stores: the 64*C output floats are ONE contiguous run.  The generic kernel
// above spends ~36 instructions per element on index arithmetic and is issue-bound (ncu: 75 %
// issue-active at 3.2 TB/s); this one moves four elements per load/store instruction.
#ifndef RCB_V4_PIXELS
#define RCB_V4_PIXELS 64
#endif
constexpr int kV4Pixels = RCB_V4_PIXELS;
#ifndef RCB_V4_MIN_HW
#define RCB_V4_MIN_HW 256
#endif
__global__ void __launch_bounds__(256) k_planes_to_rows_v4(const float *__restrict__ src, float *__restrict__ dst,
                                                           int C, int HW, long long src_img_stride, const int *gate) {
  pdl_prologue();
  if (gate_closed(gate)) return;
  extern __shared__ float tile_v4[];  // [kV4Pixels][C + 1]
  const int P = C + 1;
  const int img = blockIdx.y;
  const int hw0 = blockIdx.x * kV4Pixels;
  const int n_hw = min(kV4Pixels, HW - hw0);  // multiple of 4
  const float *s = src + (size_t)img * src_img_stride + hw0;
  const int quads_per_row = n_hw >> 2;
  // five 128-bit loads in flight per thread before the first shared-memory store (C = 80: the whole
  // tile in one trip); left as a plain loop the compiler keeps one load per trip in flight
  constexpr int kBatch = 5;
  const int n_f = C * quads_per_row;
  for (int f0 = threadIdx.x; f0 < n_f; f0 += 256 * kBatch) {
    float4 v[kBatch];
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int f = f0 + 256 * u;
      if (f < n_f) {
        const int c = f / quads_per_row, j = f - c * quads_per_row;
        v[u] = ld_stream_f4(reinterpret_cast<const float4 *>(s + (size_t)c * HW) + j);
      }
    }
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int f = f0 + 256 * u;
      if (f < n_f) {
        const int c = f / quads_per_row, j = f - c * quads_per_row;
        float *t = tile_v4 + (4 * j) * P + c;
        t[0] = v[u].x, t[P] = v[u].y, t[2 * P] = v[u].z, t[3 * P] = v[u].w;
      }
    }
  }
  __syncthreads();
  float4 *d = reinterpret_cast<float4 *>(dst + ((size_t)img * HW + hw0) * C);
  const int c_quads = C >> 2;
  for (int g = threadIdx.x; g < n_hw * c_quads; g += 256) {
    const int hw = g / c_quads, cq = g - hw * c_quads;
    const float *t = tile_v4 + hw * P + 4 * cq;
    st_stream_f4(d + g, make_float4(t[0], t[1], t[2], t[3]));
  }
}

int planes_to_rows_launch(const void *src, void *dst, int n_img, int C, int HW,
                          long long src_img_stride, int elem_bytes, cudaStream_t s, const int *gate) {
  if (n_img <= 0 || C <= 0 || HW <= 0) return RCB_OK;
  if (n_img > 65535) return RCB_ERR_UNSUPPORTED;
  if (elem_bytes == 4 && (C % 4) == 0 && (HW % 4) == 0 && HW >= RCB_V4_MIN_HW && C <= 256 && (src_img_stride % 4) == 0 &&
      (((uintptr_t)src) % 16) == 0 && (((uintptr_t)dst) % 16) == 0) {
    const size_t smem = (size_t)kV4Pixels * (C + 1) * 4;
    if (smem > 48 * 1024)
      RCB_CUDA_TRY(cudaFuncSetAttribute(k_planes_to_rows_v4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 g4(ceil_div(HW, kV4Pixels), n_img);
    RCB_CUDA_TRY(launch_pdl(k_planes_to_rows_v4, g4, 256, smem, s, (const float *)src, (float *)dst, C, HW, src_img_stride, gate));
    return RCB_OK;
  }
  dim3 grid(ceil_div(HW, 32), ceil_div(C, 32), n_img);
  if (elem_bytes == 4)
    RCB_CUDA_TRY(launch_pdl(k_planes_to_rows<float, float>, grid, 256, 0, s, (const float *)src, (float *)dst, C, HW, src_img_stride, gate));
  else if (elem_bytes == 2)
    k_planes_to_rows<unsigned short, unsigned short>
        <<<grid, 256, 0, s>>>((const unsigned short *)src, (unsigned short *)dst, C, HW, src_img_stride, gate);
  else if (elem_bytes == -2)  // fp32 in, bf16 out
    k_planes_to_rows<float, unsigned short>
        <<<grid, 256, 0, s>>>((const float *)src, (unsigned short *)dst, C, HW, src_img_stride, gate);
  else
    return RCB_ERR_ARG;
  RCB_LAUNCH_CHECK();
  return RCB_OK;
}

}  // namespace rcb

using namespace rcb;

extern "C" int rcb_planes_to_rows(const void *src, void *dst, int n_img, int C, int HW,
                                  long long src_img_stride, int elem_bytes, int device,
                                  rcb_stream_t stream) {
  if (!src || !dst) return (n_img > 0 && C > 0 && HW > 0) ? RCB_ERR_ARG : RCB_OK;
  if (n_img < 0 || C < 0 || HW < 0) return RCB_ERR_ARG;
  DeviceGuard guard(device);
  if (guard.err) return guard.err;
  return planes_to_rows_launch(src, dst, n_img, C, HW, src_img_stride, elem_bytes, (cudaStream_t)stream, nullptr);
}
