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
                                                        int C, int HW, long long src_img_stride) {
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

int planes_to_rows_launch(const void *src, void *dst, int n_img, int C, int HW,
                          long long src_img_stride, int elem_bytes, cudaStream_t s) {
  if (n_img <= 0 || C <= 0 || HW <= 0) return RCB_OK;
  if (n_img > 65535) return RCB_ERR_UNSUPPORTED;
  dim3 grid(ceil_div(HW, 32), ceil_div(C, 32), n_img);
  if (elem_bytes == 4)
    k_planes_to_rows<float, float><<<grid, 256, 0, s>>>((const float *)src, (float *)dst, C, HW, src_img_stride);
  else if (elem_bytes == 2)
    k_planes_to_rows<unsigned short, unsigned short>
        <<<grid, 256, 0, s>>>((const unsigned short *)src, (unsigned short *)dst, C, HW, src_img_stride);
  else if (elem_bytes == -2)  // fp32 in, bf16 out
    k_planes_to_rows<float, unsigned short>
        <<<grid, 256, 0, s>>>((const float *)src, (unsigned short *)dst, C, HW, src_img_stride);
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
  return planes_to_rows_launch(src, dst, n_img, C, HW, src_img_stride, elem_bytes, (cudaStream_t)stream);
}
