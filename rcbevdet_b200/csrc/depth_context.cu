// SURVEY.md 8(f-2) -- the producer of the pooling inputs, fused: LSSViewTransformer.forward
// (mmdet3d/models/necks/view_transformer.py:316-319; the BEVDepth variant :793-797):
//     depth_digit = x[:, :D];  tran_feat = x[:, D:D+C];  depth = depth_digit.softmax(dim=1)
// followed, inside the pooling op, by the channels-last copy of the context (bev_pool.py:21).
// The reference runs a softmax kernel over a strided slice and, later, a strided transpose copy of
// the other slice.  Here ONE kernel reads the depth-net output x (n_img, D + C, H, W) once and
// writes both consumers' formats: depth (n_img, D, H, W) and the context as channels-last rows
// (n_img * H * W, C) -- the rows the pooling kernels gather -- so rcb_planes_to_rows disappears from
// the step.  The backward kernel is the mirror image: softmax backward on the depth slice and the
// rows -> planes transpose of the context gradient, written straight into d(x).
//
// A CTA owns 32 consecutive pixels of one image, lane <-> pixel, warp <-> every 8th channel: all
// global accesses are 128-byte rows of one channel plane, or (context rows) one 32 * C * 4 byte run.
// A thread keeps its <= kMaxD / 8 depth values in registers between the max, the sum and the
// normalisation: x is read exactly once.  softmax = exp(x - max) / sum with expf and a true
// division, as torch's kernel computes it; the sums run in a different order (rel. 1e-7).
#include "common.cuh"

namespace rcb {

constexpr int kDcWarps = 8;
constexpr int kDcPerWarp = 32;  // depth bins a thread may hold: D <= 256

template <typename T>
__device__ __forceinline__ float dc_load(const T *p);
template <>
__device__ __forceinline__ float dc_load<float>(const float *p) { return ld_stream_f32(p); }
template <>
__device__ __forceinline__ float dc_load<__half>(const __half *p) { return __half2float(__ldg(p)); }
template <>
__device__ __forceinline__ float dc_load<__nv_bfloat16>(const __nv_bfloat16 *p) { return __bfloat162float(__ldg(p)); }

template <typename T, int kPer>
__global__ void __launch_bounds__(32 * kDcWarps)
    k_depth_context(const T *__restrict__ x, float *__restrict__ depth, float *__restrict__ rows, int D, int C, int HW,
                    long long img_stride) {
  extern __shared__ __align__(16) float dc_tile[];  // [32][C + 1] context tile
  __shared__ float s_red[kDcWarps][32];
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int img = blockIdx.y;
  const int hw0 = blockIdx.x * 32;
  const int n_j = min(32, HW - hw0);
  const bool live = lane < n_j;
  const T *xi = x + (size_t)img * img_stride + hw0 + lane;

  // ---- context planes -> shared tile (issued first: they overlap the softmax arithmetic) -------
  const int pitch = C | 1;
  for (int c0 = warp; c0 < C; c0 += kDcWarps * 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int c = c0 + u * kDcWarps;
      v[u] = (live && c < C) ? dc_load<T>(xi + (size_t)(D + c) * HW) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int c = c0 + u * kDcWarps;
      if (c < C) dc_tile[lane * pitch + c] = v[u];
    }
  }
  // ---- softmax over the depth slice: this thread's bins d = warp, warp + 8, ... -----------------
  float v[kPer];
  float m = -INFINITY;
#pragma unroll
  for (int k = 0; k < kPer; ++k) {
    const int d = warp + k * kDcWarps;
    v[k] = (live && d < D) ? dc_load<T>(xi + (size_t)d * HW) : -INFINITY;
    m = fmaxf(m, v[k]);
  }
  s_red[warp][lane] = m;
  __syncthreads();
#pragma unroll
  for (int w = 0; w < kDcWarps; ++w) m = fmaxf(m, s_red[w][lane]);
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < kPer; ++k) {
    v[k] = expf(v[k] - m);  // (-inf for the bins this thread does not have: exp -> 0)
    sum += v[k];
  }
  __syncthreads();  // the maxima have been read
  s_red[warp][lane] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int w = 0; w < kDcWarps; ++w) sum += s_red[w][lane];
  if (live) {
    float *di = depth + (size_t)img * D * HW + hw0 + lane;
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      const int d = warp + k * kDcWarps;
      if (d < D) st_stream_f32(di + (size_t)d * HW, __fdiv_rn(v[k], sum));
    }
  }
  // ---- context rows: the tile leaves as one contiguous run of n_j * C floats ---------------------
  float *ri = rows + ((size_t)img * HW + hw0) * C;
  for (int i = threadIdx.x; i < n_j * C; i += 32 * kDcWarps) {
    const int j = i / C, c = i - j * C;
    ri[i] = dc_tile[j * pitch + c];
  }
}

// d(x): softmax backward on the depth slice, rows -> planes on the context slice
template <int kPer>
__global__ void __launch_bounds__(32 * kDcWarps)
    k_depth_context_bwd(const float *__restrict__ depth, const float *__restrict__ depth_grad,
                        const float *__restrict__ rows_grad, float *__restrict__ x_grad, int D, int C, int HW) {
  extern __shared__ __align__(16) float dc_tile[];  // [32][C + 1]
  __shared__ float s_red[kDcWarps][32];
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int img = blockIdx.y;
  const int hw0 = blockIdx.x * 32;
  const int n_j = min(32, HW - hw0);
  const bool live = lane < n_j;
  const int pitch = C | 1;
  const float *ri = rows_grad + ((size_t)img * HW + hw0) * C;
  for (int i = threadIdx.x; i < n_j * C; i += 32 * kDcWarps) {
    const int j = i / C, c = i - j * C;
    dc_tile[j * pitch + c] = ld_stream_f32(ri + i);
  }
  const size_t off = (size_t)img * D * HW + hw0 + lane;
  float y[kPer], g[kPer];
  float dot = 0.f;
#pragma unroll
  for (int k = 0; k < kPer; ++k) {
    const int d = warp + k * kDcWarps;
    const bool ok = live && d < D;
    y[k] = ok ? ld_stream_f32(depth + off + (size_t)d * HW) : 0.f;
    g[k] = ok ? ld_stream_f32(depth_grad + off + (size_t)d * HW) : 0.f;
    dot = fmaf(y[k], g[k], dot);
  }
  s_red[warp][lane] = dot;
  __syncthreads();  // also: the context tile is complete
  dot = 0.f;
#pragma unroll
  for (int w = 0; w < kDcWarps; ++w) dot += s_red[w][lane];
  float *xg = x_grad + (size_t)img * (D + C) * HW + hw0 + lane;
  if (live) {
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      const int d = warp + k * kDcWarps;
      if (d < D) st_stream_f32(xg + (size_t)d * HW, y[k] * (g[k] - dot));
    }
    for (int c = warp; c < C; c += kDcWarps) st_stream_f32(xg + (size_t)(D + c) * HW, dc_tile[lane * pitch + c]);
  }
}

}  // namespace rcb

using namespace rcb;

static int dc_args_ok(int n_img, int D, int C, int HW) {
  if (n_img <= 0 || D <= 0 || C <= 0 || HW <= 0) return RCB_ERR_ARG;
  if (D > kDcWarps * kDcPerWarp || n_img > 65535 || (size_t)32 * (C | 1) * 4 > 200 * 1024) return RCB_ERR_UNSUPPORTED;
  return RCB_OK;
}

template <typename T>
static int dc_launch(const void *x, float *depth, float *rows, int n_img, int D, int C, int HW, long long img_stride,
                     cudaStream_t s) {
  const size_t smem = (size_t)32 * (C | 1) * 4;
  dim3 grid(ceil_div(HW, 32), n_img);
  const int per = ceil_div(D, kDcWarps);
  auto go = [&](auto kernel) -> int {
    if (smem > 48 * 1024) RCB_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<grid, 32 * kDcWarps, smem, s>>>(static_cast<const T *>(x), depth, rows, D, C, HW, img_stride);
    RCB_LAUNCH_CHECK();
    return RCB_OK;
  };
  if (per <= 8) return go(k_depth_context<T, 8>);
  if (per <= 16) return go(k_depth_context<T, 16>);
  return go(k_depth_context<T, kDcPerWarp>);
}

extern "C" int rcb_depth_context_split(const void *x, int x_dtype, float *depth, float *context_rows, int n_img, int D,
                                       int C, int HW, long long x_img_stride, int device, rcb_stream_t stream) {
  int rc = dc_args_ok(n_img, D, C, HW);
  if (rc != RCB_OK) return rc;
  if (!x || !depth || !context_rows || x_img_stride < (long long)(D + C) * HW) return RCB_ERR_ARG;
  DeviceGuard guard(device);
  if (guard.err) return guard.err;
  cudaStream_t s = (cudaStream_t)stream;
  switch (x_dtype) {
    case RCB_DTYPE_F32: return dc_launch<float>(x, depth, context_rows, n_img, D, C, HW, x_img_stride, s);
    case RCB_DTYPE_F16: return dc_launch<__half>(x, depth, context_rows, n_img, D, C, HW, x_img_stride, s);
    case RCB_DTYPE_BF16: return dc_launch<__nv_bfloat16>(x, depth, context_rows, n_img, D, C, HW, x_img_stride, s);
    default: return RCB_ERR_ARG;
  }
}

extern "C" int rcb_depth_context_split_bwd(const float *depth, const float *depth_grad, const float *context_rows_grad,
                                           float *x_grad, int n_img, int D, int C, int HW, int device,
                                           rcb_stream_t stream) {
  int rc = dc_args_ok(n_img, D, C, HW);
  if (rc != RCB_OK) return rc;
  if (!depth || !depth_grad || !context_rows_grad || !x_grad) return RCB_ERR_ARG;
  DeviceGuard guard(device);
  if (guard.err) return guard.err;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t smem = (size_t)32 * (C | 1) * 4;
  dim3 grid(ceil_div(HW, 32), n_img);
  const int per = ceil_div(D, kDcWarps);
  auto go = [&](auto kernel) -> int {
    if (smem > 48 * 1024) RCB_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<grid, 32 * kDcWarps, smem, s>>>(depth, depth_grad, context_rows_grad, x_grad, D, C, HW);
    RCB_LAUNCH_CHECK();
    return RCB_OK;
  };
  if (per <= 8) return go(k_depth_context_bwd<8>);
  if (per <= 16) return go(k_depth_context_bwd<16>);
  return go(k_depth_context_bwd<kDcPerWarp>);
}
