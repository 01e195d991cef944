// SURVEY.md 8(f-3) -- temporal alignment of pooled BEV features:
// BEVDepth4D.gen_grid + shift_feature (mmdet3d/models/detectors/bevdet_rc.py:585-657).
//
// The reference builds an (n, h, w, 3, 1) grid tensor, multiplies it by a per-sample 3x3 matrix with
// a batched matmul, normalises it and hands it to F.grid_sample(align_corners=True), which
// un-normalises it again.  Here one kernel does all of it per output pixel: p = tf[n] * (x, y, 1),
// the normalise / un-normalise round trip in the same fp32 operations (so the sampling positions
// round as the reference's do), then the bilinear taps with zero padding.  The 3x3 matrix (current
// key-ego BEV pixel -> adjacent-frame BEV pixel) is 9 floats per sample, prepared on the host side
// from the calibration (rcbevdet_b200/temporal.py: the handful of 4x4 products of :604-641).
//
// A thread owns one output pixel and a slice of the channels: the four taps' addresses and weights
// are computed once, the channel loop is four loads + four FMAs per channel, neighbouring threads
// read neighbouring input pixels.  The backward (gradient w.r.t. the feature map only; the grid
// comes from poses) scatters with float atomics, as the reference's grid_sample backward does.
#include "common.cuh"

namespace rcb {

struct ShiftTaps {
  int o00, o01, o10, o11;  // offsets inside a channel plane (valid where the weight is non-zero)
  float w00, w01, w10, w11;
};

// bevdet_rc.py:644-650 then grid_sample's align_corners=True un-normalisation, bilinear weights
// with zero padding (the same operations in the same order, fp32, no contraction)
__device__ __forceinline__ ShiftTaps shift_taps(const float *__restrict__ tf, int x, int y, int H, int W) {
  const float fx = (float)x, fy = (float)y;
  const float px = __fadd_rn(__fadd_rn(__fmul_rn(tf[0], fx), __fmul_rn(tf[1], fy)), tf[2]);
  const float py = __fadd_rn(__fadd_rn(__fmul_rn(tf[3], fx), __fmul_rn(tf[4], fy)), tf[5]);
  const float gx = __fsub_rn(__fmul_rn(__fdiv_rn(px, (float)W - 1.0f), 2.0f), 1.0f);
  const float gy = __fsub_rn(__fmul_rn(__fdiv_rn(py, (float)H - 1.0f), 2.0f), 1.0f);
  const float ix = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.0f), 2.0f), (float)(W - 1));
  const float iy = __fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.0f), 2.0f), (float)(H - 1));
  const float x0f = floorf(ix), y0f = floorf(iy);
  const float x1f = x0f + 1.0f, y1f = y0f + 1.0f;
  ShiftTaps t;
  t.w00 = __fmul_rn(__fsub_rn(x1f, ix), __fsub_rn(y1f, iy));  // nw
  t.w01 = __fmul_rn(__fsub_rn(ix, x0f), __fsub_rn(y1f, iy));  // ne
  t.w10 = __fmul_rn(__fsub_rn(x1f, ix), __fsub_rn(iy, y0f));  // sw
  t.w11 = __fmul_rn(__fsub_rn(ix, x0f), __fsub_rn(iy, y0f));  // se
  // NaN / huge positions sample nothing
  const bool finite = fabsf(ix) < 1e9f && fabsf(iy) < 1e9f;
  const int x0 = finite ? (int)x0f : -2, y0 = finite ? (int)y0f : -2;
  const bool vx0 = x0 >= 0 && x0 < W, vx1 = x0 + 1 >= 0 && x0 + 1 < W;
  const bool vy0 = y0 >= 0 && y0 < H, vy1 = y0 + 1 >= 0 && y0 + 1 < H;
  if (!(vx0 && vy0)) t.w00 = 0.f;
  if (!(vx1 && vy0)) t.w01 = 0.f;
  if (!(vx0 && vy1)) t.w10 = 0.f;
  if (!(vx1 && vy1)) t.w11 = 0.f;
  const int xc0 = min(max(x0, 0), W - 1), xc1 = min(max(x0 + 1, 0), W - 1);
  const int yc0 = min(max(y0, 0), H - 1), yc1 = min(max(y0 + 1, 0), H - 1);
  t.o00 = yc0 * W + xc0, t.o01 = yc0 * W + xc1, t.o10 = yc1 * W + xc0, t.o11 = yc1 * W + xc1;
  return t;
}

constexpr int kShiftChannels = 8;  // channels per thread and trip: 32 loads in flight

__global__ void __launch_bounds__(256)
    k_bev_shift(const float *__restrict__ in, const float *__restrict__ tf, float *__restrict__ out, int C, int H, int W,
                int c_per_block) {
  const int n = blockIdx.z;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= H * W) return;
  const int y = pix / W, x = pix - y * W;
  const ShiftTaps t = shift_taps(tf + n * 9, x, y, H, W);
  const size_t plane = (size_t)H * W;
  const int c0 = blockIdx.y * c_per_block, c1 = min(C, c0 + c_per_block);
  const float *src = in + ((size_t)n * C + c0) * plane;
  float *dst = out + ((size_t)n * C + c0) * plane + pix;
  for (int c = c0; c < c1; c += kShiftChannels) {
    float a[kShiftChannels], b[kShiftChannels], d[kShiftChannels], e[kShiftChannels];
#pragma unroll
    for (int u = 0; u < kShiftChannels; ++u) {
      const float *p = src + (size_t)u * plane;
      const bool ok = c + u < c1;
      a[u] = ok ? __ldg(p + t.o00) : 0.f, b[u] = ok ? __ldg(p + t.o01) : 0.f;
      d[u] = ok ? __ldg(p + t.o10) : 0.f, e[u] = ok ? __ldg(p + t.o11) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < kShiftChannels; ++u) {
      if (c + u < c1) {
        float v = __fmul_rn(a[u], t.w00);  // nw, ne, sw, se: the reference's accumulation order
        v = fmaf(b[u], t.w01, v);
        v = fmaf(d[u], t.w10, v);
        v = fmaf(e[u], t.w11, v);
        st_stream_f32(dst + (size_t)u * plane, v);
      }
    }
    src += (size_t)kShiftChannels * plane;
    dst += (size_t)kShiftChannels * plane;
  }
}

__global__ void __launch_bounds__(256)
    k_bev_shift_bwd(const float *__restrict__ out_grad, const float *__restrict__ tf, float *__restrict__ in_grad, int C,
                    int H, int W, int c_per_block) {
  const int n = blockIdx.z;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= H * W) return;
  const int y = pix / W, x = pix - y * W;
  const ShiftTaps t = shift_taps(tf + n * 9, x, y, H, W);
  const size_t plane = (size_t)H * W;
  const int c0 = blockIdx.y * c_per_block, c1 = min(C, c0 + c_per_block);
  for (int c = c0; c < c1; ++c) {
    const float g = __ldg(out_grad + ((size_t)n * C + c) * plane + pix);
    float *dst = in_grad + ((size_t)n * C + c) * plane;
    if (t.w00 != 0.f) atomicAdd(dst + t.o00, g * t.w00);
    if (t.w01 != 0.f) atomicAdd(dst + t.o01, g * t.w01);
    if (t.w10 != 0.f) atomicAdd(dst + t.o10, g * t.w10);
    if (t.w11 != 0.f) atomicAdd(dst + t.o11, g * t.w11);
  }
}

}  // namespace rcb

using namespace rcb;

static int shift_args_ok(const void *a, const void *tf, const void *b, int n, int C, int H, int W) {
  if (!a || !tf || !b) return RCB_ERR_ARG;
  if (n <= 0 || C <= 0 || H <= 1 || W <= 1) return RCB_ERR_ARG;  // (w - 1), (h - 1) normalise the grid
  if (n > 65535 || (long long)n * C * H * W >= (1ll << 40)) return RCB_ERR_UNSUPPORTED;
  return RCB_OK;
}

extern "C" int rcb_bev_shift_feature(const float *input, const float *tf, float *output, int n, int C, int H, int W,
                                     int device, rcb_stream_t stream) {
  int rc = shift_args_ok(input, tf, output, n, C, H, W);
  if (rc != RCB_OK) return rc;
  DeviceGuard guard(device);
  if (guard.err) return guard.err;
  const int c_per_block = C >= 32 ? (C + 3) / 4 : C;  // four channel slices per pixel block: enough CTAs at 128x128
  dim3 grid(ceil_div(H * W, 256), ceil_div(C, c_per_block), n);
  k_bev_shift<<<grid, 256, 0, (cudaStream_t)stream>>>(input, tf, output, C, H, W, c_per_block);
  RCB_LAUNCH_CHECK();
  return RCB_OK;
}

extern "C" int rcb_bev_shift_feature_bwd(const float *output_grad, const float *tf, float *input_grad, int n, int C,
                                         int H, int W, int device, rcb_stream_t stream) {
  int rc = shift_args_ok(output_grad, tf, input_grad, n, C, H, W);
  if (rc != RCB_OK) return rc;
  DeviceGuard guard(device);
  if (guard.err) return guard.err;
  cudaStream_t s = (cudaStream_t)stream;
  RCB_CUDA_TRY(cudaMemsetAsync(input_grad, 0, (size_t)n * C * H * W * 4, s));
  const int c_per_block = C >= 32 ? (C + 3) / 4 : C;
  dim3 grid(ceil_div(H * W, 256), ceil_div(C, c_per_block), n);
  k_bev_shift_bwd<<<grid, 256, 0, s>>>(output_grad, tf, input_grad, C, H, W, c_per_block);
  RCB_LAUNCH_CHECK();
  return RCB_OK;
}
