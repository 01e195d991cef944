// Row R -- PointPillarsScatterRCS.forward up to (not including) its two convolutions.
// Reference: mmdet3d/middle_encoders/pillar_scatter.py:64-104 (pillar scatter), :115-131 (RCS
// heat-maps: a Python loop over pillars with two .item() syncs and a numpy Gaussian each),
// mmdet3d/core/utils/gaussian.py:6-23 (gaussian_2d, float64), :26-55 (max-draw), :57-81 (fill).
//
// The sequential loop has two order-dependent effects, both expressible as order-free maxima:
//   heatmap[cell]      = max over pillars covering the cell of G_r(dx, dy)        (max is commutative)
//   heatmap_feat[cell] = rcs value of the LAST pillar (largest index) covering the cell
// so the loop becomes one splat kernel with integer atomicMax (non-negative floats order like
// their bit patterns) -- exact and deterministic -- followed by one dense write kernel that
// produces all three outputs, zeros included, with coalesced rows (no memset of the outputs,
// no per-pillar host sync).
#include "common.cuh"

namespace rcb {

struct RadarParams {
  int V, Cin, rcs_dim, B, ny, nx, cells;  // cells = ny*nx
};

// pillar_scatter.py:122-126,130: r = x^2 + y^2; relu(rcs * r) + 1 in fp32, then Python int().
__device__ __forceinline__ int rcs_radius(const float *rcs_row, int rcs_dim) {
  const float x = rcs_row[0], y = rcs_row[1];
  const float r = __fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y));
  const float t = fmaxf(__fmul_rn(rcs_row[rcs_dim - 2], r), 0.f);
  const float rad = __fadd_rn(t, 1.f);
  if (!(rad < 2147483520.f)) return 0x7fffff00;
  return (int)rad;
}

__global__ void __launch_bounds__(256) k_radar_splat(RadarParams p, const float *__restrict__ rcs,
                                                     const int *__restrict__ coors,
                                                     int *__restrict__ pillar_at,
                                                     int *__restrict__ last_cover,
                                                     int *__restrict__ heat_bits) {
  const int lane = lane_id();
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; v < p.V; v += warps) {
    const int b = __ldg(coors + v * 4), y = __ldg(coors + v * 4 + 2), x = __ldg(coors + v * 4 + 3);
    if (b < 0 || b >= p.B || y < 0 || y >= p.ny || x < 0 || x >= p.nx) continue;
    const int base = b * p.cells;
    if (lane == 0) atomicMax(pillar_at + base + y * p.nx + x, v);
    const int radius = rcs_radius(rcs + (size_t)v * p.rcs_dim, p.rcs_dim);
    // gaussian.py:40-47: clipped window
    const int left = min(x, radius), right = min(p.nx - x, radius + 1);
    const int top = min(y, radius), bottom = min(p.ny - y, radius + 1);
    const int w = left + right, h = top + bottom;
    // gaussian.py:17-23 with sigma = diameter / 6 (gaussian.py:38-39), all float64
    const double diameter = 2.0 * (double)radius + 1.0;
    const double sigma = diameter / 6.0;
    const double denom = 2.0 * sigma * sigma;
    const double eps = 2.220446049250313e-16;  // np.finfo(float64).eps * h.max(), h.max() == 1
    for (int i = lane; i < w * h; i += 32) {
      const int iy = i / w, ix = i - iy * w;
      const int dy = iy - top, dx = ix - left;
      double g = exp(-(double)(dx * dx + dy * dy) / denom);
      if (g < eps) g = 0.0;
      const int cell = base + (y + dy) * p.nx + (x + dx);
      atomicMax(heat_bits + cell, __float_as_int((float)g));
      atomicMax(last_cover + cell, v);
    }
  }
}

// 32 consecutive cells x all channels per CTA; pillar rows are read coalesced, transposed in
// shared memory, written as 128-byte channel rows.
__global__ void __launch_bounds__(256)
    k_radar_write(RadarParams p, const float *__restrict__ point_features, const float *__restrict__ rcs,
                  const int *__restrict__ pillar_at, const int *__restrict__ last_cover,
                  const int *__restrict__ heat_bits, float *__restrict__ features,
                  float *__restrict__ heatmap, float *__restrict__ heatmap_feat) {
  extern __shared__ float tile[];  // [Cin][33]
  __shared__ int s_owner[32];
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int tiles_per_sample = ceil_div(p.cells, 32);
  const int b = blockIdx.x / tiles_per_sample;
  const int cell0 = (blockIdx.x - b * tiles_per_sample) * 32;
  const int n = min(32, p.cells - cell0);
  if (threadIdx.x < 32) {
    int o = -1;
    if (threadIdx.x < n) {
      const int g = b * p.cells + cell0 + threadIdx.x;
      o = pillar_at[g];
      heatmap[g] = __int_as_float(heat_bits[g]);
      const int lc = last_cover[g];
      heatmap_feat[g] = lc >= 0 ? __ldg(rcs + (size_t)lc * p.rcs_dim + p.rcs_dim - 2) : 0.f;
    }
    s_owner[threadIdx.x] = o;
  }
  __syncthreads();
  for (int j = warp; j < 32; j += 8) {
    const int o = s_owner[j];
    for (int c = lane; c < p.Cin; c += 32)
      tile[c * 33 + j] = o >= 0 ? __ldg(point_features + (size_t)o * p.Cin + c) : 0.f;
  }
  __syncthreads();
  float *dst = features + (size_t)b * p.Cin * p.cells + cell0;
  for (int c = warp; c < p.Cin; c += 8)
    if (lane < n) st_stream_f32(dst + (size_t)c * p.cells + lane, tile[c * 33 + lane]);
}

__global__ void __launch_bounds__(256)
    k_radar_scatter_bwd(RadarParams p, const float *__restrict__ features_grad,
                        const int *__restrict__ coors, float *__restrict__ point_features_grad) {
  const int lane = lane_id();
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; v < p.V; v += warps) {
    const int b = __ldg(coors + v * 4), y = __ldg(coors + v * 4 + 2), x = __ldg(coors + v * 4 + 3);
    const bool ok = b >= 0 && b < p.B && y >= 0 && y < p.ny && x >= 0 && x < p.nx;
    for (int c = lane; c < p.Cin; c += 32)
      point_features_grad[(size_t)v * p.Cin + c] =
          ok ? __ldg(features_grad + ((size_t)b * p.Cin + c) * p.cells + y * p.nx + x) : 0.f;
  }
}

static int fill_radar(const rcb_radar_desc *d, RadarParams *p) {
  if (!d) return RCB_ERR_ARG;
  if (d->V < 0 || d->Cin <= 0 || d->rcs_dim < 3 || d->B <= 0 || d->ny <= 0 || d->nx <= 0) return RCB_ERR_ARG;
  if ((long long)d->B * d->ny * d->nx * d->Cin >= (1ll << 40)) return RCB_ERR_UNSUPPORTED;
  if ((long long)d->B * d->ny * d->nx >= (1ll << 31)) return RCB_ERR_UNSUPPORTED;
  p->V = d->V, p->Cin = d->Cin, p->rcs_dim = d->rcs_dim, p->B = d->B, p->ny = d->ny, p->nx = d->nx;
  p->cells = d->ny * d->nx;
  return RCB_OK;
}

}  // namespace rcb

using namespace rcb;

extern "C" size_t rcb_radar_workspace_bytes(const rcb_radar_desc *d) {
  RadarParams p;
  if (fill_radar(d, &p) != RCB_OK) return 0;
  return 3 * align_up((size_t)p.B * p.cells * 4, 256);
}

extern "C" int rcb_radar_rcs_scatter(const rcb_radar_desc *d, const float *point_features,
                                     const float *rcs, const int *coors, float *features,
                                     float *heatmap, float *heatmap_feat, void *workspace,
                                     size_t workspace_bytes, int device, rcb_stream_t stream) {
  RadarParams p;
  int rc = fill_radar(d, &p);
  if (rc != RCB_OK) return rc;
  if (!features || !heatmap || !heatmap_feat || !workspace) return RCB_ERR_ARG;
  if (p.V > 0 && (!point_features || !rcs || !coors)) return RCB_ERR_ARG;
  const size_t plane = align_up((size_t)p.B * p.cells * 4, 256);
  if (workspace_bytes < 3 * plane) return RCB_ERR_WORKSPACE;
  DeviceGuard guard(device);
  if (guard.err) return guard.err;
  cudaStream_t s = (cudaStream_t)stream;
  char *ws = static_cast<char *>(workspace);
  int *pillar_at = (int *)ws, *last_cover = (int *)(ws + plane), *heat_bits = (int *)(ws + 2 * plane);
  RCB_CUDA_TRY(cudaMemsetAsync(ws, 0xff, 2 * plane, s));
  RCB_CUDA_TRY(cudaMemsetAsync(heat_bits, 0, plane, s));
  const int sms = sm_count_cached(device);
  if (p.V > 0) {
    k_radar_splat<<<max(1, min(ceil_div(p.V, 8), sms * 32)), 256, 0, s>>>(p, rcs, coors, pillar_at,
                                                                          last_cover, heat_bits);
    RCB_LAUNCH_CHECK();
  }
  const size_t smem = (size_t)p.Cin * 33 * 4;
  if (smem > 200 * 1024) return RCB_ERR_UNSUPPORTED;
  if (smem > 48 * 1024)
    RCB_CUDA_TRY(cudaFuncSetAttribute(k_radar_write, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_radar_write<<<p.B * ceil_div(p.cells, 32), 256, smem, s>>>(p, point_features, rcs, pillar_at,
                                                               last_cover, heat_bits, features, heatmap,
                                                               heatmap_feat);
  RCB_LAUNCH_CHECK();
  return RCB_OK;
}

extern "C" int rcb_radar_scatter_bwd(const rcb_radar_desc *d, const float *features_grad,
                                     const int *coors, float *point_features_grad, int device,
                                     rcb_stream_t stream) {
  RadarParams p;
  int rc = fill_radar(d, &p);
  if (rc != RCB_OK) return rc;
  if (p.V == 0) return RCB_OK;
  if (!features_grad || !coors || !point_features_grad) return RCB_ERR_ARG;
  DeviceGuard guard(device);
  if (guard.err) return guard.err;
  const int sms = sm_count_cached(device);
  k_radar_scatter_bwd<<<max(1, min(ceil_div(p.V, 8), sms * 32)), 256, 0, (cudaStream_t)stream>>>(
      p, features_grad, coors, point_features_grad);
  RCB_LAUNCH_CHECK();
  return RCB_OK;
}
