// Row F -- bev_pool_v2 forward: out[cell, c] = sum_i depth[ranks_depth[i]] * feat[ranks_feat[i], c]
// over the points i of the cell's interval.
// Reference: mmdet3d/ops/bev_pool_v2/src/bev_pool_cuda.cu:21-48 (kernel), :125-131 (launch),
// mmdet3d/ops/bev_pool_v2/bev_pool.py:16-41,86-92 (zero-fill before, permute copy after).
//
//  k_fwd_cells (pool_fwd_cells.cu) is the product path: it needs the dense per-cell CSR `cell_start`,
//      i.e. intervals sorted by cell and tiling [0, K) -- what prepare emits and what
//      rcb_pool_validate proves.  Every cell is written, zeros included, directly in the layout the
//      caller wants: no memset pass, no permute pass.
//
//  k_pool_fwd_intervals (here) is the general path for any ranks the reference accepts: one warp
//      per interval, writes only non-empty cells into a pre-zeroed `out`.
#include "common.cuh"

namespace rcb {

// ---------------------------------------------------------------------------------------------
// General path: the reference's contract verbatim (any ranks, only interval cells written).
// ---------------------------------------------------------------------------------------------
template <typename FeatT>
__global__ void __launch_bounds__(256)
    k_pool_fwd_intervals(int n_intervals, int C, int cells_per_sample, int layout,
                         const float *__restrict__ depth, const FeatT *__restrict__ feat,
                         const int *__restrict__ ranks_depth, const int *__restrict__ ranks_feat,
                         const int *__restrict__ ranks_bev, const int *__restrict__ interval_starts,
                         const int *__restrict__ interval_lengths, float *__restrict__ out) {
  const int lane = lane_id();
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int iv = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; iv < n_intervals; iv += warps) {
    const int start = __ldg(interval_starts + iv), len = __ldg(interval_lengths + iv);
    const int cell = __ldg(ranks_bev + start);
    for (int c0 = 0; c0 < C; c0 += 32) {
      const int c = c0 + lane;
      float acc = 0.f;
      for (int base = 0; base < len; base += 32) {
        const int m = min(32, len - base);
        int row = 0;
        float w = 0.f;
        if (lane < m) {
          row = __ldg(ranks_feat + start + base + lane);
          w = __ldg(depth + __ldg(ranks_depth + start + base + lane));
        }
        for (int k = 0; k < m; ++k) {
          const int rk = __shfl_sync(kFull, row, k);
          const float wk = __shfl_sync(kFull, w, k);
          if (c < C) acc = fmaf(to_f32<FeatT>(feat[(size_t)rk * C + c]), wk, acc);
        }
      }
      if (c < C) {
        if (layout == RCB_LAYOUT_B_C_CELLS) {
          const int bb = cell / cells_per_sample, cc = cell - bb * cells_per_sample;
          out[((size_t)bb * C + c) * cells_per_sample + cc] = acc;
        } else {
          out[(size_t)cell * C + c] = acc;
        }
      }
    }
  }
}

template <typename FeatT>
static int launch_intervals(const rcb_pool_desc *d, const float *depth, const void *feat,
                            const int *rd, const int *rf, const int *rb, const int *starts,
                            const int *lengths, float *out, int sms, cudaStream_t s) {
  const int grid = max(1, min(ceil_div(d->n_intervals, 8), sms * 16));
  k_pool_fwd_intervals<FeatT><<<grid, 256, 0, s>>>(d->n_intervals, d->C, d->Z * d->Y * d->X, d->layout,
                                                   depth, static_cast<const FeatT *>(feat), rd, rf, rb,
                                                   starts, lengths, out);
  RCB_LAUNCH_CHECK();
  return RCB_OK;
}

bool fwd_cells_eligible(const rcb_pool_desc *d, const void *feat, const int *cell_start);
int fwd_cells_launch(const rcb_pool_desc *d, const float *depth, const void *feat, const int *ranks_depth,
                     const int *ranks_feat, const int *cell_start, float *out, cudaStream_t s);

int check_pool_desc(const rcb_pool_desc *d) {
  if (!d) return RCB_ERR_ARG;
  if (d->n_points < 0 || d->n_intervals < 0 || d->C <= 0 || d->B <= 0 || d->Z <= 0 || d->Y <= 0 ||
      d->X <= 0 || d->n_depth < 0 || d->n_pixels < 0)
    return RCB_ERR_ARG;
  if ((long long)d->B * d->Z * d->Y * d->X >= (1ll << 31) / 4) return RCB_ERR_UNSUPPORTED;
  if (d->layout != RCB_LAYOUT_CELLS_C && d->layout != RCB_LAYOUT_B_C_CELLS) return RCB_ERR_ARG;
  if (d->feat_dtype != RCB_DTYPE_F32 && d->feat_dtype != RCB_DTYPE_BF16 && d->feat_dtype != RCB_DTYPE_F16)
    return RCB_ERR_ARG;
  return RCB_OK;
}

}  // namespace rcb

using namespace rcb;

extern "C" int rcb_bev_pool_v2_fwd(const rcb_pool_desc *d, const float *depth, const void *feat,
                                   const int *ranks_depth, const int *ranks_feat,
                                   const int *ranks_bev, const int *interval_lengths,
                                   const int *interval_starts, const int *cell_start, float *out,
                                   int device, rcb_stream_t stream) {
  int rc = check_pool_desc(d);
  if (rc != RCB_OK) return rc;
  if (!out) return RCB_ERR_ARG;
  DeviceGuard guard(device);
  if (guard.err) return guard.err;
  cudaStream_t s = (cudaStream_t)stream;
  const int cps = d->Z * d->Y * d->X;
  const size_t out_bytes = (size_t)d->B * cps * d->C * 4;

  if (fwd_cells_eligible(d, feat, cell_start)) {
    if (d->n_points > 0 && (!depth || !feat || !ranks_depth || !ranks_feat)) return RCB_ERR_ARG;
    return fwd_cells_launch(d, depth, feat, ranks_depth, ranks_feat, cell_start, out, s);
  }
  // A CSR without interval arrays is the sync-free fused chain (n_points is only an upper bound
  // there): it cannot fall back to the interval kernel, so say so instead of returning zeros.
  if (cell_start != nullptr && d->n_points > 0 && (!interval_lengths || !interval_starts)) return RCB_ERR_UNSUPPORTED;
  if (launch_gate()) return RCB_ERR_UNSUPPORTED;  // the interval kernel takes no launch gate
  // general path: zero-fill (bev_pool.py:27) then one warp per interval
  RCB_CUDA_TRY(cudaMemsetAsync(out, 0, out_bytes, s));
  if (d->n_intervals == 0) return RCB_OK;
  if (!depth || !feat || !ranks_depth || !ranks_feat || !ranks_bev || !interval_lengths ||
      !interval_starts)
    return RCB_ERR_ARG;
  const int sms = sm_count_cached(device);
  switch (d->feat_dtype) {
    case RCB_DTYPE_F32:
      return launch_intervals<float>(d, depth, feat, ranks_depth, ranks_feat, ranks_bev,
                                     interval_starts, interval_lengths, out, sms, s);
    case RCB_DTYPE_BF16:
      return launch_intervals<__nv_bfloat16>(d, depth, feat, ranks_depth, ranks_feat, ranks_bev,
                                             interval_starts, interval_lengths, out, sms, s);
    default:
      return launch_intervals<__half>(d, depth, feat, ranks_depth, ranks_feat, ranks_bev,
                                      interval_starts, interval_lengths, out, sms, s);
  }
}

