// Plan checks for caller-supplied ranks (rows F/B/L).  The reference's extension trusts its
// inputs completely (mmdet3d/ops/bev_pool_v2/src/bev_pool.cpp:30-57: raw data_ptr, no checks).
// The fast kernels here rely on properties that hold for everything voxel_pooling_prepare_v2
// produces (view_transformer.py:250-262) but that the op's signature does not promise, so for
// ranks that did not come from this library's prepare they are verified on the device first:
//
//   RANGES_OK     every rank indexes inside its tensor
//   INTERVALS_OK  intervals tile [0, n_points) contiguously, every length > 0, and every point of
//                 an interval carries the interval's cell in ranks_bev
//   SORTED_CELLS  interval cells strictly increasing  -> dense CSR `cell_start`, tile kernel
//   STRUCTURED    ranks_depth unique, ranks_feat == pixel_of(ranks_depth) -> sort-free backward
#include "common.cuh"

namespace rcb {

__global__ void __launch_bounds__(256)
    k_validate_points(int n_points, int n_depth, int n_pixels, int n_cells, int D, int HW,
                      const int *__restrict__ ranks_depth, const int *__restrict__ ranks_feat,
                      const int *__restrict__ ranks_bev, int *__restrict__ point_cell,
                      int *__restrict__ bad) {
  const int stride = gridDim.x * blockDim.x;
  int bad_local = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_points; i += stride) {
    const int rd = ranks_depth[i], rf = ranks_feat[i], rb = ranks_bev[i];
    const bool in_range = rd >= 0 && rd < n_depth && rf >= 0 && rf < n_pixels && rb >= 0 && rb < n_cells;
    if (!in_range) {
      bad_local |= RCB_PLAN_RANGES_OK | RCB_PLAN_STRUCTURED;
      continue;
    }
    if (D > 0 && HW > 0) {
      const int DHW = D * HW;
      const int bn = rd / DHW;
      if (rf != bn * HW + (rd - bn * DHW) % HW) bad_local |= RCB_PLAN_STRUCTURED;
      // uniqueness of ranks_depth: first writer claims the slot
      if (atomicCAS(point_cell + rd, -1, rb) != -1) bad_local |= RCB_PLAN_STRUCTURED;
    } else {
      bad_local |= RCB_PLAN_STRUCTURED;
    }
  }
  if (bad_local) atomicOr(bad, bad_local);
}

__global__ void __launch_bounds__(256)
    k_validate_intervals(int n_points, int n_intervals, int n_cells, const int *__restrict__ ranks_bev,
                         const int *__restrict__ starts, const int *__restrict__ lengths,
                         int *__restrict__ bad) {
  const int stride = gridDim.x * blockDim.x;
  int bad_local = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_intervals; i += stride) {
    const int s = starts[i], l = lengths[i];
    const int expect_end = (i + 1 < n_intervals) ? starts[i + 1] : n_points;
    // l <= n_points - s first: it bounds the scan below even when starts[i + 1] is garbage, and keeps
    // s + l from overflowing
    if (l <= 0 || s < 0 || s >= n_points || l > n_points - s || (i == 0 && s != 0) || s + l != expect_end) {
      bad_local |= RCB_PLAN_INTERVALS_OK | RCB_PLAN_SORTED_CELLS;
      continue;
    }
    const int c = ranks_bev[s];
    if (c < 0 || c >= n_cells) {
      bad_local |= RCB_PLAN_RANGES_OK | RCB_PLAN_SORTED_CELLS;
      continue;
    }
    for (int k = 1; k < l; ++k)
      if (ranks_bev[s + k] != c) {
        bad_local |= RCB_PLAN_INTERVALS_OK;
        break;
      }
    if (i > 0) {
      const int sp = starts[i - 1];
      if (sp < 0 || sp >= n_points || ranks_bev[sp] >= c) bad_local |= RCB_PLAN_SORTED_CELLS;
    }
  }
  if (bad_local) atomicOr(bad, bad_local);
}

__global__ void k_finish_flags(int n_points, int n_intervals, const int *bad, int *flags_out) {
  int f = RCB_PLAN_ALL & ~(*bad);
  if ((n_points == 0) != (n_intervals == 0)) f &= ~(RCB_PLAN_INTERVALS_OK | RCB_PLAN_SORTED_CELLS);
  // the fast kernels need everything below them in the chain
  if (!(f & RCB_PLAN_RANGES_OK)) f = 0;
  if (!(f & RCB_PLAN_INTERVALS_OK)) f &= ~RCB_PLAN_SORTED_CELLS;
  *flags_out = f;
}

// cell_start[c] for every cell from sorted intervals: interval i fills (cell_{i-1}, cell_i].
__global__ void __launch_bounds__(256)
    k_build_cellmap(int n_points, int n_intervals, int n_cells, const int *__restrict__ ranks_bev,
                    const int *__restrict__ starts, int *__restrict__ cell_start) {
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= n_intervals; i += stride) {
    const int prev = i > 0 ? ranks_bev[starts[i - 1]] : -1;
    const int cur = i < n_intervals ? ranks_bev[starts[i]] : n_cells;
    const int val = i < n_intervals ? starts[i] : n_points;
    for (int c = prev + 1; c <= cur; ++c) cell_start[c] = val;
  }
}

}  // namespace rcb

using namespace rcb;

extern "C" size_t rcb_pool_validate_workspace_bytes(const rcb_pool_desc *d) {
  (void)d;
  return 256;
}

extern "C" int rcb_pool_validate(const rcb_pool_desc *d, const int *ranks_depth, const int *ranks_feat,
                                 const int *ranks_bev, const int *interval_starts,
                                 const int *interval_lengths, int *point_cell, int *flags_out,
                                 void *workspace, size_t workspace_bytes, int device,
                                 rcb_stream_t stream) {
  int rc = check_pool_desc(d);
  if (rc != RCB_OK) return rc;
  if (!flags_out || !workspace || !point_cell) return RCB_ERR_ARG;
  if (workspace_bytes < 4) return RCB_ERR_WORKSPACE;
  if (d->n_points > 0 && (!ranks_depth || !ranks_feat || !ranks_bev)) return RCB_ERR_ARG;
  if (d->n_intervals > 0 && (!interval_starts || !interval_lengths)) return RCB_ERR_ARG;
  DeviceGuard guard(device);
  if (guard.err) return guard.err;
  cudaStream_t s = (cudaStream_t)stream;
  int *bad = static_cast<int *>(workspace);
  const int n_cells = d->B * d->Z * d->Y * d->X;
  const int sms = sm_count_cached(device);
  RCB_CUDA_TRY(cudaMemsetAsync(bad, 0, 4, s));
  RCB_CUDA_TRY(cudaMemsetAsync(point_cell, 0xff, (size_t)d->n_depth * 4, s));
  if (d->n_points > 0) {
    k_validate_points<<<min(ceil_div(d->n_points, 256), sms * 16), 256, 0, s>>>(
        d->n_points, d->n_depth, d->n_pixels, n_cells, d->D, d->HW, ranks_depth, ranks_feat, ranks_bev,
        point_cell, bad);
    RCB_LAUNCH_CHECK();
  }
  if (d->n_intervals > 0 && d->n_points > 0) {
    k_validate_intervals<<<min(ceil_div(d->n_intervals, 256), sms * 16), 256, 0, s>>>(
        d->n_points, d->n_intervals, n_cells, ranks_bev, interval_starts, interval_lengths, bad);
    RCB_LAUNCH_CHECK();
  }
  k_finish_flags<<<1, 1, 0, s>>>(d->n_points, d->n_intervals, bad, flags_out);
  RCB_LAUNCH_CHECK();
  return RCB_OK;
}

extern "C" int rcb_pool_build_cellmap(const rcb_pool_desc *d, const int *ranks_bev,
                                      const int *interval_starts, int *cell_start, int device,
                                      rcb_stream_t stream) {
  int rc = check_pool_desc(d);
  if (rc != RCB_OK) return rc;
  if (!cell_start) return RCB_ERR_ARG;
  if (d->n_intervals > 0 && (!ranks_bev || !interval_starts)) return RCB_ERR_ARG;
  DeviceGuard guard(device);
  if (guard.err) return guard.err;
  const int n_cells = d->B * d->Z * d->Y * d->X;
  const int sms = sm_count_cached(device);
  k_build_cellmap<<<max(1, min(ceil_div(d->n_intervals + 1, 256), sms * 16)), 256, 0,
                    (cudaStream_t)stream>>>(d->n_points, d->n_intervals, n_cells, ranks_bev,
                                            interval_starts, cell_start);
  RCB_LAUNCH_CHECK();
  return RCB_OK;
}
