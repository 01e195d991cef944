/*
 * CPU oracle for bev_pool_v2 forward / backward.  TEST INFRASTRUCTURE ONLY --
 * never linked into, or called by, the product library (librcbevdet_b200.so).
 *
 * Plain-C restatement of the two kernels of the reference
 *   mmdet3d/ops/bev_pool_v2/src/bev_pool_cuda.cu:21-48   (bev_pool_v2_kernel)
 *   mmdet3d/ops/bev_pool_v2/src/bev_pool_cuda.cu:67-121  (bev_pool_grad_kernel)
 * with the CUDA thread index turned into loops.  nvcc contracts `a += b * c`
 * into one fused multiply-add by default, so every accumulation below is an
 * explicit fmaf() in the same order the reference thread performs it; the file
 * is compiled with -ffp-contract=off so nothing else fuses.
 *
 * Parity status: pinned by the reference's known-answer test
 * (mmdet3d/ops/bev_pool_v2/bev_pool.py:145-176) in tests/test_oracle.py and, on
 * the GPU box, against the reference kernels themselves (oracle/_ref).
 *
 * `threads` > 1 spreads intervals over OpenMP threads (each interval is still
 * summed sequentially, so results do not depend on the thread count).
 */
#include <math.h>
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* out: (n_cells, c) channels-last, pre-zeroed by the caller (bev_pool.py:27). */
void oracle_bev_pool_v2_fwd(int c, int n_intervals, const float *depth, const float *feat,
                            const int *ranks_depth, const int *ranks_feat, const int *ranks_bev,
                            const int *interval_starts, const int *interval_lengths, float *out,
                            int threads) {
  (void)threads;
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads > 0 ? threads : 1)
  for (int index = 0; index < n_intervals; ++index) {        /* .cu:30-33 */
    const int start = interval_starts[index];                /* .cu:34 */
    const int length = interval_lengths[index];              /* .cu:35 */
    float *cur_out = out + (size_t)ranks_bev[start] * c;     /* .cu:45-46 */
    for (int cur_c = 0; cur_c < c; ++cur_c) {
      float psum = 0.f;                                      /* .cu:36 */
      for (int i = 0; i < length; ++i) {                     /* .cu:39-43 */
        const float d = depth[ranks_depth[start + i]];
        const float f = feat[(size_t)ranks_feat[start + i] * c + cur_c];
        psum = fmaf(f, d, psum);
      }
      cur_out[cur_c] = psum;                                 /* .cu:47 */
    }
  }
}

/* Intervals here are runs of equal ranks_feat (bev_pool.py:47-57).
 * depth_grad / feat_grad pre-zeroed by the caller (bev_pool.py:67-68). */
void oracle_bev_pool_v2_bwd(int c, int n_intervals, const float *out_grad, const float *depth,
                            const float *feat, const int *ranks_depth, const int *ranks_feat,
                            const int *ranks_bev, const int *interval_starts,
                            const int *interval_lengths, float *depth_grad, float *feat_grad,
                            int threads) {
  (void)threads;
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads > 0 ? threads : 1)
  for (int idx = 0; idx < n_intervals; ++idx) {              /* .cu:79-80 */
    const int start = interval_starts[idx];
    const int length = interval_lengths[idx];
    for (int i = 0; i < length; ++i) {                       /* .cu:91-105 */
      const float *og = out_grad + (size_t)ranks_bev[start + i] * c;
      const float *ft = feat + (size_t)ranks_feat[start + i] * c;
      float grad_sum = 0.f;
      for (int cur_c = 0; cur_c < c; ++cur_c) grad_sum = fmaf(og[cur_c], ft[cur_c], grad_sum);
      depth_grad[ranks_depth[start + i]] = grad_sum;
    }
    float *fg = feat_grad + (size_t)ranks_feat[start] * c;   /* .cu:107-120 */
    for (int cur_c = 0; cur_c < c; ++cur_c) {
      float grad_sum = 0.f;
      for (int i = 0; i < length; ++i) {
        const float g = out_grad[(size_t)ranks_bev[start + i] * c + cur_c];
        const float d = depth[ranks_depth[start + i]];
        grad_sum = fmaf(g, d, grad_sum);
      }
      fg[cur_c] = grad_sum;
    }
  }
}

/*
 * voxel_pooling_prepare_v2 (mmdet3d/models/necks/view_transformer.py:207-265) as a stable
 * counting sort -- the same result as oracle.py's numpy restatement (argsort kind="stable"),
 * which is the one pinned against the reference's golden vectors; tests/test_oracle.py checks
 * the two against each other.  Exists so that the CPU baseline in bench.py is a compiled,
 * multi-threaded program rather than a numpy script.
 *
 * Outputs have capacity P (ranks) and min(P, B*cells) (intervals); counts = {n_kept, n_intervals}.
 * Returns 0, or -1 if allocation fails / the grid is outside the fp32-exact range (2^24 cells).
 */
#include <stdlib.h>
#include <string.h>

int oracle_voxel_pooling_prepare_v2(int B, int N, int D, int H, int W, const float *coor,
                                    const float *lower, const float *interval, const float *size,
                                    int *ranks_bev, int *ranks_depth, int *ranks_feat,
                                    int *interval_starts, int *interval_lengths, int *counts,
                                    int threads) {
  const long long P = (long long)B * N * D * H * W;
  const long long per_sample = (long long)N * D * H * W;
  const int HW = H * W;
  const long long DHW = (long long)D * HW;
  const float sz0 = size[0], sz1 = size[1], sz2 = size[2];
  const long long cells_ll = (long long)B * (long long)sz0 * (long long)sz1 * (long long)sz2;
  if (cells_ll > (1ll << 24) || cells_ll <= 0) return -1;
  const int n_cells = (int)cells_ll;
  int *cell = (int *)malloc((size_t)(P > 0 ? P : 1) * sizeof(int));
  int *start = (int *)calloc((size_t)n_cells + 1, sizeof(int));
  if (!cell || !start) {
    free(cell);
    free(start);
    return -1;
  }
  (void)threads;
  /* :230-240, 246-249 -- per point: fp32 subtract, fp32 divide, .long() truncation, range test on
   * the truncated value (compared in fp32), rank accumulated in fp32 */
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
  for (long long p = 0; p < P; ++p) {
    const float vx = (coor[3 * p] - lower[0]) / interval[0];
    const float vy = (coor[3 * p + 1] - lower[1]) / interval[1];
    const float vz = (coor[3 * p + 2] - lower[2]) / interval[2];
    /* x86 cvttss2si: NaN and out-of-range give INT64_MIN, as torch's CPU .long() does */
    const long long ix = (vx == vx && vx > -9.2e18f && vx < 9.2e18f) ? (long long)vx : (-9223372036854775807ll - 1);
    const long long iy = (vy == vy && vy > -9.2e18f && vy < 9.2e18f) ? (long long)vy : (-9223372036854775807ll - 1);
    const long long iz = (vz == vz && vz > -9.2e18f && vz < 9.2e18f) ? (long long)vz : (-9223372036854775807ll - 1);
    const int kept = ix >= 0 && (float)ix < sz0 && iy >= 0 && (float)iy < sz1 && iz >= 0 && (float)iz < sz2;
    if (!kept) {
      cell[p] = -1;
      continue;
    }
    const float b = (float)(p / per_sample);
    float r = b * ((sz2 * sz1) * sz0);
    r = r + (float)iz * (sz1 * sz0);
    r = r + ((float)iy * sz0 + (float)ix);
    cell[p] = (int)r;
  }
  /* :250 -- stable sort by rank == counting sort that visits points in ascending order.
   * Samples never share a cell, so they are histogrammed and scattered independently. */
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 0 ? threads : 1)
  for (int b = 0; b < B; ++b)
    for (long long p = b * per_sample; p < (b + 1) * per_sample; ++p)
      if (cell[p] >= 0) start[cell[p] + 1] += 1;
  int n_iv = 0;
  for (int c = 0; c < n_cells; ++c) { /* :254-262 run boundaries */
    const int len = start[c + 1];
    if (len > 0) {
      interval_starts[n_iv] = start[c];
      interval_lengths[n_iv] = len;
      ++n_iv;
    }
    start[c + 1] = start[c] + len;
  }
  counts[0] = start[n_cells];
  counts[1] = n_iv;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 0 ? threads : 1)
  for (int b = 0; b < B; ++b)
    for (long long p = b * per_sample; p < (b + 1) * per_sample; ++p) {
      const int c = cell[p];
      if (c < 0) continue;
      const int slot = start[c]++;
      ranks_bev[slot] = c;
      ranks_depth[slot] = (int)p;                                       /* :223 */
      ranks_feat[slot] = (int)((p / DHW) * HW + (p % DHW) % HW);        /* :225-228 */
    }
  free(cell);
  free(start);
  return 0;
}
