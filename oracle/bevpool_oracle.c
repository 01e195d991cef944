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
