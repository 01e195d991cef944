/*
 * rcbevdet_b200 -- C ABI of the B200-native (sm_100a) camera->BEV pooling path.
 *
 * This is the drop-in boundary.  Every entry point is `extern "C"`, takes plain device
 * pointers (what torch's `tensor.data_ptr()` returns), sizes and a CUDA stream; there are
 * no torch types, no globals and no allocation inside (the caller owns every buffer,
 * workspace included).  All calls are re-entrant and asynchronous on `stream`; they set
 * the CUDA device from the `device` argument themselves, so they can be called from the
 * autograd engine's worker thread (the reference relies on OptionalCUDAGuard for that:
 * mmdet3d/ops/bev_pool_v2/src/bev_pool.cpp:42,88).
 *
 * Return value: 0 on success, a negative RCB_ERR_* for argument errors, or a positive
 * `cudaError_t`.  (The reference's entry points are `void` and unchecked.)
 *
 * Which reference interface each entry point replaces is cited on the declaration.
 */
#ifndef RCBEVDET_B200_H_
#define RCBEVDET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st *rcb_stream_t; /* == cudaStream_t */

#define RCB_OK 0
#define RCB_ERR_ARG (-1)         /* null pointer / negative size / inconsistent shape  */
#define RCB_ERR_WORKSPACE (-2)   /* workspace too small                                */
#define RCB_ERR_UNSUPPORTED (-3) /* size outside the supported envelope */
#define RCB_ERR_ALIGN (-4)       /* pointer not aligned as required                    */

#define RCB_DTYPE_F32 0
#define RCB_DTYPE_BF16 1
#define RCB_DTYPE_F16 2

/* Output / gradient layouts of the pooled BEV tensor */
#define RCB_LAYOUT_CELLS_C 0 /* (B*Z*Y*X, C) channels last: what bev_pool_v2_forward writes   */
#define RCB_LAYOUT_B_C_CELLS 1 /* (B, C, Z*Y*X): what bev_pool_v2() returns (bev_pool.py:91)  */

/* plan flags (rcb_pool_desc.flags, produced by rcb_pool_validate or implied by
 * rcb_voxel_pooling_prepare_v2) */
#define RCB_PLAN_RANGES_OK 1      /* every rank is inside its tensor                                   */
#define RCB_PLAN_INTERVALS_OK 2   /* intervals tile [0, n_points) contiguously, lengths > 0           */
#define RCB_PLAN_SORTED_CELLS 4   /* interval cells strictly increasing -> zero-filling tile kernel    */
#define RCB_PLAN_STRUCTURED 8     /* ranks_depth unique and ranks_feat == pixel_of(ranks_depth)        */
#define RCB_PLAN_ALL 15

int rcb_version(void);
const char *rcb_error_string(int code);
/* sm count, L2 bytes, compute capability of `device` (host query, no kernel). */
int rcb_device_info(int device, int *sm_count, int *l2_bytes, int *cc_major, int *cc_minor);

/* ------------------------------------------------------------------------------------------
 * Row P -- LSSViewTransformer.voxel_pooling_prepare_v2
 *          (mmdet3d/models/necks/view_transformer.py:207-265)
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int B, N, D, H, W;  /* coor is (B, N, D, H, W, 3) float32 contiguous                          */
  float lower[3];     /* self.grid_lower_bound (x, y, z)  view_transformer.py:80                 */
  float interval[3];  /* self.grid_interval                view_transformer.py:81                 */
  float size[3];      /* self.grid_size (must be integral) view_transformer.py:82-83              */
} rcb_prepare_desc;

size_t rcb_prepare_workspace_bytes(const rcb_prepare_desc *d);

/*
 * Outputs (all int32, device):
 *   ranks_bev / ranks_depth / ranks_feat : capacity P = B*N*D*H*W, first n_kept valid, sorted
 *       by ranks_bev; inside a cell by (ranks_feat, depth bin), i.e. the depth bins of one camera
 *       pixel are adjacent (the reference's argsort leaves this order unspecified).
 *   interval_starts / interval_lengths   : capacity min(P, B*gx*gy*gz), first n_intervals valid.
 *   point_cell  [P]            : global BEV cell of every frustum point, -1 if dropped
 *                                (the inverse map the backward pass uses).
 *   cell_start  [B*G + 1]      : exclusive prefix of points per BEV cell.
 *   counts      [4]            : {n_kept, n_intervals, 0, 0}; read back by the host to slice.
 */
int rcb_voxel_pooling_prepare_v2(const rcb_prepare_desc *d, const float *coor, int *ranks_bev,
                                 int *ranks_depth, int *ranks_feat, int *interval_starts,
                                 int *interval_lengths, int *point_cell, int *cell_start,
                                 int *counts, void *workspace, size_t workspace_bytes, int device,
                                 rcb_stream_t stream);

/*
 * LSSViewTransformer.get_lidar_coor (view_transformer.py:115-157) fused into the
 * prepare stage.  The frustum points are generated inside the first kernel from the calibration
 * instead of being materialised as `coor` (12 bytes per point written, read back and -- end to end
 * -- shipped over PCIe).  All pointers are device pointers, float32:
 *   u [W], v [H], d [D] : the frustum template's axes (create_frustum, view_transformer.py:85-113:
 *                         linspace(0, W_in-1, W), linspace(0, H_in-1, H), arange(*depth_cfg) or the
 *                         SID bins)
 *   cam [B*N][24]       : per camera  inverse(post_rots) 3x3 row-major | post_trans[3] |
 *                         combine = sensor2ego[:3,:3] @ inverse(cam2imgs) 3x3 | sensor2ego[:3,3]
 *   bda [B][9]          : the BEV augmentation 3x3, row-major
 * Operation order per point (every product and sum separately rounded, rows as (m0*x + m1*y) + m2*z):
 *   p = (u,v,d) - post_trans;  p = inv_post_rot * p;  p = (p.x*p.z, p.y*p.z, p.z);
 *   p = combine * p + trans;  p = bda * p.
 * The reference leaves this order to its BLAS (batched 3x3 matmul); outputs are otherwise those of
 * rcb_voxel_pooling_prepare_v2 on the coor so defined.
 */
typedef struct {
  const float *u, *v, *d;
  const float *cam;
  const float *bda;
} rcb_frustum_desc;

int rcb_voxel_pooling_prepare_from_calib(const rcb_prepare_desc *d, const rcb_frustum_desc *fr,
                                         int *ranks_bev, int *ranks_depth, int *ranks_feat,
                                         int *interval_starts, int *interval_lengths, int *point_cell,
                                         int *cell_start, int *counts, void *workspace,
                                         size_t workspace_bytes, int device, rcb_stream_t stream);

/* The first stage alone: point_cell[B*N*D*H*W] (BEV cell of every frustum point, -1 = outside the grid,
 * view_transformer.py:230-249 per point), from `coor` (fr == NULL) or from the calibration (coor ==
 * NULL).  It is all the strip plan (rcb_strip_plan_build) needs: a chain that never hands ranks to
 * its caller -- LSSViewTransformer.voxel_pooling_v2, view_transformer.py:180-205 -- can pool without
 * sorting at all (rcbevdet_b200/view_pool.py, strips mode "chain"). */
int rcb_frustum_point_cells(const rcb_prepare_desc *d, const float *coor, const rcb_frustum_desc *fr,
                            int *point_cell, int device, rcb_stream_t stream);

/* The prepare pipeline in stages: bit 1 = cells (point_cell + the bucket histograms kept in
 * `workspace`), bit 2 = the two sort kernels (ranks_*, cell_start), bit 4 = interval_starts /
 * interval_lengths / counts.  Same workspace across the calls of one pipeline; outputs of stages that
 * are not requested may be NULL.  Grids up to 2^22 cells (the two-level sort); RCB_ERR_UNSUPPORTED
 * beyond.  The sort-free chain runs stage 1 and enqueues stage 2 under a launch gate. */
int rcb_voxel_pooling_prepare_staged(const rcb_prepare_desc *d, const float *coor, const rcb_frustum_desc *fr,
                                     int stages, int *ranks_bev, int *ranks_depth, int *ranks_feat,
                                     int *interval_starts, int *interval_lengths, int *point_cell, int *cell_start,
                                     int *counts, void *workspace, size_t workspace_bytes, int device,
                                     rcb_stream_t stream);

/* Launch gate of the calling host thread (NULL clears it).  While it is set, the kernels behind
 * rcb_voxel_pooling_prepare_v2 / _from_calib / _staged, rcb_bev_pool_v2_fwd (CSR path) and rcb_bev_pool_v2_bwd
 * (structured path, its out_grad transpose included) are launched as usual but exit at once when
 * *gate == 0 on the device; entry points that would take a kernel without a gate return
 * RCB_ERR_UNSUPPORTED.  With gate = the status word of a strip plan this enqueues the general
 * kernels as the fallback of the strip kernels (which exit when the word is non-zero) without
 * reading the word back: exactly one of the two families does the work. */
int rcb_set_launch_gate(const int *gate);

/* Test hook (synchronous, allocates 16 bytes itself): the kernels replace the fp32 division of
 * view_transformer.py:231 by a reciprocal-based sequence that must round identically.  Sweeps all
 * 2^32 numerators for one divisor; *mismatches = quotients that differ from IEEE division in any
 * bit, *first_bad_plus_1 = bit pattern of the first such numerator + 1 (0 when none). */
int rcb_debug_exactdiv_sweep(float divisor, unsigned long long *mismatches,
                             unsigned long long *first_bad_plus_1, int device);

/* ------------------------------------------------------------------------------------------
 * Rows F, B, L -- bev_pool_v2_forward / bev_pool_v2_backward
 *   (mmdet3d/ops/bev_pool_v2/src/bev_pool.cpp:30-57, 74-104; kernels bev_pool_cuda.cu:21-48, 67-121)
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int n_points;    /* K = ranks_*.numel()                                                        */
  int n_intervals; /* I = interval_lengths.size(0)                 (bev_pool.cpp:40)             */
  int C;           /* channels = feat.size(4)                      (bev_pool.cpp:39)             */
  int B, Z, Y, X;  /* bev_feat_shape[0..3]                         (bev_pool.py:27)              */
  int n_depth;     /* depth.numel()  = B*N*D*H*W                                                  */
  int n_pixels;    /* feat rows      = B*N*H*W                                                    */
  int D;           /* depth bins      } describe the frustum; needed only for RCB_PLAN_STRUCTURED */
  int HW;          /* H*W per camera  }                                                           */
  int H;           /* feature rows per camera (0 = unknown); lets the backward walk pixel columns   */
  int layout;      /* RCB_LAYOUT_* of `out` / `out_grad`                                          */
  int feat_dtype;  /* RCB_DTYPE_* of `feat` (depth, out and all gradients are float32)            */
  int flags;       /* RCB_PLAN_* bits known to hold for these ranks                               */
} rcb_pool_desc;

/* Checks the RCB_PLAN_* properties of caller-supplied ranks on the device (the reference trusts
 * them blindly) and writes the bit mask to *flags_out (device int).  point_cell[n_depth] is
 * filled with the BEV cell of every depth element (-1 = not referenced): valid as the backward's
 * inverse map when RCB_PLAN_STRUCTURED comes back set. */
size_t rcb_pool_validate_workspace_bytes(const rcb_pool_desc *d);
int rcb_pool_validate(const rcb_pool_desc *d, const int *ranks_depth, const int *ranks_feat,
                      const int *ranks_bev, const int *interval_starts, const int *interval_lengths,
                      int *point_cell, int *flags_out, void *workspace, size_t workspace_bytes,
                      int device, rcb_stream_t stream);

/* cell_start[B*Z*Y*X + 1]: dense CSR over BEV cells (exclusive prefix of points per cell) from
 * sorted intervals.  Requires RCB_PLAN_SORTED_CELLS.  rcb_voxel_pooling_prepare_v2 emits the
 * same array directly. */
int rcb_pool_build_cellmap(const rcb_pool_desc *d, const int *ranks_bev, const int *interval_starts,
                           int *cell_start, int device, rcb_stream_t stream);

/*
 * Forward.  depth: float32 [n_depth]; feat: [n_pixels, C] channels last, `feat_dtype`;
 * out: float32 [B*Z*Y*X*C] in `layout`, fully written on return (the reference needs it
 * pre-zeroed, bev_pool.py:27, and a permute copy afterwards, bev_pool.py:91 -- neither here).
 *   cell_start != NULL (ranks with RCB_PLAN_SORTED_CELLS): cell-stationary kernel (one warp pools a
 *     cell; C % 4 == 0 up to 128 channels, C % 8 == 0 up to 256).
 *   cell_start == NULL: general path for arbitrary ranks (memset + one warp per interval).
 * Argument order of the rank arrays follows bev_pool_v2_forward (bev_pool.cpp:30-38):
 * interval_lengths BEFORE interval_starts.
 */
int rcb_bev_pool_v2_fwd(const rcb_pool_desc *d, const float *depth, const void *feat,
                        const int *ranks_depth, const int *ranks_feat, const int *ranks_bev,
                        const int *interval_lengths, const int *interval_starts,
                        const int *cell_start, float *out, int device, rcb_stream_t stream);

/*
 * Backward (bev_pool.cpp:74-104).  out_grad float32 in `layout`; depth_grad float32 [n_depth];
 * feat_grad float32 [n_pixels, C].  Both gradients are fully written (zeros where nothing
 * contributes; the reference zero-fills them first, bev_pool.py:67-68).
 *   RCB_PLAN_STRUCTURED + point_cell != NULL: deterministic pixel-stationary kernel; no sort
 *     (the reference re-sorts by ranks_feat on every call, bev_pool.py:47-57).
 *   otherwise: general path (depth_grad exact; feat_grad accumulated with float atomics).
 * workspace >= rcb_pool_bwd_workspace_bytes() (holds out_grad as rows when layout is B_C_CELLS).
 */
size_t rcb_pool_bwd_workspace_bytes(const rcb_pool_desc *d);
int rcb_bev_pool_v2_bwd(const rcb_pool_desc *d, const float *out_grad, const float *depth,
                        const void *feat, const int *ranks_depth, const int *ranks_feat,
                        const int *ranks_bev, const int *point_cell, float *depth_grad,
                        float *feat_grad, void *workspace, size_t workspace_bytes, int device,
                        rcb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Rows F, B on STRIPS (csrc/strips.cu): the same two operators for structured ranks (every frustum
 * point used at most once, RCB_PLAN_STRUCTURED, cells sorted), driven by a plan that groups the
 * points of a vertical run of 16 pixels of one image column (a strip) by BEV cell.  The context
 * rows of a strip are read once and held in registers; one partial row per (cell, strip) segment
 * crosses memory instead of one context row per (cell, pixel) pair (9-13x fewer on the R50 grid).
 * Results equal rcb_bev_pool_v2_fwd / _bwd up to fp32 summation order; both are bit-reproducible.
 *
 *   plan: device buffer of rcb_strip_plan_bytes(); its first int is a status word, 0 = usable.
 *         Non-zero (a strip meets more distinct cells than the plan reserves: only ranks that do not
 *         come from a camera frustum do that) makes the strip kernels exit at once; the caller then
 *         runs -- or has enqueued behind, gated on the same word -- the general entry points.
 *   rows: scratch of rcb_strip_rows_bytes() (one row of C floats per reserved segment; only the used
 *         rows are touched).
 *   cell_start (plan build only): unused since the plan builds its per-cell lists from its own counts
 *         (may be NULL); the plan needs point_cell only -- see rcb_frustum_point_cells.
 * Supported: C in {64, 80, 128} forward, {64, 80} backward; D <= 256; B*Z*Y*X <= 2^24.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int n_img;   /* camera images = B*N                       */
  int D, H, W; /* depth bins, feature rows / columns per image */
  int n_cells; /* B*Z*Y*X                                   */
} rcb_strip_desc;

size_t rcb_strip_plan_bytes(const rcb_strip_desc *d);
size_t rcb_strip_rows_bytes(const rcb_strip_desc *d, int C);
int rcb_strip_plan_build(const rcb_strip_desc *d, const int *point_cell, const int *cell_start, void *plan,
                         size_t plan_bytes, int device, rcb_stream_t stream);
int rcb_bev_pool_v2_fwd_strips(const rcb_pool_desc *d, const rcb_strip_desc *sd, const void *plan,
                               const float *depth, const void *feat, float *out, void *rows, size_t rows_bytes,
                               int device, rcb_stream_t stream);
int rcb_bev_pool_v2_bwd_strips(const rcb_pool_desc *d, const rcb_strip_desc *sd, const void *plan,
                               const float *out_grad, const float *depth, const void *feat, float *depth_grad,
                               float *feat_grad, void *rows, size_t rows_bytes, int device, rcb_stream_t stream);

/* (n_img, C, HW) with image stride `src_img_stride` elements -> (n_img, HW, C) contiguous.
 * Replaces the `feat.contiguous()` transpose copy of bev_pool.py:21 and `out_grad.contiguous()`
 * of bev_pool.py:69.  elem_bytes: 4 (fp32), 2 (bf16/fp16), -2 (fp32 in, bf16 out). */
int rcb_planes_to_rows(const void *src, void *dst, int n_img, int C, int HW,
                       long long src_img_stride, int elem_bytes, int device, rcb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Row R -- PointPillarsScatterRCS.forward up to the two convolutions
 *   (mmdet3d/models/middle_encoders/pillar_scatter.py:64-104, 115-131;
 *    mmdet3d/core/utils/gaussian.py:6-23, 26-55, 57-81)
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int V;        /* pillars                                  */
  int Cin;      /* point feature channels                   */
  int rcs_dim;  /* columns of rcs (7); RCS value = col rcs_dim-2, x,y = cols 0,1 */
  int B, ny, nx;
} rcb_radar_desc;

size_t rcb_radar_workspace_bytes(const rcb_radar_desc *d);

/* point_features (V,Cin) f32; rcs (V,rcs_dim) f32; coors (V,4) int32 [b, z, y, x].
 * features (B,Cin,ny,nx), heatmap (B,ny,nx), heatmap_feat (B,1,ny,nx): all fully written. */
int rcb_radar_rcs_scatter(const rcb_radar_desc *d, const float *point_features, const float *rcs,
                          const int *coors, float *features, float *heatmap, float *heatmap_feat,
                          void *workspace, size_t workspace_bytes, int device, rcb_stream_t stream);

/* features_grad (B,Cin,ny,nx) -> point_features_grad (V,Cin): the gather that is the scatter's
 * gradient (only `features` depends on a differentiable input). */
int rcb_radar_scatter_bwd(const rcb_radar_desc *d, const float *features_grad, const int *coors,
                          float *point_features_grad, int device, rcb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * SURVEY.md 8(f-3) -- BEVDepth4D.gen_grid + shift_feature
 *   (mmdet3d/models/detectors/bevdet_rc.py:585-657): temporal alignment of a pooled BEV feature map.
 * input / output (n, C, H, W) float32 contiguous; tf [n][9]: row-major 3x3 that maps a key-frame BEV
 * pixel (x, y, 1) to the adjacent frame's pixel (the `tf` of bevdet_rc.py:642).  Per output pixel:
 * p = tf * (x, y, 1); normalise by (W-1, H-1) to [-1, 1] (:646-650); grid_sample's align_corners=True
 * un-normalisation; bilinear taps, zero padding -- one kernel, no grid tensor.
 * The backward is the gradient w.r.t. `input` (float atomics, like grid_sample's); input_grad is
 * zero-filled inside.
 * ------------------------------------------------------------------------------------------ */
int rcb_bev_shift_feature(const float *input, const float *tf, float *output, int n, int C, int H, int W,
                          int device, rcb_stream_t stream);
int rcb_bev_shift_feature_bwd(const float *output_grad, const float *tf, float *input_grad, int n, int C,
                              int H, int W, int device, rcb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * SURVEY.md 8(f-2) -- the producer of the pooling inputs, fused
 *   (mmdet3d/models/necks/view_transformer.py:316-319 / :793-797: channel split of the depth-net
 *    output + softmax over the depth bins; mmdet3d/ops/bev_pool_v2/bev_pool.py:21: channels-last copy
 *    of the context).
 * x: (n_img, >= D + C, H, W), `x_dtype` (RCB_DTYPE_*), image stride x_img_stride elements.
 * depth (n_img, D, H, W) float32 = softmax over channels [0, D); context_rows (n_img * H * W, C)
 * float32 = channels [D, D + C) channels last: the `feat` rows of rcb_bev_pool_v2_fwd.  D <= 256.
 * Backward: x_grad (n_img, D + C, H, W) float32 from depth (the forward's output), depth_grad and
 * context_rows_grad.
 * ------------------------------------------------------------------------------------------ */
int rcb_depth_context_split(const void *x, int x_dtype, float *depth, float *context_rows, int n_img, int D,
                            int C, int HW, long long x_img_stride, int device, rcb_stream_t stream);
int rcb_depth_context_split_bwd(const float *depth, const float *depth_grad, const float *context_rows_grad,
                                float *x_grad, int n_img, int D, int C, int HW, int device,
                                rcb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * SURVEY.md 8(f-4) -- the body of the TensorRT plugin `mmdeploy::bev_pool_v2`
 *   (ONNX node: mmdet3d/ops/bev_pool_v2/bev_pool.py:98-119; eager twin :121-142; engine tools
 *    tools/convert_bevdet_to_TRT.py:357-364).  Shaped like IPluginV2DynamicExt::enqueue /
 *   getWorkspaceSize: device pointers in ONNX input order -- inputs[0..6] = depth (N,D,H,W) f32,
 *   feat (N,H,W,C) `feat_dtype`, ranks_depth, ranks_feat, ranks_bev, interval_starts,
 *   interval_lengths (int32) -- outputs[0] = (1, out_height, out_width, C) f32, fully written.
 *   No allocation, no host synchronisation.  sorted_cells = 1 asserts that the interval cells are
 *   strictly increasing (ranks from voxel_pooling_prepare_v2): cell-stationary kernel, CSR rebuilt
 *   in `workspace`; 0: general kernel for any ranks.  INTEGRATION.md has the C++ plugin shell.
 * ------------------------------------------------------------------------------------------ */
size_t rcb_trt_bev_pool_v2_workspace_bytes(int out_height, int out_width, int sorted_cells);
int rcb_trt_bev_pool_v2_enqueue(const void *const *inputs, void *const *outputs, int n_cams, int D, int H,
                                int W, int C, int n_points, int n_intervals, int out_height, int out_width,
                                int feat_dtype, int sorted_cells, void *workspace, size_t workspace_bytes,
                                int device, rcb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RCBEVDET_B200_H_ */
