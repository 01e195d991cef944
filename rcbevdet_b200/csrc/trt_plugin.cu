// SURVEY.md 8(f-4) -- the body of the TensorRT plugin `mmdeploy::bev_pool_v2`.
// Reference: the ONNX node mmdet3d/ops/bev_pool_v2/bev_pool.py:98-119 creates (inputs depth, feat,
// ranks_depth, ranks_feat, ranks_bev, interval_starts, interval_lengths; attributes out_height,
// out_width) and its eager twin :121-142; the engine-building tools load the plugin library with
// mmdeploy's load_tensorrt_plugin() (tools/convert_bevdet_to_TRT.py:28,220,357-364).  The plugin itself
// lives in mmdeploy (third party, not vendored in the reference); its contract is the eager forward:
//     depth (N, D, H, W) fp32, feat (N, H, W, C) channels last, int32 ranks  ->  (1, out_h, out_w, C).
//
// TensorRT is not in this image, so the IPluginV2DynamicExt shell cannot be compiled here; what it
// calls is: rcb_trt_bev_pool_v2_enqueue has the shape of IPluginV2DynamicExt::enqueue (arrays of
// input / output device pointers in ONNX input order, a caller-owned workspace, a stream, no host
// synchronisation, no allocation), rcb_trt_bev_pool_v2_workspace_bytes that of getWorkspaceSize.
// INTEGRATION.md holds the ~50-line C++ shell a maintainer compiles against NvInfer.h.
//
// Inside enqueue the ranks are runtime tensors the plugin cannot inspect on the host, so:
//   sorted_cells = 0 (default): the general kernel -- memset + one warp per interval, any ranks the
//       reference accepts (what mmdeploy's plugin does);
//   sorted_cells = 1 (plugin attribute the exporter may set: ranks come from
//       voxel_pooling_prepare_v2, interval cells strictly increasing): the dense CSR is rebuilt in the
//       workspace from the intervals and the cell-stationary kernel runs (no memset, empty cells
//       written as zeros by the same kernel).
#include "common.cuh"

extern "C" size_t rcb_trt_bev_pool_v2_workspace_bytes(int out_height, int out_width, int sorted_cells) {
  if (!sorted_cells || out_height <= 0 || out_width <= 0) return 0;
  return rcb::align_up(((size_t)out_height * out_width + 1) * sizeof(int), 256);
}

extern "C" int rcb_trt_bev_pool_v2_enqueue(const void *const *inputs, void *const *outputs, int n_cams, int D, int H,
                                           int W, int C, int n_points, int n_intervals, int out_height,
                                           int out_width, int feat_dtype, int sorted_cells, void *workspace,
                                           size_t workspace_bytes, int device, rcb_stream_t stream) {
  if (!inputs || !outputs || !outputs[0]) return RCB_ERR_ARG;
  if (n_cams <= 0 || D <= 0 || H <= 0 || W <= 0 || C <= 0 || out_height <= 0 || out_width <= 0 || n_points < 0 ||
      n_intervals < 0)
    return RCB_ERR_ARG;
  rcb_pool_desc d{};
  d.n_points = n_points, d.n_intervals = n_intervals, d.C = C;
  d.B = 1, d.Z = 1, d.Y = out_height, d.X = out_width;                      // bev_pool.py:134-135
  d.n_depth = n_cams * D * H * W, d.n_pixels = n_cams * H * W;
  d.D = D, d.HW = H * W, d.H = H;
  d.layout = RCB_LAYOUT_CELLS_C;                                            // (1, out_h, out_w, C): bev_pool.py:140-141
  d.feat_dtype = feat_dtype, d.flags = 0;
  const float *depth = static_cast<const float *>(inputs[0]);
  const void *feat = inputs[1];
  const int *ranks_depth = static_cast<const int *>(inputs[2]), *ranks_feat = static_cast<const int *>(inputs[3]);
  const int *ranks_bev = static_cast<const int *>(inputs[4]);
  const int *interval_starts = static_cast<const int *>(inputs[5]), *interval_lengths = static_cast<const int *>(inputs[6]);
  float *out = static_cast<float *>(outputs[0]);
  const int *cell_start = nullptr;
  if (sorted_cells && n_intervals > 0) {
    if (!workspace || workspace_bytes < rcb_trt_bev_pool_v2_workspace_bytes(out_height, out_width, 1)) return RCB_ERR_WORKSPACE;
    d.flags = RCB_PLAN_RANGES_OK | RCB_PLAN_INTERVALS_OK | RCB_PLAN_SORTED_CELLS;
    int rc = rcb_pool_build_cellmap(&d, ranks_bev, interval_starts, static_cast<int *>(workspace), device, stream);
    if (rc != RCB_OK) return rc;
    cell_start = static_cast<const int *>(workspace);
  }
  return rcb_bev_pool_v2_fwd(&d, depth, feat, ranks_depth, ranks_feat, ranks_bev, interval_lengths, interval_starts,
                             cell_start, out, device, stream);
}
