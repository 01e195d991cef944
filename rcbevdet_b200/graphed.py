"""CUDA-graph replay of the calibration-driven chain for inference.

One sample (or a few) is host-launch-bound when the chain is launched eagerly: 11 kernel launches
through ctypes and a dozen allocations cost ~250 us of host time for ~100 us of device time (B = 1 R50).
The chain (`voxel_pooling_v2_from_calib`, view_transformer.py:290-294 + :180-205) reads nothing back
and allocates only through torch's caching allocator, so it captures as it is; this class owns the
static input buffers, captures once per input signature and replays."""
from __future__ import annotations

import torch

from .prepare import pack_calib
from .view_pool import voxel_pooling_v2_from_calib


class GraphedViewPool:
    """pool = GraphedViewPool(axes, grid_lower_bound, grid_interval, grid_size)
    bev = pool(calib, depth, feat)      # calib: get_lidar_coor's six tensors or a packed (cam, bda) pair

    Forward only (inference: the reference's `accelerate` use case, but with a per-frame calibration).
    The returned tensor is the graph's output buffer: it is overwritten by the next call with the same
    input signature -- clone it to keep it."""

    def __init__(self, axes, grid_lower_bound, grid_interval, grid_size, collapse_z=True, channels_last=False):
        self.axes = axes
        self.grid = (grid_lower_bound, grid_interval, grid_size)
        self.collapse_z, self.channels_last = collapse_z, channels_last
        self._graphs = {}

    def _capture(self, cam, bda, depth, feat):
        dev = depth.device
        static = tuple(torch.empty_like(t, device=dev) for t in (cam, bda, depth, feat))
        for dst, src in zip(static, (cam, bda, depth, feat)):
            dst.copy_(src, non_blocking=True)
        axes = tuple(t.to(dev) for t in self.axes)

        def run():
            with torch.no_grad():
                return voxel_pooling_v2_from_calib((static[0], static[1]), axes, static[2], static[3], *self.grid,
                                                   collapse_z=self.collapse_z, channels_last=self.channels_last)

        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):          # warm-up outside the capture: kernel attributes, allocator pools
                run()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            out = run()
        return graph, static, out

    def __call__(self, calib, depth, feat):
        if not depth.is_cuda:
            raise RuntimeError("GraphedViewPool runs on CUDA tensors only (there is no CPU fallback)")
        cam, bda = pack_calib(*calib) if len(calib) == 6 else calib
        cam = cam.to(device=depth.device, dtype=torch.float32, non_blocking=True)
        bda = bda.to(device=depth.device, dtype=torch.float32, non_blocking=True)
        key = (depth.device, tuple(depth.shape), depth.dtype, tuple(feat.shape), feat.dtype, tuple(cam.shape))
        entry = self._graphs.get(key)
        if entry is None:
            entry = self._graphs[key] = self._capture(cam, bda, depth.detach(), feat.detach())
        graph, static, out = entry
        for dst, src in zip(static, (cam, bda, depth.detach(), feat.detach())):
            dst.copy_(src, non_blocking=True)
        graph.replay()
        return out
