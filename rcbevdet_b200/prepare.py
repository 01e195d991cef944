"""voxel_pooling_prepare_v2 on the GPU (row P of SURVEY.md section 8).

Mirrors LSSViewTransformer.voxel_pooling_prepare_v2 (mmdet3d/models/necks/view_transformer.py:
207-265): same argument, same 5-tuple in the same order, int32 contiguous tensors on
`coor.device`, or five `None` when no point falls inside the grid.  One CUDA pipeline
(csrc/prepare.cu) and one 16-byte read-back replace ~45 torch kernels and 4 host syncs.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib, plan as _plan


def _grid3(v, name):
    if isinstance(v, torch.Tensor):
        v = v.detach().to("cpu", torch.float32).tolist()
    v = [float(x) for x in v]
    if len(v) != 3:
        raise ValueError(f"{name} must have 3 entries (x, y, z)")
    return v


def _desc(coor, lower, interval, size):
    if coor.dim() != 6 or coor.shape[-1] != 3:
        raise ValueError(f"coor must be (B, N, D, H, W, 3), got {tuple(coor.shape)}")
    if not coor.is_cuda:
        raise RuntimeError("rcbevdet_b200.voxel_pooling_prepare_v2 runs on CUDA tensors only "
                           "(there is no CPU fallback)")
    B, N, D, H, W, _ = coor.shape
    d = _lib.PrepareDesc()
    d.B, d.N, d.D, d.H, d.W = B, N, D, H, W
    # fp32 values of the reference's grid tensors (view_transformer.py:80-83)
    d.lower[:] = _grid3(lower, "grid_lower_bound")
    d.interval[:] = _grid3(interval, "grid_interval")
    d.size[:] = _grid3(size, "grid_size")
    return d


class PreparedRanks:
    """Device-resident result of the prepare pipeline, before any host read-back."""

    __slots__ = ("ranks_bev", "ranks_depth", "ranks_feat", "interval_starts", "interval_lengths",
                 "point_cell", "cell_start", "counts", "grid", "B", "D", "HW", "H", "P", "n_cells")


def prepare_async(coor, grid_lower_bound, grid_interval, grid_size):
    """Launch the pipeline; nothing is read back.  Outputs have capacity P (ranks) and
    min(P, cells) (intervals); `counts` (device int32[4]) holds {n_kept, n_intervals}."""
    desc = _desc(coor, grid_lower_bound, grid_interval, grid_size)
    coor = coor.detach()
    if coor.dtype != torch.float32:
        coor = coor.float()
    coor = coor.contiguous()
    if coor.data_ptr() % 16:                 # an offset view: the kernels fetch 16-byte quads
        coor = coor.clone()
    dev = coor.device
    lib = _lib.lib()
    P = desc.B * desc.N * desc.D * desc.H * desc.W
    gx, gy, gz = (int(desc.size[k]) for k in range(3))
    n_cells = desc.B * gx * gy * gz
    ws_bytes = lib.rcb_prepare_workspace_bytes(ctypes.byref(desc))
    if ws_bytes == 0:
        raise RuntimeError("voxel_pooling_prepare_v2: unsupported geometry (grid_size must be integral, "
                           "B*Z*Y*X <= 2^24 as the reference ranks in fp32, B*N*D*H*W < 2^31)")
    i32 = dict(dtype=torch.int32, device=dev)
    r = PreparedRanks()
    # +4: the vectorised kernels store whole quads
    r.ranks_bev = torch.empty(P, **i32)
    r.ranks_depth = torch.empty(P, **i32)
    r.ranks_feat = torch.empty(P, **i32)
    n_iv = max(1, min(P, n_cells))
    r.interval_starts = torch.empty(n_iv, **i32)
    r.interval_lengths = torch.empty(n_iv, **i32)
    r.point_cell = torch.empty(P + 4, **i32)
    r.cell_start = torch.empty(n_cells + 1, **i32)
    r.counts = torch.empty(4, **i32)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    _lib.check(lib.rcb_voxel_pooling_prepare_v2(
        ctypes.byref(desc), _lib.ptr(coor), _lib.ptr(r.ranks_bev), _lib.ptr(r.ranks_depth),
        _lib.ptr(r.ranks_feat), _lib.ptr(r.interval_starts), _lib.ptr(r.interval_lengths),
        _lib.ptr(r.point_cell), _lib.ptr(r.cell_start), _lib.ptr(r.counts), _lib.ptr(ws), ws_bytes,
        dev.index, _lib.stream_ptr(dev)), "rcb_voxel_pooling_prepare_v2")
    r.grid = (gz, gy, gx)
    r.B, r.D, r.HW, r.H, r.P, r.n_cells = desc.B, desc.D, desc.H * desc.W, desc.H, P, n_cells
    return r


def voxel_pooling_prepare_v2(coor, grid_lower_bound, grid_interval, grid_size):
    """-> (ranks_bev, ranks_depth, ranks_feat, interval_starts, interval_lengths), the reference's
    order (view_transformer.py:263-265), or (None,)*5 when nothing is kept (:258-259).

    Order inside an interval is the stable one (ascending ranks_depth); the reference's
    argsort (:250) leaves it unspecified."""
    r = prepare_async(coor, grid_lower_bound, grid_interval, grid_size)
    n_kept, n_iv = r.counts[:2].tolist()  # the one host sync (the reference has four)
    if n_kept == 0 or n_iv == 0:
        return None, None, None, None, None
    out = (r.ranks_bev[:n_kept], r.ranks_depth[:n_kept], r.ranks_feat[:n_kept],
           r.interval_starts[:n_iv], r.interval_lengths[:n_iv])
    p = _plan.PoolPlan(_lib.PLAN_ALL, r.cell_start, r.point_cell, r.D, r.HW, r.n_cells, r.P)
    _plan.attach(p, out[1], out[2], out[0], out[3], out[4])
    return out


def install(view_transformer_cls):
    """Make `view_transformer_cls.voxel_pooling_prepare_v2` (LSSViewTransformer and subclasses)
    run on this library; `view_transform`, `voxel_pooling_v2`, `init_acceleration_v2` etc. then
    work unchanged on top of it."""

    def _method(self, coor):
        return voxel_pooling_prepare_v2(coor, self.grid_lower_bound, self.grid_interval, self.grid_size)

    _method.__name__ = "voxel_pooling_prepare_v2"
    _method.__doc__ = voxel_pooling_prepare_v2.__doc__
    view_transformer_cls.voxel_pooling_prepare_v2 = _method
    return view_transformer_cls
