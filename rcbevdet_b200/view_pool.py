"""LSSViewTransformer.voxel_pooling_v2 (mmdet3d/models/necks/view_transformer.py:180-205) as one
device-side chain: prepare -> pool, with no host synchronisation in between.

The reference method must read four sizes back to the host (boolean-mask compactions, `where`)
before it can launch the pool kernel.  The tile kernel here is driven by the dense per-cell CSR
that prepare leaves on the device, so neither the number of kept points nor the number of
intervals is needed on the host; an empty result (no point inside the grid) simply comes out as
zeros, which is what the reference's special case returns (:184-194).
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib, bev_pool as _bp, plan as _plan, strips as _strips
from .prepare import point_cells_async, prepare_async, prepare_from_calib_async


class _ViewPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, feat, prepared, channels_last=False):
        dev = depth.device
        depth_c = depth.detach().contiguous().float()
        rows = _bp.feat_rows(feat.detach())
        gz, gy, gx = prepared.grid
        C = rows.shape[1]
        d = _lib.PoolDesc()
        d.n_points, d.n_intervals, d.C = prepared.P, 0, C   # upper bounds; the tile kernel reads the CSR
        d.B, d.Z, d.Y, d.X = prepared.B, gz, gy, gx
        d.n_depth, d.n_pixels = depth_c.numel(), rows.shape[0]
        d.D, d.HW, d.H = prepared.D, prepared.HW, prepared.H
        d.layout = _lib.LAYOUT_CELLS_C if channels_last else _lib.LAYOUT_B_C_CELLS
        d.feat_dtype, d.flags = _bp._DTYPES[rows.dtype], _lib.PLAN_ALL
        if d.n_depth != prepared.P or d.n_pixels * d.D != d.n_depth:
            raise ValueError("depth / feat shapes do not match the frustum that `coor` describes")
        shape = (prepared.B, gz, gy, gx, C) if channels_last else (prepared.B, C, gz, gy, gx)
        out = torch.empty(shape, dtype=torch.float32, device=dev)
        plan = _plan.PoolPlan(_lib.PLAN_ALL, prepared.cell_start, prepared.point_cell, prepared.D,
                              prepared.HW, prepared.n_cells, prepared.P)
        sp = _strips.for_forward(plan, d)          # "on" only: the plan of this chain is never reused
        if sp is not None:
            _strips.forward(sp, d, depth_c, rows, out)
        else:
            _bp.pool_forward(d, depth_c, rows, prepared.ranks_depth, prepared.ranks_feat, prepared.ranks_bev, None,
                             None, prepared.cell_start, out)
        ctx.save_for_backward(depth_c, rows, prepared.ranks_depth, prepared.ranks_feat, prepared.ranks_bev)
        ctx.rcb = (d, plan, tuple(feat.shape), feat.dtype, tuple(depth.shape), depth.dtype)
        return out.permute(0, 4, 1, 2, 3) if channels_last else out   # logically (B, C, Z, Y, X) either way

    @staticmethod
    def backward(ctx, out_grad):
        desc, plan, feat_shape, feat_dtype, depth_shape, depth_dtype = ctx.rcb
        depth_grad, feat_grad = _bp._backward(out_grad, ctx.saved_tensors, desc, plan, feat_shape,
                                              feat_dtype, depth_shape, depth_dtype, True)
        return depth_grad, feat_grad, None, None


def _chain_desc(cells, depth_c, rows, channels_last):
    gz, gy, gx = cells.grid
    d = _lib.PoolDesc()
    d.n_points, d.n_intervals, d.C = cells.P, 0, rows.shape[1]
    d.B, d.Z, d.Y, d.X = cells.B, gz, gy, gx
    d.n_depth, d.n_pixels = depth_c.numel(), rows.shape[0]
    d.D, d.HW, d.H = cells.D, cells.HW, cells.H
    d.layout = _lib.LAYOUT_CELLS_C if channels_last else _lib.LAYOUT_B_C_CELLS
    d.feat_dtype, d.flags = _bp._DTYPES[rows.dtype], _lib.PLAN_ALL
    if d.n_depth != cells.P or d.n_pixels * d.D != d.n_depth:
        raise ValueError("depth / feat shapes do not match the frustum that `coor` describes")
    return d


class _ChainPool(torch.autograd.Function):
    """The sort-free chain (strips mode "chain").  Forward: strip kernels on a plan built from point_cell
    alone; behind them, gated on the plan's status word, the sorted pipeline + the cell-stationary
    kernel (nothing is read back; the kernels of the family that does not apply exit at once).
    Backward: the pixel-stationary kernel, which needs point_cell only."""

    @staticmethod
    def forward(ctx, depth, feat, cells, sp, channels_last=False):
        dev = depth.device
        depth_c = depth.detach().contiguous().float()
        rows = _bp.feat_rows(feat.detach())
        d = _chain_desc(cells, depth_c, rows, channels_last)
        gz, gy, gx = cells.grid
        shape = (cells.B, gz, gy, gx, d.C) if channels_last else (cells.B, d.C, gz, gy, gx)
        out = torch.empty(shape, dtype=torch.float32, device=dev)
        _strips.forward(sp, d, depth_c, rows, out)
        with _lib.launch_gate(sp.status_tensor()):
            prepared = cells.launch_sort()
            _bp.pool_forward(d, depth_c, rows, prepared.ranks_depth, prepared.ranks_feat, prepared.ranks_bev, None,
                             None, prepared.cell_start, out)
        ctx.save_for_backward(depth_c, rows, cells.point_cell)
        ctx.rcb = (d, tuple(feat.shape), feat.dtype, tuple(depth.shape), depth.dtype)
        return out.permute(0, 4, 1, 2, 3) if channels_last else out

    @staticmethod
    def backward(ctx, out_grad):
        desc, feat_shape, feat_dtype, depth_shape, depth_dtype = ctx.rcb
        depth, rows, point_cell = ctx.saved_tensors
        dev = depth.device
        out_grad = out_grad.float()
        desc = _lib.PoolDesc.from_buffer_copy(desc)
        if out_grad.dim() == 5 and not out_grad.is_contiguous() and out_grad.permute(0, 2, 3, 4, 1).is_contiguous():
            desc.layout = _lib.LAYOUT_CELLS_C            # rows in place
        else:
            out_grad = out_grad.contiguous()
            desc.layout = _lib.LAYOUT_B_C_CELLS
        depth_grad = torch.empty(depth.shape, dtype=torch.float32, device=dev)
        feat_grad = torch.empty(rows.shape, dtype=torch.float32, device=dev)
        # The pixel-stationary kernel: it needs point_cell only (no sort), is never slower than the strip
        # backward (B = 8 R50: 94 against 100 us, 152 us under image rotation; hi-res B = 1: 73 against
        # 132 us) and takes every geometry, so the backward needs no plan, no gate and no fallback.
        lib = _lib.lib()
        ws_bytes = lib.rcb_pool_bwd_workspace_bytes(ctypes.byref(desc))
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
        _lib.check(lib.rcb_bev_pool_v2_bwd(
            ctypes.byref(desc), _lib.ptr(out_grad), _lib.ptr(depth), _lib.ptr(rows), None, None, None,
            _lib.ptr(point_cell), _lib.ptr(depth_grad), _lib.ptr(feat_grad), _lib.ptr(ws), ws_bytes, dev.index,
            _lib.stream_ptr(dev)), "rcb_bev_pool_v2_bwd")
        depth_grad, feat_grad = depth_grad.view(depth_shape), feat_grad.view(feat_shape)
        return (depth_grad if depth_dtype == torch.float32 else depth_grad.to(depth_dtype),
                feat_grad if feat_dtype == torch.float32 else feat_grad.to(feat_dtype), None, None, None)


def _chain(cells, C, depth, feat, channels_last):
    """The sort-free chain when it applies (mode, channel count, geometry), else None.  `feat` is the
    (B, N, H, W, C) view the pooling operators take (view_transformer.py:195)."""
    if cells is None or _strips.MODE != "chain" or C not in _strips.FWD_CHANNELS:
        return None
    sp = _strips.build(cells.point_cell, None, cells.n_img, cells.D, cells.H, cells.W, cells.n_cells)
    if sp is None:
        return None
    return _ChainPool.apply(depth, feat, cells, sp, channels_last)


def _chain_wanted(C, return_prepared):
    return _strips.MODE == "chain" and not return_prepared and C in _strips.FWD_CHANNELS


def _collapse_z(bev):
    """view_transformer.py:203-204: (B, C, Z, Y, X) -> (B, C*Z, Y, X).  Z == 1 is a view (it keeps a
    channels-last result channels-last); the general case concatenates like the reference."""
    return bev.squeeze(2) if bev.shape[2] == 1 else torch.cat(bev.unbind(dim=2), 1)


def fused_path_supports(C):
    """The sync-free chain drives the CSR (cell-stationary) forward kernels, which own whole
    128-bit channel quads: C % 4 == 0 up to 128 channels, C % 8 == 0 up to 256.  Any other channel
    count (the reference accepts every C) takes the two-call route below, which reads the counts
    back once."""
    return 0 < C <= 256 and C % 4 == 0 and (C <= 128 or C % 8 == 0)


def voxel_pooling_v2(coor, depth, feat, grid_lower_bound, grid_interval, grid_size, collapse_z=True,
                     return_prepared=False, channels_last=False):
    """coor (B,N,D,H,W,3) fp32; depth (B,N,D,H,W); feat (B,N,C,H,W) -- exactly the arguments of
    the reference method.  Returns bev_feat (B, C*Z, Y, X) (collapse_z) or (B,C,Z,Y,X)."""
    C = int(feat.shape[2])
    if not fused_path_supports(C):
        # view_transformer.py:180-205 literally: prepare (one read-back of the counts), then the op
        # through its general kernels, with the real n_kept / n_intervals
        from .prepare import voxel_pooling_prepare_v2
        if return_prepared:
            raise ValueError("return_prepared needs the fused path (see fused_path_supports)")
        rb, rd, rf, st, ln = voxel_pooling_prepare_v2(coor, grid_lower_bound, grid_interval, grid_size)
        B = int(coor.shape[0])
        gx, gy, gz = (int(float(v)) for v in (grid_size.tolist() if isinstance(grid_size, torch.Tensor)
                                              else grid_size))
        if rb is None:                                           # :184-194
            bev = torch.zeros((B, C, gz, gy, gx), dtype=torch.float32, device=depth.device)
            bev = bev + 0.0 * (depth.sum() + feat.sum())         # keep the graph connected
        else:
            bev = _bp.bev_pool_v2(depth, feat.permute(0, 1, 3, 4, 2), rd, rf, rb, (B, gz, gy, gx, C), st, ln)
        return torch.cat(bev.unbind(dim=2), 1) if collapse_z else bev
    if _chain_wanted(C, return_prepared):
        cells = point_cells_async(coor=coor, grid_lower_bound=grid_lower_bound, grid_interval=grid_interval,
                                  grid_size=grid_size)
        bev = _chain(cells, C, depth, feat.permute(0, 1, 3, 4, 2), channels_last)
        if bev is not None:
            return _collapse_z(bev) if collapse_z else bev
    prepared = prepare_async(coor, grid_lower_bound, grid_interval, grid_size)
    feat = feat.permute(0, 1, 3, 4, 2)                       # view_transformer.py:195
    bev = _ViewPool.apply(depth, feat, prepared, channels_last)
    if collapse_z:
        bev = _collapse_z(bev)                               # view_transformer.py:203-204
    return (bev, prepared) if return_prepared else bev


def voxel_pooling_v2_from_calib(calib, axes, depth, feat, grid_lower_bound, grid_interval, grid_size,
                                collapse_z=True, return_prepared=False, channels_last=False):
    """get_lidar_coor + voxel_pooling_v2 (view_transformer.py:290-294) as one device-side chain:
    `calib` = (sensor2ego, ego2global, cam2imgs, post_rots, post_trans, bda) or a packed (cam, bda)
    pair, `axes` = frustum_axes(...); depth (B,N,D,H,W); feat (B,N,C,H,W).  The frustum points
    are generated inside the first prepare kernel; nothing of size P crosses PCIe or HBM as `coor`."""
    C = int(feat.shape[2])
    if _chain_wanted(C, return_prepared):
        cells = point_cells_async(calib=calib, axes=axes, grid_lower_bound=grid_lower_bound,
                                  grid_interval=grid_interval, grid_size=grid_size, device=depth.device)
        bev = _chain(cells, C, depth, feat.permute(0, 1, 3, 4, 2), channels_last)
        if bev is not None:
            return _collapse_z(bev) if collapse_z else bev
    prepared = prepare_from_calib_async(calib, axes, grid_lower_bound, grid_interval, grid_size, device=depth.device)
    if not fused_path_supports(C):
        from .prepare import _finish
        if return_prepared:
            raise ValueError("return_prepared needs the fused path (see fused_path_supports)")
        rb, rd, rf, st, ln = _finish(prepared)
        gz, gy, gx = prepared.grid
        if rb is None:
            bev = torch.zeros((prepared.B, C, gz, gy, gx), dtype=torch.float32, device=depth.device)
            bev = bev + 0.0 * (depth.sum() + feat.sum())
        else:
            bev = _bp.bev_pool_v2(depth, feat.permute(0, 1, 3, 4, 2), rd, rf, rb, (prepared.B, gz, gy, gx, C), st, ln)
        return torch.cat(bev.unbind(dim=2), 1) if collapse_z else bev
    bev = _ViewPool.apply(depth, feat.permute(0, 1, 3, 4, 2), prepared, channels_last)
    if collapse_z:
        bev = _collapse_z(bev)
    return (bev, prepared) if return_prepared else bev
