"""The producer side of the pooling: LSSViewTransformer.forward's
    depth_digit = x[:, :D];  tran_feat = x[:, D:D+C];  depth = depth_digit.softmax(dim=1)
(mmdet3d/models/necks/view_transformer.py:316-319, BEVDepth variant :793-797) fused with the
channels-last copy of the context that the pooling op makes (mmdet3d/ops/bev_pool_v2/bev_pool.py:21).

    depth_context_split(x, D, C) -> depth (n_img, D, H, W) fp32, context (n_img, H, W, C) fp32
one kernel, x read once; its backward is one kernel too (softmax backward + the transpose back).

    lss_view_transform(x, n_cams, D, C, calib, axes, grid...) -> (bev_feat, depth)
is forward()'s tail from the depth-net output on: split -> get_lidar_coor + prepare (fused, from the
calibration) -> bev_pool_v2, without the context transpose pass and without a host sync.
There is no CPU path.
"""
from __future__ import annotations

import torch

from . import _lib
from .bev_pool import _DTYPES

__all__ = ["depth_context_split", "lss_view_transform"]


class _DepthContextSplit(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, D, C):
        if not x.is_cuda:
            raise RuntimeError("depth_context_split runs on CUDA tensors only (rcbevdet_b200 has no CPU fallback)")
        if x.dim() != 4 or x.shape[1] < D + C:
            raise ValueError(f"x must be (n_img, >= D + C, H, W), got {tuple(x.shape)} for D={D}, C={C}")
        xd = x.detach()
        if xd.dtype not in _DTYPES:
            xd = xd.float()
        xd = xd.contiguous()
        n_img, c_tot, H, W = xd.shape
        dev = xd.device
        depth = torch.empty((n_img, D, H, W), dtype=torch.float32, device=dev)
        rows = torch.empty((n_img, H, W, C), dtype=torch.float32, device=dev)
        _lib.check(_lib.lib().rcb_depth_context_split(_lib.ptr(xd), _DTYPES[xd.dtype], _lib.ptr(depth), _lib.ptr(rows),
                                                      n_img, D, C, H * W, c_tot * H * W, dev.index,
                                                      _lib.stream_ptr(dev)), "rcb_depth_context_split")
        ctx.save_for_backward(depth)
        ctx.rcb = (D, C, c_tot, x.dtype)
        return depth, rows

    @staticmethod
    def backward(ctx, g_depth, g_rows):
        (depth,) = ctx.saved_tensors
        D, C, c_tot, dtype = ctx.rcb
        n_img, _, H, W = depth.shape
        dev = depth.device
        g_depth = torch.zeros_like(depth) if g_depth is None else g_depth.contiguous().float()
        g_rows = (torch.zeros((n_img, H, W, C), dtype=torch.float32, device=dev) if g_rows is None
                  else g_rows.contiguous().float())
        if c_tot == D + C:
            dx = torch.empty((n_img, c_tot, H, W), dtype=torch.float32, device=dev)
            target = dx
        else:  # channels the split does not read get a zero gradient
            dx = torch.zeros((n_img, c_tot, H, W), dtype=torch.float32, device=dev)
            target = torch.empty((n_img, D + C, H, W), dtype=torch.float32, device=dev)
        _lib.check(_lib.lib().rcb_depth_context_split_bwd(_lib.ptr(depth), _lib.ptr(g_depth), _lib.ptr(g_rows),
                                                          _lib.ptr(target), n_img, D, C, H * W, dev.index,
                                                          _lib.stream_ptr(dev)), "rcb_depth_context_split_bwd")
        if target is not dx:
            dx[:, :D + C] = target
        return dx.to(dtype), None, None


def depth_context_split(x, D, C):
    """x (n_img, D + C, H, W), the depth-net output -> (depth, context):
    depth = x[:, :D].softmax(dim=1) as (n_img, D, H, W) fp32 and context = x[:, D:D+C] channels
    last, (n_img, H, W, C) fp32 contiguous -- view it as (B, N, H, W, C) and it IS the `feat`
    argument of bev_pool_v2 (view_transformer.py:278-279), already in the layout the kernels read."""
    return _DepthContextSplit.apply(x, int(D), int(C))


def lss_view_transform(x, n_cams, D, C, calib, axes, grid_lower_bound, grid_interval, grid_size, collapse_z=True,
                       channels_last=False):
    """LSSViewTransformer.forward from the depth-net output on (view_transformer.py:316-320 +
    view_transform_core's non-accelerated branch :290-294) as one device-side chain:
    x (B*N, D + C, H, W); calib = get_lidar_coor's six tensors (or a packed pair from pack_calib);
    axes = frustum_axes(...).  Returns (bev_feat, depth) like the reference: bev_feat (B, C*Z, Y, X)
    (collapse_z) and depth (B*N, D, H, W)."""
    from .prepare import point_cells_async, prepare_from_calib_async
    from .view_pool import _ViewPool, _chain, _chain_wanted, _collapse_z, fused_path_supports
    if not fused_path_supports(C):
        raise ValueError(f"lss_view_transform needs C % 4 == 0 (C % 8 above 128 channels, C <= 256), got {C}")
    bn, _, H, W = x.shape
    if bn % n_cams:
        raise ValueError("x.shape[0] must be B * n_cams")
    B = bn // n_cams
    depth, ctx_cl = depth_context_split(x, D, C)
    bev = None
    if _chain_wanted(C, False):                                 # strips mode "chain": pool without sorting
        cells = point_cells_async(calib=calib, axes=axes, grid_lower_bound=grid_lower_bound,
                                  grid_interval=grid_interval, grid_size=grid_size, device=x.device)
        bev = _chain(cells, C, depth.view(B, n_cams, D, H, W), ctx_cl.view(B, n_cams, H, W, C), channels_last)
    if bev is not None:
        return (_collapse_z(bev) if collapse_z else bev), depth
    prepared = prepare_from_calib_async(calib, axes, grid_lower_bound, grid_interval, grid_size, device=x.device)
    bev = _ViewPool.apply(depth.view(B, n_cams, D, H, W), ctx_cl.view(B, n_cams, H, W, C), prepared, channels_last)
    if collapse_z:
        bev = _collapse_z(bev)
    return bev, depth
