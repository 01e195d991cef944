"""bev_pool_v2 for B200: the reference's operator surface on hand-written sm_100a kernels.

Same names, argument order and return layouts as mmdet3d/ops/bev_pool_v2/bev_pool.py:
  * QuickCumsumCuda(torch.autograd.Function)  -- forward returns (B, Z, Y, X, C)   (:16-41)
  * bev_pool_v2(...) -> (B, C, Z, Y, X) contiguous float32                          (:86-92)
  * TRTBEVPoolv2 (ONNX symbolic `mmdeploy::bev_pool_v2` + eager forward)            (:95-142)
so `from mmdet3d.ops.bev_pool_v2.bev_pool import bev_pool_v2, TRTBEVPoolv2` can be pointed here
and LSSViewTransformer.view_transform runs unchanged (view_transformer.py:11,199,284).

What differs underneath: no zero-fill pass, no permute copy (the kernel writes the final layout),
the non-contiguous (B,N,H,W,C) context view is transposed by one kernel (or consumed as bf16/fp16
rows), the backward needs no argsort, and everything runs on the caller's current stream.
There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib, plan as _plan, strips as _strips

__all__ = ["bev_pool_v2", "TRTBEVPoolv2", "QuickCumsumCuda"]

_DTYPES = {torch.float32: _lib.DTYPE_F32, torch.bfloat16: _lib.DTYPE_BF16, torch.float16: _lib.DTYPE_F16}


def _require_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError(f"bev_pool_v2: `{name}` must be a CUDA tensor; rcbevdet_b200 has no CPU fallback")


def feat_rows(feat):
    """Context features as contiguous channels-last rows (n_pixels, C).

    The reference calls feat.contiguous() (bev_pool.py:21), a strided transpose copy, because
    `feat` arrives as the (B,N,H,W,C) permuted view of an NCHW tensor (view_transformer.py:195).
    Here that case goes through one tiled transpose kernel; rows that are already contiguous are
    used in place.  fp32 stays fp32; fp16/bf16 stay 16-bit (widened in registers, fp32
    accumulation -- identical to the reference's `.float()` up-cast)."""
    if feat.dtype not in _DTYPES:
        feat = feat.float()
    C = feat.shape[-1]
    if feat.is_contiguous():
        return feat.view(-1, C)
    if feat.dim() == 5:
        B, N, H, W, _ = feat.shape
        sb, sn, sh, sw, sc = feat.stride()
        planar = sw == 1 and sh == W and sc == H * W and (B == 1 or sb == N * sn) and sn >= C * H * W
        if planar and B * N <= 65535:
            rows = torch.empty((B * N * H * W, C), dtype=feat.dtype, device=feat.device)
            _lib.check(_lib.lib().rcb_planes_to_rows(_lib.ptr(feat), _lib.ptr(rows), B * N, C, H * W, sn,
                                                     feat.element_size(), feat.device.index,
                                                     _lib.stream_ptr(feat.device)), "rcb_planes_to_rows")
            return rows
    return feat.contiguous().view(-1, C)


def _pool_desc(depth, rows, ranks_depth, interval_lengths, bev_feat_shape, layout):
    B, Z, Y, X, C = (int(s) for s in bev_feat_shape)
    if C != rows.shape[1]:
        raise ValueError(f"bev_feat_shape[-1]={C} does not match feat channels {rows.shape[1]}")
    d = _lib.PoolDesc()
    d.n_points, d.n_intervals, d.C = ranks_depth.numel(), interval_lengths.numel(), C
    d.B, d.Z, d.Y, d.X = B, Z, Y, X
    d.n_depth, d.n_pixels = depth.numel(), rows.shape[0]
    if depth.dim() == 5 and depth.shape[0] * depth.shape[1] * depth.shape[3] * depth.shape[4] == rows.shape[0]:
        d.D, d.HW, d.H = depth.shape[2], depth.shape[3] * depth.shape[4], depth.shape[3]
    else:
        d.D, d.HW, d.H = 0, 0, 0
    d.layout, d.feat_dtype, d.flags = layout, _DTYPES[rows.dtype], 0
    return d


def pool_forward(desc, depth, rows, ranks_depth, ranks_feat, ranks_bev, interval_lengths, interval_starts,
                 cell_start, out):
    """rcb_bev_pool_v2_fwd on the caller's current stream."""
    dev = out.device
    _lib.check(_lib.lib().rcb_bev_pool_v2_fwd(
        ctypes.byref(desc), _lib.ptr(depth), _lib.ptr(rows), _lib.ptr(ranks_depth), _lib.ptr(ranks_feat),
        _lib.ptr(ranks_bev), _lib.ptr(interval_lengths), _lib.ptr(interval_starts), _lib.ptr(cell_start),
        _lib.ptr(out), dev.index, _lib.stream_ptr(dev)), "rcb_bev_pool_v2_fwd")


def _forward(depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape, interval_starts,
             interval_lengths, layout):
    for t, n in ((depth, "depth"), (feat, "feat"), (ranks_depth, "ranks_depth"), (ranks_feat, "ranks_feat"),
                 (ranks_bev, "ranks_bev"), (interval_starts, "interval_starts"),
                 (interval_lengths, "interval_lengths")):
        _require_cuda(t, n)
    dev = depth.device
    depth = depth.detach().contiguous().float()          # bev_pool.py:19
    rows = feat_rows(feat.detach())                      # bev_pool.py:20
    if not (ranks_depth.numel() == ranks_feat.numel() == ranks_bev.numel()):
        raise ValueError("ranks_depth, ranks_feat and ranks_bev must have the same length")
    if interval_starts.numel() != interval_lengths.numel():
        raise ValueError("interval_starts and interval_lengths must have the same length")
    originals = (ranks_depth, ranks_feat, ranks_bev, interval_starts, interval_lengths)
    desc = _pool_desc(depth, rows, ranks_depth, interval_lengths, bev_feat_shape, layout)
    n_cells = desc.B * desc.Z * desc.Y * desc.X
    # the plan is keyed on the caller's own tensors; int64 / strided ranks are converted to int32
    # once, inside the plan (bev_pool.py:18,22-25 converts on every call)
    plan = _plan.lookup(*originals, n_cells, desc.n_depth)
    if plan is None:
        plan = _plan.derive(desc, originals)
    ranks_depth, ranks_feat, ranks_bev, interval_starts, interval_lengths = plan.ranks
    desc.flags = plan.flags
    if plan.structured:
        desc.D, desc.HW = plan.D, plan.HW
    B, Z, Y, X, C = desc.B, desc.Z, desc.Y, desc.X, desc.C
    shape = (B, C, Z, Y, X) if layout == _lib.LAYOUT_B_C_CELLS else (B, Z, Y, X, C)
    out = torch.empty(shape, dtype=torch.float32, device=dev)
    sp = _strips.for_forward(plan, desc)
    if sp is not None:
        _strips.forward(sp, desc, depth, rows, out)
    else:
        pool_forward(desc, depth, rows, ranks_depth, ranks_feat, ranks_bev, interval_lengths, interval_starts,
                     plan.cell_start if plan.sorted_cells else None, out)
    saved = (depth, rows, ranks_depth, ranks_feat, ranks_bev)
    return out, saved, desc, plan


def _backward(out_grad, saved, desc, plan, feat_shape, feat_dtype, depth_shape, depth_dtype, logical_bczyx=None):
    """`logical_bczyx`: out_grad has the shape (B, C, Z, Y, X) (what bev_pool_v2 returns) rather than
    (B, Z, Y, X, C); None = what desc.layout says.  Its MEMORY decides which kernel path runs: a
    gradient that is channels-last in memory (the forward was asked for `channels_last`, or the
    consumer runs in torch.channels_last) is used in place as rows; a (B, C, cells)-contiguous one goes
    through the transpose kernel, as the reference's out_grad.contiguous() does (bev_pool.py:69)."""
    depth, rows, ranks_depth, ranks_feat, ranks_bev = saved
    dev = depth.device
    if logical_bczyx is None:
        logical_bczyx = desc.layout == _lib.LAYOUT_B_C_CELLS
    out_grad = out_grad.float()
    bwd_desc = _lib.PoolDesc.from_buffer_copy(desc)      # the forward's descriptor is shared with the caller
    if not logical_bczyx:
        out_grad = out_grad.contiguous()                 # bev_pool.py:69
        bwd_desc.layout = _lib.LAYOUT_CELLS_C
    elif out_grad.dim() == 5 and not out_grad.is_contiguous() and out_grad.permute(0, 2, 3, 4, 1).is_contiguous():
        bwd_desc.layout = _lib.LAYOUT_CELLS_C            # rows in place: no transpose pass
    else:
        out_grad = out_grad.contiguous()
        bwd_desc.layout = _lib.LAYOUT_B_C_CELLS
    desc = bwd_desc
    depth_grad = torch.empty(depth.shape, dtype=torch.float32, device=dev)
    feat_grad = torch.empty(rows.shape, dtype=torch.float32, device=dev)
    sp = _strips.for_backward(plan, desc)
    if sp is not None:
        _strips.backward(sp, desc, out_grad, depth, rows, depth_grad, feat_grad)
        depth_grad, feat_grad = depth_grad.view(depth_shape), feat_grad.view(feat_shape)
        return (depth_grad if depth_dtype == torch.float32 else depth_grad.to(depth_dtype),
                feat_grad if feat_dtype == torch.float32 else feat_grad.to(feat_dtype))
    lib = _lib.lib()
    ws_bytes = lib.rcb_pool_bwd_workspace_bytes(ctypes.byref(desc))
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
    _lib.check(lib.rcb_bev_pool_v2_bwd(
        ctypes.byref(desc), _lib.ptr(out_grad), _lib.ptr(depth), _lib.ptr(rows), _lib.ptr(ranks_depth),
        _lib.ptr(ranks_feat), _lib.ptr(ranks_bev), _lib.ptr(plan.point_cell if plan.structured else None),
        _lib.ptr(depth_grad), _lib.ptr(feat_grad), _lib.ptr(ws), ws_bytes, dev.index,
        _lib.stream_ptr(dev)), "rcb_bev_pool_v2_bwd")
    depth_grad = depth_grad.view(depth_shape)
    feat_grad = feat_grad.view(feat_shape)
    if depth_dtype != torch.float32:
        depth_grad = depth_grad.to(depth_dtype)
    if feat_dtype != torch.float32:
        feat_grad = feat_grad.to(feat_dtype)
    return depth_grad, feat_grad


class _PoolFunction(torch.autograd.Function):
    _layout = _lib.LAYOUT_CELLS_C

    @classmethod
    def _fwd(cls, ctx, depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape, interval_starts,
             interval_lengths):
        out, saved, desc, plan = _forward(depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape,
                                          interval_starts, interval_lengths, cls._layout)
        ctx.save_for_backward(*saved)
        ctx.rcb = (desc, plan, tuple(feat.shape), feat.dtype, tuple(depth.shape), depth.dtype)
        return out

    @staticmethod
    def _bwd(ctx, out_grad):
        desc, plan, feat_shape, feat_dtype, depth_shape, depth_dtype = ctx.rcb[:6]
        logical = ctx.rcb[6] if len(ctx.rcb) > 6 else None
        depth_grad, feat_grad = _backward(out_grad, ctx.saved_tensors, desc, plan, feat_shape, feat_dtype,
                                          depth_shape, depth_dtype, logical)
        return depth_grad, feat_grad, None, None, None, None, None, None


class QuickCumsumCuda(_PoolFunction):
    """Drop-in for the reference's autograd Function (bev_pool.py:11-83).  forward returns the
    pooled features channels-last, (B, Z, Y, X, C); backward returns gradients for `depth` and
    `feat` only."""
    _layout = _lib.LAYOUT_CELLS_C

    @staticmethod
    def forward(ctx, depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape, interval_starts,
                interval_lengths):
        return QuickCumsumCuda._fwd(ctx, depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape,
                                    interval_starts, interval_lengths)

    @staticmethod
    def backward(ctx, out_grad):
        return _PoolFunction._bwd(ctx, out_grad)


class _BevPoolV2Fused(_PoolFunction):
    """QuickCumsumCuda + the permute(0,4,1,2,3).contiguous() of bev_pool.py:91 in one kernel:
    forward returns (B, C, Z, Y, X) directly."""
    _layout = _lib.LAYOUT_B_C_CELLS

    @staticmethod
    def forward(ctx, depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape, interval_starts,
                interval_lengths):
        return _BevPoolV2Fused._fwd(ctx, depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape,
                                    interval_starts, interval_lengths)

    @staticmethod
    def backward(ctx, out_grad):
        return _PoolFunction._bwd(ctx, out_grad)


class _BevPoolV2ChannelsLast(_PoolFunction):
    """bev_pool_v2 with a channels-last result: same shape (B, C, Z, Y, X), memory (B, Z, Y, X, C).
    The forward writes rows directly and the backward takes a channels-last gradient in place: neither
    the permute copy of bev_pool.py:91 nor its mirror image at :69 exists in any form."""
    _layout = _lib.LAYOUT_CELLS_C

    @staticmethod
    def forward(ctx, depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape, interval_starts,
                interval_lengths):
        out = _BevPoolV2ChannelsLast._fwd(ctx, depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape,
                                          interval_starts, interval_lengths)
        ctx.rcb = ctx.rcb + (True,)                      # out_grad is logically (B, C, Z, Y, X)
        return out.permute(0, 4, 1, 2, 3)

    @staticmethod
    def backward(ctx, out_grad):
        return _PoolFunction._bwd(ctx, out_grad)


def bev_pool_v2(depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape, interval_starts,
                interval_lengths, channels_last=False):
    """depth (B,N,D,H,W); feat (B,N,H,W,C) (any strides); ranks/intervals as produced by
    voxel_pooling_prepare_v2; bev_feat_shape = (B,Z,Y,X,C).  Returns (B,C,Z,Y,X) contiguous float32
    (bev_pool.py:86-92).  channels_last=True (an extension; the reference has no such mode): the same
    shape with channels-last memory -- no transposed write, and a channels-last gradient is consumed
    in place by the backward (a BEV encoder run in torch.channels_last produces one)."""
    fn = _BevPoolV2ChannelsLast if channels_last else _BevPoolV2Fused
    return fn.apply(depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape, interval_starts,
                    interval_lengths)


class TRTBEVPoolv2(torch.autograd.Function):
    """bev_pool.py:95-142: ONNX export stub + eager forward used by the TensorRT-style model
    wrappers (bevdet.py:549-551)."""

    @staticmethod
    def symbolic(g, depth, feat, ranks_depth, ranks_feat, ranks_bev, interval_starts, interval_lengths,
                 out_height=128, out_width=128):
        return g.op("mmdeploy::bev_pool_v2", depth, feat, ranks_depth, ranks_feat, ranks_bev,
                    interval_starts, interval_lengths, out_height_i=out_height, out_width_i=out_width)

    @staticmethod
    def forward(g, depth, feat, ranks_depth, ranks_feat, ranks_bev, interval_starts, interval_lengths,
                out_height=128, out_width=128):
        feat = feat.unsqueeze(0)      # (N,H,W,C) -> (1,N,H,W,C)
        depth = depth.unsqueeze(0)    # (N,D,H,W) -> (1,N,D,H,W)
        bev_feat_shape = (depth.shape[0], 1, out_height, out_width, feat.shape[-1])
        bev_feat = bev_pool_v2(depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape,
                               interval_starts, interval_lengths)
        return bev_feat.squeeze(2).permute(0, 2, 3, 1)
