"""Drive the reference's own CUDA kernels (oracle/_ref/bev_pool_v2_ext.so, built unmodified by
oracle/build_ref.py) with the exact host-side sequence of the reference's autograd Function
(mmdet3d/ops/bev_pool_v2/bev_pool.py:16-41 forward, :43-83 backward, :86-92 permute).
TEST INFRASTRUCTURE ONLY: the checker for the GPU parity tests and an optional timing line in
bench.py; never imported by the product package."""
from __future__ import annotations

import torch

from . import build_ref


def available():
    return build_ref.load() is not None and torch.cuda.is_available()


def forward(depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape, interval_starts, interval_lengths):
    """-> (B, Z, Y, X, C) float32, what QuickCumsumCuda.forward returns (bev_pool.py:16-41)."""
    ext = build_ref.load()
    ranks_bev = ranks_bev.int()
    depth = depth.contiguous().float()
    feat = feat.contiguous().float()
    ranks_depth = ranks_depth.contiguous().int()
    ranks_feat = ranks_feat.contiguous().int()
    interval_lengths = interval_lengths.contiguous().int()
    interval_starts = interval_starts.contiguous().int()
    out = feat.new_zeros(tuple(int(s) for s in bev_feat_shape))
    ext.bev_pool_v2_forward(depth, feat, out, ranks_depth, ranks_feat, ranks_bev, interval_lengths, interval_starts)
    return out


def backward(out_grad, depth, feat, ranks_depth, ranks_feat, ranks_bev):
    """-> (depth_grad, feat_grad) following bev_pool.py:43-83 (re-sort by ranks_feat, rebuild the
    intervals, zero-filled gradients, kernel)."""
    ext = build_ref.load()
    depth = depth.contiguous().float()
    feat = feat.contiguous().float()
    order = ranks_feat.argsort()
    ranks_feat, ranks_depth, ranks_bev = ranks_feat[order], ranks_depth[order], ranks_bev[order]
    kept = torch.ones(ranks_bev.shape[0], device=ranks_bev.device, dtype=torch.bool)
    kept[1:] = ranks_feat[1:] != ranks_feat[:-1]
    interval_starts_bp = torch.where(kept)[0].int()
    interval_lengths_bp = torch.zeros_like(interval_starts_bp)
    interval_lengths_bp[:-1] = interval_starts_bp[1:] - interval_starts_bp[:-1]
    interval_lengths_bp[-1] = ranks_bev.shape[0] - interval_starts_bp[-1]
    depth_grad = depth.new_zeros(depth.shape)
    feat_grad = feat.new_zeros(feat.shape)
    ext.bev_pool_v2_backward(out_grad.contiguous(), depth_grad, feat_grad, depth, feat,
                             ranks_depth.contiguous().int(), ranks_feat.contiguous().int(),
                             ranks_bev.contiguous().int(), interval_lengths_bp.contiguous(),
                             interval_starts_bp.contiguous())
    return depth_grad, feat_grad


def bev_pool_v2(depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape, interval_starts, interval_lengths):
    """bev_pool.py:86-92."""
    x = forward(depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape, interval_starts, interval_lengths)
    return x.permute(0, 4, 1, 2, 3).contiguous()
