"""Temporal alignment of pooled BEV features: BEVDepth4D.gen_grid and
shift_feature (mmdet3d/models/detectors/bevdet_rc.py:585-657), the step that follows the pooling for
every adjacent frame when `align_after_view_transfromation` is set.

The reference materialises an (n, h, w, 3, 1) grid, runs a batched 3x3 matmul over it, normalises,
and lets F.grid_sample un-normalise and sample.  Here the per-sample 3x3 pixel transform is built
from the calibration with the reference's own handful of small matrix products (same operations,
same dtype), and ONE kernel per call does the rest per output pixel (csrc/bev_shift.cu).
There is no CPU path.
"""
from __future__ import annotations

import torch

from . import _lib

__all__ = ["gen_grid_transform", "shift_feature"]


def gen_grid_transform(n, sensor2keyegos, bda, bda_adj, grid_interval, grid_lower_bound):
    """-> (n, 3, 3) float32: key-frame BEV pixel (x, y, 1) -> adjacent-frame BEV pixel.
    bevdet_rc.py:604-642 on the (n, 1, 4, 4) calibration matrices; `grid_interval`,
    `grid_lower_bound` are the view transformer's (x, y, z) grid tensors."""
    c02l0 = sensor2keyegos[0][:, 0:1, :, :]                       # :604
    c12l0 = sensor2keyegos[1][:, 0:1, :, :]                       # :607
    like = dict(dtype=c02l0.dtype, device=c02l0.device)

    def hom(m3):                                                  # :610-612
        m = torch.zeros((n, 1, 4, 4), **like)
        m[:, :, :3, :3] = m3.unsqueeze(1)
        m[:, :, 3, 3] = 1
        return m
    b4 = hom(bda)
    c02l0 = b4.matmul(c02l0)                                      # :613
    if bda_adj is not None:                                       # :614-617
        b4 = hom(bda_adj)
    c12l0 = b4.matmul(c12l0)                                      # :618
    l02l1 = c02l0.matmul(torch.inverse(c12l0))[:, 0, :, :]        # :622 (n, 4, 4)
    keep = [0, 1, 3]
    l02l1 = l02l1[:, keep, :][:, :, keep]                         # :631-633
    f2b = torch.zeros((3, 3), **like)                             # :635-641
    f2b[0, 0], f2b[1, 1] = float(grid_interval[0]), float(grid_interval[1])
    f2b[0, 2], f2b[1, 2] = float(grid_lower_bound[0]), float(grid_lower_bound[1])
    f2b[2, 2] = 1
    f2b = f2b.view(1, 3, 3)
    return torch.inverse(f2b).matmul(l02l1).matmul(f2b).float().contiguous()   # :642


class _ShiftFeature(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, tf):
        if not input.is_cuda:
            raise RuntimeError("shift_feature runs on CUDA tensors only (rcbevdet_b200 has no CPU fallback)")
        x = input.detach().contiguous().float()                   # @force_fp32 (:653)
        n, C, H, W = x.shape
        tf = tf.to(device=x.device, dtype=torch.float32).contiguous()
        out = torch.empty_like(x)
        _lib.check(_lib.lib().rcb_bev_shift_feature(_lib.ptr(x), _lib.ptr(tf), _lib.ptr(out), n, C, H, W,
                                                    x.device.index, _lib.stream_ptr(x.device)), "rcb_bev_shift_feature")
        ctx.save_for_backward(tf)
        ctx.rcb = (tuple(x.shape), input.dtype)
        return out.to(input.dtype)

    @staticmethod
    def backward(ctx, out_grad):
        (tf,) = ctx.saved_tensors
        (n, C, H, W), dtype = ctx.rcb
        g = out_grad.contiguous().float()
        gi = torch.empty((n, C, H, W), dtype=torch.float32, device=g.device)
        _lib.check(_lib.lib().rcb_bev_shift_feature_bwd(_lib.ptr(g), _lib.ptr(tf), _lib.ptr(gi), n, C, H, W,
                                                        g.device.index, _lib.stream_ptr(g.device)),
                   "rcb_bev_shift_feature_bwd")
        return gi.to(dtype), None


def shift_feature(input, sensor2keyegos, bda, bda_adj=None, grid_interval=None, grid_lower_bound=None):
    """Drop-in for BEVDepth4D.shift_feature (bevdet_rc.py:654-657): input (n, C, h, w) pooled BEV
    features of an adjacent frame, sensor2keyegos = [current, adjacent] (n, N, 4, 4), bda (n, 3, 3).
    The two grid tensors are the view transformer's (`self.img_view_transformer.grid_interval`,
    `.grid_lower_bound` in the reference, which reads them from `self`)."""
    if grid_interval is None or grid_lower_bound is None:
        raise ValueError("shift_feature needs the view transformer's grid_interval and grid_lower_bound")
    tf = gen_grid_transform(input.shape[0], sensor2keyegos, bda, bda_adj, grid_interval, grid_lower_bound)
    return _ShiftFeature.apply(input, tf)
