"""RCS-aware radar BEV scatter (row R): PointPillarsScatterRCS.forward
(mmdet3d/models/middle_encoders/pillar_scatter.py:106-135) with the per-pillar Python loop
(two .item() syncs + a numpy Gaussian per pillar) replaced by two CUDA kernels."""
from __future__ import annotations

import ctypes

import torch
from torch import nn

from . import _lib


def _desc(point_features, rcs, batch_size, ny, nx):
    d = _lib.RadarDesc()
    d.V, d.Cin, d.rcs_dim = point_features.shape[0], point_features.shape[1], rcs.shape[1]
    d.B, d.ny, d.nx = int(batch_size), int(ny), int(nx)
    return d


class _RadarScatter(torch.autograd.Function):
    @staticmethod
    def forward(ctx, point_features, rcs, coors, batch_size, ny, nx):
        if not (point_features.is_cuda and rcs.is_cuda and coors.is_cuda):
            raise RuntimeError("radar_rcs_scatter runs on CUDA tensors only (no CPU fallback)")
        dev = point_features.device
        pf = point_features.detach().contiguous().float()
        rc = rcs.detach().contiguous().float()
        co = coors.detach().int().contiguous()
        if co.dim() != 2 or co.shape[1] != 4 or co.shape[0] != pf.shape[0] or rc.shape[0] != pf.shape[0]:
            raise ValueError("expected point_features (V,C), rcs (V,R), coors (V,4) = [b, z, y, x]")
        d = _desc(pf, rc, batch_size, ny, nx)
        lib = _lib.lib()
        f32 = dict(dtype=torch.float32, device=dev)
        features = torch.empty((d.B, d.Cin, d.ny, d.nx), **f32)
        heatmap = torch.empty((d.B, d.ny, d.nx), **f32)
        heatmap_feat = torch.empty((d.B, 1, d.ny, d.nx), **f32)
        ws_bytes = lib.rcb_radar_workspace_bytes(ctypes.byref(d))
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
        _lib.check(lib.rcb_radar_rcs_scatter(ctypes.byref(d), _lib.ptr(pf), _lib.ptr(rc), _lib.ptr(co),
                                             _lib.ptr(features), _lib.ptr(heatmap), _lib.ptr(heatmap_feat),
                                             _lib.ptr(ws), ws_bytes, dev.index, _lib.stream_ptr(dev)),
                   "rcb_radar_rcs_scatter")
        ctx.save_for_backward(co)
        ctx.rcb = (d, point_features.dtype)
        ctx.mark_non_differentiable(heatmap, heatmap_feat)
        return features, heatmap, heatmap_feat

    @staticmethod
    def backward(ctx, g_features, g_heatmap, g_heatmap_feat):
        (co,) = ctx.saved_tensors
        d, dtype = ctx.rcb
        g = g_features.contiguous().float()
        out = torch.empty((d.V, d.Cin), dtype=torch.float32, device=g.device)
        _lib.check(_lib.lib().rcb_radar_scatter_bwd(ctypes.byref(d), _lib.ptr(g), _lib.ptr(co), _lib.ptr(out),
                                                    g.device.index, _lib.stream_ptr(g.device)),
                   "rcb_radar_scatter_bwd")
        return out.to(dtype), None, None, None, None, None


def radar_rcs_scatter(point_features, rcs, coors, batch_size, ny, nx):
    """-> features (B,Cin,ny,nx), heatmap (B,ny,nx), heatmap_feat (B,1,ny,nx): the three tensors
    the reference builds at pillar_scatter.py:117-131 before its two convolutions.  Only
    `features` carries a gradient (to `point_features`); the heat-maps depend on raw radar
    attributes only (the reference reads them through .item())."""
    return _RadarScatter.apply(point_features, rcs, coors, batch_size, ny, nx)


class PointPillarsScatterRCS(nn.Module):
    """Same constructor, parameters (`compress`, `rcs_att`) and forward signature as the
    reference module (pillar_scatter.py:106-135), so its checkpoints load unchanged."""

    def __init__(self, in_channels, output_shape):
        super().__init__()
        self.output_shape = output_shape
        self.ny, self.nx = output_shape[0], output_shape[1]
        self.in_channels = in_channels
        self.compress = nn.Conv2d(in_channels * 2, in_channels, 3, padding=1)
        self.rcs_att = nn.Conv2d(2, in_channels, 1)

    def forward(self, voxel_features, coors, batch_size=None):
        point_features, rcs = voxel_features
        if batch_size is None:
            batch_size = int(coors[:, 0].max().item()) + 1 if coors.numel() else 1
        features, heatmap, heatmap_feat = radar_rcs_scatter(point_features, rcs, coors, batch_size,
                                                            self.ny, self.nx)
        rcs_att = self.rcs_att(torch.cat([heatmap.unsqueeze(dim=1), heatmap_feat], dim=1))
        return self.compress(torch.cat([features, rcs_att], dim=1))
