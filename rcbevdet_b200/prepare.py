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


def frustum_axes(depth_cfg=None, input_size=None, downsample=None, frustum=None, device=None):
    """(u [W], v [H], d [D]) float32: the axes of the reference's frustum template
    (LSSViewTransformer.create_frustum, view_transformer.py:85-113), which is separable:
    frustum[d, h, w] = (u[w], v[h], dbin[d]).  Pass either the reference module's own
    `frustum` tensor (D, H, W, 3) -- SID depth bins included -- or the three config values."""
    if frustum is not None:
        f = frustum.detach().to(torch.float32)
        u, v, d = f[0, 0, :, 0], f[0, :, 0, 1], f[:, 0, 0, 2]
    else:
        H_in, W_in = input_size
        Hf, Wf = H_in // downsample, W_in // downsample
        d = torch.arange(*depth_cfg, dtype=torch.float)                     # :98
        u = torch.linspace(0, W_in - 1, Wf, dtype=torch.float)              # :107
        v = torch.linspace(0, H_in - 1, Hf, dtype=torch.float)              # :109
    axes = tuple(t.contiguous() for t in (u, v, d))
    return tuple(t.to(device) for t in axes) if device is not None else axes


def pack_calib(sensor2ego, ego2global, cam2imgs, post_rots, post_trans, bda):
    """The argument tuple of get_lidar_coor (view_transformer.py:115) -> (cam (B,N,24), bda (B,9))
    float32 on the inputs' device: per camera inverse(post_rots) | post_trans | combine =
    sensor2ego[:3,:3] @ inverse(cam2imgs) (:147) | sensor2ego[:3,3].  The two small inverses and
    the product are taken in float64 and rounded once (the reference runs them in fp32 through
    LAPACK/cuSOLVER, whose rounding depends on the build); 48 matrices per batch, so this costs
    nothing -- do it on the host when the calibration arrives on the host.  ego2global is unused
    by the reference too."""
    del ego2global
    B, N = sensor2ego.shape[:2]
    inv_pr = torch.linalg.inv(post_rots.double()).float()
    combine = (sensor2ego[:, :, :3, :3].double() @ torch.linalg.inv(cam2imgs.double())).float()
    cam = torch.cat((inv_pr.reshape(B, N, 9), post_trans.reshape(B, N, 3).float(), combine.reshape(B, N, 9),
                     sensor2ego[:, :, :3, 3].reshape(B, N, 3).float()), -1)
    return cam.contiguous(), bda.reshape(B, 9).float().contiguous()


def _launch(desc, dev, coor=None, frustum=None):
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
    outs = (_lib.ptr(r.ranks_bev), _lib.ptr(r.ranks_depth), _lib.ptr(r.ranks_feat), _lib.ptr(r.interval_starts),
            _lib.ptr(r.interval_lengths), _lib.ptr(r.point_cell), _lib.ptr(r.cell_start), _lib.ptr(r.counts),
            _lib.ptr(ws), ws_bytes, dev.index, _lib.stream_ptr(dev))
    if coor is not None:
        _lib.check(lib.rcb_voxel_pooling_prepare_v2(ctypes.byref(desc), _lib.ptr(coor), *outs),
                   "rcb_voxel_pooling_prepare_v2")
    else:
        fd = _lib.FrustumDesc()
        fd.u, fd.v, fd.d, fd.cam, fd.bda = (t.data_ptr() for t in frustum)
        _lib.check(lib.rcb_voxel_pooling_prepare_from_calib(ctypes.byref(desc), ctypes.byref(fd), *outs),
                   "rcb_voxel_pooling_prepare_from_calib")
    r.grid = (gz, gy, gx)
    r.B, r.D, r.HW, r.H, r.P, r.n_cells = desc.B, desc.D, desc.H * desc.W, desc.H, P, n_cells
    return r


def _coor_args(coor, grid_lower_bound, grid_interval, grid_size):
    desc = _desc(coor, grid_lower_bound, grid_interval, grid_size)
    coor = coor.detach()
    if coor.dtype != torch.float32:
        coor = coor.float()
    return desc, coor.contiguous()


def _calib_args(calib, axes, grid_lower_bound, grid_interval, grid_size, device):
    cam, bda = pack_calib(*calib) if len(calib) == 6 else calib
    if device is None:
        device = cam.device if cam.is_cuda else torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("rcbevdet_b200 prepare runs on CUDA only (there is no CPU fallback)")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    cam, bda = (t.to(device=device, dtype=torch.float32, non_blocking=True).contiguous() for t in (cam, bda))
    u, v, d = (t.to(device=device, dtype=torch.float32, non_blocking=True).contiguous() for t in axes)
    B, N = cam.shape[:2]
    desc = _lib.PrepareDesc()
    desc.B, desc.N, desc.D, desc.H, desc.W = B, N, d.numel(), v.numel(), u.numel()
    desc.lower[:] = _grid3(grid_lower_bound, "grid_lower_bound")
    desc.interval[:] = _grid3(grid_interval, "grid_interval")
    desc.size[:] = _grid3(grid_size, "grid_size")
    return desc, device, (u, v, d, cam, bda)


class FrustumCells:
    """The first prepare stage alone: point_cell (BEV cell of every frustum point, -1 = outside), and
    what is needed to run the rest of the pipeline on the same input later (`launch_sort`)."""

    __slots__ = ("desc", "device", "coor", "frustum", "point_cell", "workspace", "grid", "B", "D", "H", "W", "HW",
                 "P", "n_img", "n_cells")

    def _call(self, stages, r=None):
        fd = None
        if self.frustum is not None:
            fd = _lib.FrustumDesc()
            fd.u, fd.v, fd.d, fd.cam, fd.bda = (t.data_ptr() for t in self.frustum)
        outs = [None] * 5 + [_lib.ptr(self.point_cell), None, None]
        if r is not None:
            outs = [_lib.ptr(r.ranks_bev), _lib.ptr(r.ranks_depth), _lib.ptr(r.ranks_feat), None, None,
                    _lib.ptr(self.point_cell), _lib.ptr(r.cell_start), None]
        _lib.check(_lib.lib().rcb_voxel_pooling_prepare_staged(
            ctypes.byref(self.desc), _lib.ptr(self.coor), ctypes.byref(fd) if fd is not None else None, stages, *outs,
            _lib.ptr(self.workspace), self.workspace.numel(), self.device.index, _lib.stream_ptr(self.device)),
            "rcb_voxel_pooling_prepare_staged")

    def launch_sort(self):
        """Stage 2 (the two sort kernels) on the histograms stage 1 left in the workspace: ranks and the
        dense CSR, no interval arrays (the cell-stationary forward reads the CSR)."""
        i32 = dict(dtype=torch.int32, device=self.device)
        r = PreparedRanks()
        r.ranks_bev, r.ranks_depth, r.ranks_feat = (torch.empty(self.P, **i32) for _ in range(3))
        r.cell_start = torch.empty(self.n_cells + 1, **i32)
        r.point_cell, r.interval_starts, r.interval_lengths, r.counts = self.point_cell, None, None, None
        r.grid, r.B, r.D, r.HW, r.H, r.P, r.n_cells = self.grid, self.B, self.D, self.HW, self.H, self.P, self.n_cells
        self._call(2, r)
        return r


def staged_supported(B, cells_per_sample):
    """The staged pipeline (and the launch gate) exist for the two-level sort only: <= 1024 buckets of
    <= 2^12 cells."""
    return B * -(-cells_per_sample // 4096) <= 1024


def point_cells_async(coor=None, calib=None, axes=None, grid_lower_bound=None, grid_interval=None, grid_size=None,
                      device=None):
    """Stage 1 of the prepare pipeline: BEV cell of every frustum point, from `coor` or from the
    calibration; no sort, no read-back.  None when the geometry is outside the staged pipeline's
    envelope."""
    r = FrustumCells()
    if coor is not None:
        r.desc, r.coor = _coor_args(coor, grid_lower_bound, grid_interval, grid_size)
        r.device, r.frustum = r.coor.device, None
    else:
        r.desc, r.device, r.frustum = _calib_args(calib, axes, grid_lower_bound, grid_interval, grid_size, device)
        r.coor = None
    d = r.desc
    gx, gy, gz = (int(d.size[k]) for k in range(3))
    if not staged_supported(d.B, gx * gy * gz):
        return None
    ws_bytes = _lib.lib().rcb_prepare_workspace_bytes(ctypes.byref(d))
    if ws_bytes == 0:
        return None
    r.grid = (gz, gy, gx)
    r.B, r.D, r.H, r.W, r.HW, r.n_img = d.B, d.D, d.H, d.W, d.H * d.W, d.B * d.N
    r.P, r.n_cells = d.B * d.N * d.D * d.H * d.W, d.B * gx * gy * gz
    r.point_cell = torch.empty(r.P + 4, dtype=torch.int32, device=r.device)
    r.workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=r.device)
    r._call(1)
    return r


def prepare_async(coor, grid_lower_bound, grid_interval, grid_size):
    """Launch the pipeline; nothing is read back.  Outputs have capacity P (ranks) and
    min(P, cells) (intervals); `counts` (device int32[4]) holds {n_kept, n_intervals}."""
    desc, coor = _coor_args(coor, grid_lower_bound, grid_interval, grid_size)
    return _launch(desc, coor.device, coor=coor)


def prepare_from_calib_async(calib, axes, grid_lower_bound, grid_interval, grid_size, device=None):
    """get_lidar_coor fused into prepare.  `calib` = the six tensors of
    get_lidar_coor's signature (or an already packed (cam, bda) pair from pack_calib), `axes` =
    frustum_axes(...).  Host tensors are packed on the host and only ~5 KB are uploaded; `coor`
    (12 bytes per frustum point) is never materialised."""
    desc, device, frustum = _calib_args(calib, axes, grid_lower_bound, grid_interval, grid_size, device)
    return _launch(desc, device, frustum=frustum)


def _finish(r):
    n_kept, n_iv = r.counts[:2].tolist()  # the one host sync (the reference has four)
    if n_kept == 0 or n_iv == 0:
        return None, None, None, None, None
    out = (r.ranks_bev[:n_kept], r.ranks_depth[:n_kept], r.ranks_feat[:n_kept],
           r.interval_starts[:n_iv], r.interval_lengths[:n_iv])
    p = _plan.PoolPlan(_lib.PLAN_ALL, r.cell_start, r.point_cell, r.D, r.HW, r.n_cells, r.P)
    _plan.attach(p, out[1], out[2], out[0], out[3], out[4])
    return out


def voxel_pooling_prepare_from_calib(calib, axes, grid_lower_bound, grid_interval, grid_size, device=None):
    """get_lidar_coor + voxel_pooling_prepare_v2 (view_transformer.py:115-157, 207-265) in one
    pipeline: same 5-tuple as voxel_pooling_prepare_v2."""
    return _finish(prepare_from_calib_async(calib, axes, grid_lower_bound, grid_interval, grid_size, device))


def voxel_pooling_prepare_v2(coor, grid_lower_bound, grid_interval, grid_size):
    """-> (ranks_bev, ranks_depth, ranks_feat, interval_starts, interval_lengths), the reference's
    order (view_transformer.py:263-265), or (None,)*5 when nothing is kept (:258-259).

    Order inside an interval: by (ranks_feat, depth bin) -- the depth bins of one camera pixel
    adjacent; the reference's argsort (:250) leaves it unspecified."""
    return _finish(prepare_async(coor, grid_lower_bound, grid_interval, grid_size))


def install(view_transformer_cls, fuse_lidar_coor=False):
    """Make `view_transformer_cls.voxel_pooling_prepare_v2` (LSSViewTransformer and subclasses)
    run on this library; `view_transform`, `voxel_pooling_v2`, `init_acceleration_v2` etc. then
    work unchanged on top of it.

    fuse_lidar_coor=True additionally replaces the non-accelerated branch of `view_transform_core`
    (view_transformer.py:290-294: get_lidar_coor -> voxel_pooling_v2) by the calibration-driven
    chain, which never materialises `coor` and needs no host synchronisation."""

    def _method(self, coor):
        return voxel_pooling_prepare_v2(coor, self.grid_lower_bound, self.grid_interval, self.grid_size)

    _method.__name__ = "voxel_pooling_prepare_v2"
    _method.__doc__ = voxel_pooling_prepare_v2.__doc__
    view_transformer_cls.voxel_pooling_prepare_v2 = _method
    if fuse_lidar_coor:
        from .view_pool import voxel_pooling_v2_from_calib
        reference_core = view_transformer_cls.view_transform_core

        def _core(self, input, depth, tran_feat):
            if self.accelerate:
                return reference_core(self, input, depth, tran_feat)
            B, N, C, H, W = input[0].shape
            axes = frustum_axes(frustum=self.frustum)
            bev = voxel_pooling_v2_from_calib(tuple(input[1:7]), axes, depth.view(B, N, self.D, H, W),
                                              tran_feat.view(B, N, self.out_channels, H, W),
                                              self.grid_lower_bound, self.grid_interval, self.grid_size,
                                              collapse_z=self.collapse_z)
            return bev, depth

        _core.__name__ = "view_transform_core"
        view_transformer_cls.view_transform_core = _core
    return view_transformer_cls
