"""Strip plans: the second-level plan of the pooling kernels (csrc/strips.cu).

A strip is a vertical run of 16 pixels of one image column; its frustum points walk along one line
of BEV cells, so grouping them by (cell, strip) leaves 9-13x fewer partial rows to move than the
(cell, pixel) pairs of the cell-stationary kernels.  The plan depends only on the ranks (i.e. on the
calibration), like the ranks themselves: it is built once next to them (`voxel_pooling_prepare_v2`,
`init_acceleration_v2` in the reference's terms, view_transformer.py:159-178) and reused by every
forward / backward that runs on those ranks.
"""
from __future__ import annotations

import ctypes
import os

import torch

from . import _lib

FWD_CHANNELS = (64, 80, 128)
BWD_CHANNELS = (64, 80)

# Which pooling kernels bev_pool_v2 / voxel_pooling_v2 run on structured ranks:
#   "off" (default): cell-/pixel-stationary kernels only
#   "auto": the forward switches to the strip kernels from the SECOND forward that sees the same
#           ranks (cached ranks: the reference's accelerate mode, init_acceleration_v2 -- building the
#           plan costs about as much as one forward, so it pays from the second use on); backward
#           unchanged.  The two kernel families sum in different orders (both within rel 1e-5 of the
#           reference, each bit-reproducible): with "auto" the first and the second call on the same
#           inputs differ in the last bits, which is why it is not the default.
#   "on"  : strip kernels for forward and backward, plan built at first use
#   "chain": bev_pool_v2 as "off"; voxel_pooling_v2 / voxel_pooling_v2_from_calib / lss_view_transform --
#           the chains that never hand ranks to their caller -- pool WITHOUT SORTING: frustum cells ->
#           strip plan (it needs point_cell only) -> strip forward, and behind it, gated on the plan's
#           status word on the device, the sorted pipeline + cell-stationary forward as the fallback
#           (no read-back: exactly one family does the work); backward on the pixel-stationary kernel
#           (point_cell only).  view_pool.py.
# Environment: RCB_STRIPS=off|auto|on|chain.
MODE = os.environ.get("RCB_STRIPS", "off").lower()


def set_mode(mode):
    global MODE
    if mode not in ("auto", "on", "off", "chain"):
        raise ValueError("mode must be 'auto', 'on', 'off' or 'chain'")
    MODE = mode


class StripPlan:
    """desc: rcb_strip_desc; buf: the device plan buffer; status(): 0 when the strip kernels may run."""

    __slots__ = ("desc", "buf", "cell_start", "_status")

    def __init__(self, desc, buf, cell_start):
        self.desc, self.buf, self.cell_start = desc, buf, cell_start
        self._status = None

    def status_tensor(self):
        return self.buf[:4].view(torch.int32)

    def status(self):
        """Host copy of the status word (one 4-byte read-back, cached)."""
        if self._status is None:
            self._status = int(self.status_tensor().item())
        return self._status

    def rows(self, C):
        """Scratch for the segment rows of one call: reserved address space from torch's caching
        allocator (stream-ordered, so calls on different streams do not share it); only the rows of
        existing segments are ever touched."""
        n = _lib.lib().rcb_strip_rows_bytes(ctypes.byref(self.desc), C)
        return torch.empty(n, dtype=torch.uint8, device=self.buf.device)


def supported(n_img, D, H, W, n_cells):
    d = _lib.StripDesc()
    d.n_img, d.D, d.H, d.W, d.n_cells = n_img, D, H, W, n_cells
    return _lib.lib().rcb_strip_plan_bytes(ctypes.byref(d)) > 0


def build(point_cell, cell_start, n_img, D, H, W, n_cells):
    """Launch the plan kernels on the current stream (no read-back).  Returns None when the geometry is
    outside the strip kernels' envelope (D > 256, more than 2^24 cells)."""
    d = _lib.StripDesc()
    d.n_img, d.D, d.H, d.W, d.n_cells = int(n_img), int(D), int(H), int(W), int(n_cells)
    lib = _lib.lib()
    n = lib.rcb_strip_plan_bytes(ctypes.byref(d))
    if n == 0:
        return None
    dev = point_cell.device
    buf = torch.empty(n, dtype=torch.uint8, device=dev)
    # cell_start is no longer read by the plan (it builds its per-cell lists from its own counts)
    _lib.check(lib.rcb_strip_plan_build(ctypes.byref(d), _lib.ptr(point_cell), _lib.ptr(cell_start), _lib.ptr(buf),
                                        n, dev.index, _lib.stream_ptr(dev)), "rcb_strip_plan_build")
    return StripPlan(d, buf, cell_start)


def forward(plan, pool_desc, depth, rows_feat, out):
    ws = plan.rows(pool_desc.C)
    dev = out.device
    _lib.check(_lib.lib().rcb_bev_pool_v2_fwd_strips(
        ctypes.byref(pool_desc), ctypes.byref(plan.desc), _lib.ptr(plan.buf),
        _lib.ptr(depth), _lib.ptr(rows_feat), _lib.ptr(out), _lib.ptr(ws), ws.numel(), dev.index,
        _lib.stream_ptr(dev)), "rcb_bev_pool_v2_fwd_strips")


def backward(plan, pool_desc, out_grad, depth, rows_feat, depth_grad, feat_grad):
    ws = plan.rows(pool_desc.C)
    dev = out_grad.device
    _lib.check(_lib.lib().rcb_bev_pool_v2_bwd_strips(
        ctypes.byref(pool_desc), ctypes.byref(plan.desc), _lib.ptr(plan.buf),
        _lib.ptr(out_grad), _lib.ptr(depth), _lib.ptr(rows_feat), _lib.ptr(depth_grad), _lib.ptr(feat_grad),
        _lib.ptr(ws), ws.numel(), dev.index, _lib.stream_ptr(dev)), "rcb_bev_pool_v2_bwd_strips")


def _for_plan(plan, desc):
    """The StripPlan of a PoolPlan (built on first request; one 4-byte read-back of its status), or None
    when the strip kernels cannot take these ranks."""
    if plan.strips is None:
        ok = (plan.structured and plan.sorted_cells and desc.D > 0 and desc.H > 0 and desc.HW > 0
              and desc.HW % desc.H == 0 and desc.n_pixels % desc.HW == 0
              and desc.n_pixels * desc.D == desc.n_depth)
        if ok and torch.cuda.is_current_stream_capturing():
            return None                                     # the status read-back cannot be captured
        sp = build(plan.point_cell, plan.cell_start, desc.n_pixels // desc.HW, desc.D, desc.H,
                   desc.HW // desc.H, plan.n_cells) if ok else None
        plan.strips = sp if (sp is not None and sp.status() == 0) else False
    return plan.strips or None


def for_forward(plan, desc):
    if MODE in ("off", "chain") or desc.C not in FWD_CHANNELS:
        return None
    plan.uses += 1
    if MODE == "auto":
        if plan.uses < 2 and plan.strips is None:
            return None
        # one or two samples on a small grid: the cell-stationary kernel in its small-batch CTA shapes is
        # as fast or faster (R50 grid, B = 1 / 2: 22 / 35 us against 26 / 35 us on strips; on the 256^2
        # grid the strips win from the first sample on: 85 against 116 us)
        if desc.B <= 2 and desc.Z * desc.Y * desc.X < 32768 and plan.strips is None:
            return None
    return _for_plan(plan, desc)


def for_backward(plan, desc):
    if MODE != "on" or desc.C not in BWD_CHANNELS:
        return None
    return _for_plan(plan, desc)
