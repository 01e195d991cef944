"""Pool plans: what the CUDA kernels need to know about a set of ranks beyond the five arrays
of the reference's signature.

A plan is (a) attached to the tensors `voxel_pooling_prepare_v2` returns, where every property
holds by construction, or (b) derived on the device by `rcb_pool_validate` for ranks that came
from somewhere else (the reference's own prepare, hand-written test vectors), and cached while
those tensors are alive and unmodified.  The reference has no such notion: its extension trusts
its inputs (mmdet3d/ops/bev_pool_v2/src/bev_pool.cpp:30-57).
"""
from __future__ import annotations

import collections
import ctypes
import threading

import torch

from . import _lib

_ATTR = "_rcb_plan"


class PoolPlan:
    """flags: RCB_PLAN_* bits.  cell_start: int32 [n_cells + 1] dense CSR over BEV cells (None
    unless SORTED_CELLS).  point_cell: int32 [n_depth] BEV cell of each depth element, -1 if
    unused (None unless STRUCTURED).  D / HW describe the frustum when STRUCTURED."""

    __slots__ = ("flags", "cell_start", "point_cell", "D", "HW", "n_cells", "n_depth", "keys", "hold", "ranks",
                 "strips", "uses")

    def __init__(self, flags, cell_start, point_cell, D, HW, n_cells, n_depth, keys=None, hold=None,
                 ranks=None):
        self.flags = int(flags)
        self.cell_start = cell_start
        self.point_cell = point_cell
        self.D = int(D)
        self.HW = int(HW)
        self.n_cells = int(n_cells)
        self.n_depth = int(n_depth)
        self.keys = keys
        self.hold = hold
        # the five rank arrays as int32 contiguous tensors (what the kernels read); for ranks that
        # arrive as int64 / strided views these are converted ONCE and kept here, so the cache hits
        # on the caller's own tensors instead of on a fresh copy per call
        self.ranks = ranks
        # second-level plan of the strip kernels (rcbevdet_b200/strips.py): None = not built yet,
        # False = built and refused (or geometry outside the envelope); `uses` counts forwards
        self.strips = None
        self.uses = 0

    @property
    def sorted_cells(self):
        return bool(self.flags & _lib.PLAN_SORTED_CELLS) and self.cell_start is not None

    @property
    def structured(self):
        return bool(self.flags & _lib.PLAN_STRUCTURED) and self.point_cell is not None


def _key(t):
    # tensors created under torch.inference_mode() do not track a version counter (reading
    # `_version` raises); they are immutable outside inference mode, so 0 is as good a version
    version = 0 if t.is_inference() else t._version
    return (t.data_ptr(), t.numel(), version, t.dtype, t.device, tuple(t.stride()))


def attach(plan, ranks_depth, ranks_feat, ranks_bev, interval_starts, interval_lengths):
    plan.keys = tuple(_key(t) for t in (ranks_depth, ranks_feat, ranks_bev, interval_starts, interval_lengths))
    plan.ranks = (ranks_depth, ranks_feat, ranks_bev, interval_starts, interval_lengths)
    for t in (ranks_depth, ranks_bev):
        setattr(t, _ATTR, plan)


_cache = collections.OrderedDict()
_cache_lock = threading.Lock()
_CACHE_MAX = 4


def clear_cache():
    with _cache_lock:
        _cache.clear()


def lookup(ranks_depth, ranks_feat, ranks_bev, interval_starts, interval_lengths, n_cells, n_depth):
    keys = tuple(_key(t) for t in (ranks_depth, ranks_feat, ranks_bev, interval_starts, interval_lengths))
    for t in (ranks_depth, ranks_bev):
        plan = getattr(t, _ATTR, None)
        if plan is not None and plan.keys == keys and plan.n_cells == n_cells and plan.n_depth == n_depth:
            return plan
    with _cache_lock:
        plan = _cache.get((keys, n_cells, n_depth))
        if plan is not None:
            _cache.move_to_end((keys, n_cells, n_depth))
    return plan


def _as_kernel_ranks(t):
    import torch as _t
    return t if (t.dtype == _t.int32 and t.is_contiguous()) else t.int().contiguous()  # bev_pool.py:18,22-25


def derive(desc, originals):
    """Validate foreign ranks on the device (one flag read-back) and cache the result under the
    identity of the caller's ORIGINAL tensors (ranks_depth, ranks_feat, ranks_bev, interval_starts,
    interval_lengths); int64 / strided ranks are converted to int32 once and kept in the plan.

    The cache key is (data_ptr, numel, version counter, dtype, strides): ranks must not be modified
    behind the version counter (`.data.copy_`, external kernels) while a plan is cached -- call
    clear_cache() after such a write."""
    ranks_depth, ranks_feat, ranks_bev, interval_starts, interval_lengths = (
        _as_kernel_ranks(t) for t in originals)
    dev = ranks_depth.device if ranks_depth.numel() else ranks_bev.device
    n_cells = desc.B * desc.Z * desc.Y * desc.X
    lib = _lib.lib()
    flags_dev = torch.empty(1, dtype=torch.int32, device=dev)
    point_cell = torch.empty(max(desc.n_depth, 1), dtype=torch.int32, device=dev)
    ws_bytes = lib.rcb_pool_validate_workspace_bytes(ctypes.byref(desc))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    st = _lib.stream_ptr(dev)
    _lib.check(lib.rcb_pool_validate(ctypes.byref(desc), _lib.ptr(ranks_depth), _lib.ptr(ranks_feat),
                                     _lib.ptr(ranks_bev), _lib.ptr(interval_starts),
                                     _lib.ptr(interval_lengths), _lib.ptr(point_cell), _lib.ptr(flags_dev),
                                     _lib.ptr(ws), ws_bytes, dev.index, st), "rcb_pool_validate")
    flags = int(flags_dev.item())
    if not flags & _lib.PLAN_RANGES_OK:
        raise RuntimeError("bev_pool_v2: a rank indexes outside its tensor (ranks_depth < depth.numel(), "
                           "ranks_feat < feat rows, ranks_bev < B*Z*Y*X are required)")
    cell_start = None
    if flags & _lib.PLAN_SORTED_CELLS:
        cell_start = torch.empty(n_cells + 1, dtype=torch.int32, device=dev)
        _lib.check(lib.rcb_pool_build_cellmap(ctypes.byref(desc), _lib.ptr(ranks_bev),
                                              _lib.ptr(interval_starts), _lib.ptr(cell_start), dev.index, st),
                   "rcb_pool_build_cellmap")
    if not flags & _lib.PLAN_STRUCTURED:
        point_cell = None
    tensors = (ranks_depth, ranks_feat, ranks_bev, interval_starts, interval_lengths)
    keys = tuple(_key(t) for t in originals)
    plan = PoolPlan(flags, cell_start, point_cell, desc.D, desc.HW, n_cells, desc.n_depth, keys=keys,
                    hold=tuple(originals), ranks=tensors)  # holding the tensors keeps their addresses from being reused
    with _cache_lock:
        _cache[(keys, n_cells, desc.n_depth)] = plan
        while len(_cache) > _CACHE_MAX:
            _cache.popitem(last=False)
    return plan
