"""ctypes binding of librcbevdet_b200.so (the C ABI declared in include/rcbevdet_b200.h).

There is no fallback: if the CUDA library is missing every operator of this package raises.
"""
from __future__ import annotations

import contextlib
import ctypes
import os
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
# RCB_LIB_PATH selects another build of the SAME library (kernel experiments, tools/build_variant.sh)
LIB_PATH = os.environ.get("RCB_LIB_PATH") or os.path.join(_PKG, "lib", "librcbevdet_b200.so")

RCB_OK = 0
DTYPE_F32, DTYPE_BF16, DTYPE_F16 = 0, 1, 2
LAYOUT_CELLS_C, LAYOUT_B_C_CELLS = 0, 1
PLAN_RANGES_OK, PLAN_INTERVALS_OK, PLAN_SORTED_CELLS, PLAN_STRUCTURED, PLAN_ALL = 1, 2, 4, 8, 15


class PrepareDesc(ctypes.Structure):
    _fields_ = [("B", ctypes.c_int), ("N", ctypes.c_int), ("D", ctypes.c_int), ("H", ctypes.c_int),
                ("W", ctypes.c_int), ("lower", ctypes.c_float * 3), ("interval", ctypes.c_float * 3),
                ("size", ctypes.c_float * 3)]


class FrustumDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("u", "v", "d", "cam", "bda")]


class PoolDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in
                ("n_points", "n_intervals", "C", "B", "Z", "Y", "X", "n_depth", "n_pixels", "D", "HW",
                 "H", "layout", "feat_dtype", "flags")]


class StripDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("n_img", "D", "H", "W", "n_cells")]


class RadarDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("V", "Cin", "rcs_dim", "B", "ny", "nx")]


_vp, _i, _sz, _ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_longlong

# name -> (restype, argtypes); mirrors include/rcbevdet_b200.h one to one
SIGNATURES = {
    "rcb_version": (_i, []),
    "rcb_error_string": (ctypes.c_char_p, [_i]),
    "rcb_device_info": (_i, [_i] + [ctypes.POINTER(ctypes.c_int)] * 4),
    "rcb_prepare_workspace_bytes": (_sz, [ctypes.POINTER(PrepareDesc)]),
    "rcb_voxel_pooling_prepare_v2": (_i, [ctypes.POINTER(PrepareDesc)] + [_vp] * 9 + [_vp, _sz, _i, _vp]),
    "rcb_voxel_pooling_prepare_from_calib": (_i, [ctypes.POINTER(PrepareDesc), ctypes.POINTER(FrustumDesc)] +
                                             [_vp] * 8 + [_vp, _sz, _i, _vp]),
    "rcb_frustum_point_cells": (_i, [ctypes.POINTER(PrepareDesc), _vp, ctypes.POINTER(FrustumDesc), _vp, _i, _vp]),
    "rcb_set_launch_gate": (_i, [_vp]),
    "rcb_voxel_pooling_prepare_staged": (_i, [ctypes.POINTER(PrepareDesc), _vp, ctypes.POINTER(FrustumDesc), _i] +
                                         [_vp] * 8 + [_vp, _sz, _i, _vp]),
    "rcb_debug_exactdiv_sweep": (_i, [ctypes.c_float, ctypes.POINTER(ctypes.c_ulonglong),
                                      ctypes.POINTER(ctypes.c_ulonglong), _i]),
    "rcb_pool_validate_workspace_bytes": (_sz, [ctypes.POINTER(PoolDesc)]),
    "rcb_pool_validate": (_i, [ctypes.POINTER(PoolDesc)] + [_vp] * 7 + [_vp, _sz, _i, _vp]),
    "rcb_pool_build_cellmap": (_i, [ctypes.POINTER(PoolDesc), _vp, _vp, _vp, _i, _vp]),
    "rcb_bev_pool_v2_fwd": (_i, [ctypes.POINTER(PoolDesc)] + [_vp] * 9 + [_i, _vp]),
    "rcb_pool_bwd_workspace_bytes": (_sz, [ctypes.POINTER(PoolDesc)]),
    "rcb_bev_pool_v2_bwd": (_i, [ctypes.POINTER(PoolDesc)] + [_vp] * 9 + [_vp, _sz, _i, _vp]),
    "rcb_strip_plan_bytes": (_sz, [ctypes.POINTER(StripDesc)]),
    "rcb_strip_rows_bytes": (_sz, [ctypes.POINTER(StripDesc), _i]),
    "rcb_strip_plan_build": (_i, [ctypes.POINTER(StripDesc), _vp, _vp, _vp, _sz, _i, _vp]),
    "rcb_bev_pool_v2_fwd_strips": (_i, [ctypes.POINTER(PoolDesc), ctypes.POINTER(StripDesc)] + [_vp] * 5 + [_sz, _i, _vp]),
    "rcb_bev_pool_v2_bwd_strips": (_i, [ctypes.POINTER(PoolDesc), ctypes.POINTER(StripDesc)] + [_vp] * 7 + [_sz, _i, _vp]),
    "rcb_planes_to_rows": (_i, [_vp, _vp, _i, _i, _i, _ll, _i, _i, _vp]),
    "rcb_radar_workspace_bytes": (_sz, [ctypes.POINTER(RadarDesc)]),
    "rcb_radar_rcs_scatter": (_i, [ctypes.POINTER(RadarDesc)] + [_vp] * 6 + [_vp, _sz, _i, _vp]),
    "rcb_radar_scatter_bwd": (_i, [ctypes.POINTER(RadarDesc), _vp, _vp, _vp, _i, _vp]),
    "rcb_bev_shift_feature": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "rcb_bev_shift_feature_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "rcb_depth_context_split": (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _i, _ll, _i, _vp]),
    "rcb_trt_bev_pool_v2_workspace_bytes": (_sz, [_i, _i, _i]),
    "rcb_trt_bev_pool_v2_enqueue": (_i, [ctypes.POINTER(_vp), ctypes.POINTER(_vp)] + [_i] * 11 + [_vp, _sz, _i, _vp]),
    "rcb_depth_context_split_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
}

_lib = None
_lock = threading.Lock()


def lib():
    """The loaded library.  Raises RuntimeError (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} is missing: build it with `python -m rcbevdet_b200.build` "
                        "(nvcc, sm_100a).  rcbevdet_b200 has no CPU or PyTorch fallback.")
                handle = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(handle, name)
                    fn.restype, fn.argtypes = res, args
                _lib = handle
    return _lib


def check(code, what):
    if code != RCB_OK:
        msg = lib().rcb_error_string(int(code)).decode()
        raise RuntimeError(f"{what} failed: {msg} (code {int(code)})")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr(device):
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


@contextlib.contextmanager
def launch_gate(word):
    """Inside the block the general entry points (prepare, cell-stationary forward, pixel-stationary
    backward) launch kernels that exit at once when the device int32 `word[0]` is zero
    (rcb_set_launch_gate); None = no gate."""
    if word is None:
        yield
        return
    lib().rcb_set_launch_gate(ptr(word))
    try:
        yield
    finally:
        lib().rcb_set_launch_gate(None)
