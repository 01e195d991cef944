"""CPU oracle for the camera->BEV pooling hot path.  TEST INFRASTRUCTURE ONLY.

This module restates, on the CPU, the algorithms of the reference
(mook0126/RCBEVDet) for the one hot path this repo accelerates.  It is the
checker the CUDA path is compared against.  It must never be imported by the
product package ``rcbevdet_b200``: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may use it.

Parity status: PINNED.
  * pool forward/backward  - pinned by the reference's only known-answer test
    (mmdet3d/ops/bev_pool_v2/bev_pool.py:145-176) in tests/test_oracle.py, and on
    the GPU box against the reference's own CUDA kernels compiled unmodified into
    oracle/_ref/ (see oracle/build_ref.py).
  * prepare (ranks/intervals) and radar RCS scatter - pinned by golden vectors
    produced by executing the reference's own Python code in the build container
    (oracle/gen_golden.py -> tests/golden/*.npz).

Every function cites the reference file:line it follows.  Integer work is
bit-exact; float work mirrors the reference's rounding sequence.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

# ---------------------------------------------------------------------------
# Row P: LSSViewTransformer.voxel_pooling_prepare_v2
#        (mmdet3d/models/necks/view_transformer.py:207-265)
# ---------------------------------------------------------------------------


def grid_infos(x, y, z):
    """create_grid_infos (view_transformer.py:67-83): fp32 values of Python-double
    expressions -> (lower_bound, interval, size), each float32[3]."""
    cfgs = [x, y, z]
    lower = np.array([c[0] for c in cfgs], dtype=np.float64).astype(np.float32)
    interval = np.array([c[2] for c in cfgs], dtype=np.float64).astype(np.float32)
    size = np.array([(c[1] - c[0]) / c[2] for c in cfgs], dtype=np.float64).astype(np.float32)
    return lower, interval, size


def voxel_index(coor, lower, interval):
    """view_transformer.py:230-232: fp32 subtract, fp32 divide (two separately
    rounded ops), then .long() == truncate toward zero."""
    coor = np.asarray(coor, dtype=np.float32)
    v = (coor - lower.astype(np.float32)) / interval.astype(np.float32)
    assert v.dtype == np.float32
    with np.errstate(invalid="ignore"):
        return np.trunc(v).astype(np.int64)


def voxel_pooling_prepare_v2(coor, lower, interval, size):
    """Restatement of view_transformer.py:207-265.

    coor: float32 (B, N, D, H, W, 3).  lower/interval/size: float32[3].
    Returns (ranks_bev, ranks_depth, ranks_feat, interval_starts,
    interval_lengths) as int32 arrays, or five ``None`` when nothing is kept
    (:258-259).  Tie order inside an interval is the STABLE one (ascending
    ranks_depth); the reference's argsort (:250) leaves it unspecified, so the
    reference output is compared after canonicalise().
    """
    coor = np.asarray(coor, dtype=np.float32)
    B, N, D, H, W, _ = coor.shape
    P = B * N * D * H * W
    # :223-228
    ranks_depth = np.arange(P, dtype=np.int32)
    ranks_feat = np.arange(P // D, dtype=np.int32).reshape(B, N, 1, H, W)
    ranks_feat = np.broadcast_to(ranks_feat, (B, N, D, H, W)).reshape(-1)
    # :230-235
    v = voxel_index(coor, lower, interval).reshape(P, 3)
    batch = np.repeat(np.arange(B, dtype=np.int64), P // B)
    # :238-240 -- int64 values compared against 0-dim fp32 tensors (promotes to fp32)
    vf = v.astype(np.float32)
    kept = ((v[:, 0] >= 0) & (vf[:, 0] < size[0]) & (v[:, 1] >= 0) & (vf[:, 1] < size[1]) &
            (v[:, 2] >= 0) & (vf[:, 2] < size[2]))
    v, batch = v[kept], batch[kept]
    ranks_depth, ranks_feat = ranks_depth[kept], ranks_feat[kept]
    # :246-249 -- int64 * 0-dim fp32 promotes to fp32: ranks_bev is an fp32 tensor
    f32 = np.float32
    ranks_bev = batch.astype(f32) * f32(f32(size[2] * size[1]) * size[0])
    ranks_bev = ranks_bev + v[:, 2].astype(f32) * f32(size[1] * size[0])
    ranks_bev = ranks_bev + (v[:, 1].astype(f32) * size[0] + v[:, 0].astype(f32))
    assert ranks_bev.dtype == np.float32
    # :250-252 (stable flavour)
    order = np.argsort(ranks_bev, kind="stable")
    ranks_bev, ranks_depth, ranks_feat = ranks_bev[order], ranks_depth[order], ranks_feat[order]
    # :254-262
    if ranks_bev.shape[0] == 0:
        return None, None, None, None, None
    first = np.ones(ranks_bev.shape[0], dtype=bool)
    first[1:] = ranks_bev[1:] != ranks_bev[:-1]
    interval_starts = np.nonzero(first)[0].astype(np.int32)
    interval_lengths = np.empty_like(interval_starts)
    interval_lengths[:-1] = interval_starts[1:] - interval_starts[:-1]
    interval_lengths[-1] = ranks_bev.shape[0] - interval_starts[-1]
    return (ranks_bev.astype(np.int32), ranks_depth.astype(np.int32), ranks_feat.astype(np.int32),
            interval_starts, interval_lengths)


def voxel_pooling_prepare_v2_c(coor, lower, interval, size, threads=1):
    """Same contract as voxel_pooling_prepare_v2 above, computed by the compiled C restatement
    (bevpool_oracle.c, stable counting sort).  Used as the timed CPU baseline."""
    coor = np.ascontiguousarray(coor, dtype=np.float32)
    B, N, D, H, W, _ = coor.shape
    P = B * N * D * H * W
    n_cells = int(B * int(size[0]) * int(size[1]) * int(size[2]))
    fp, ip = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int)
    lo, iv, sz = (np.ascontiguousarray(a, dtype=np.float32) for a in (lower, interval, size))
    rb, rd, rf = (np.empty(max(P, 1), np.int32) for _ in range(3))
    st, ln = (np.empty(max(1, min(P, n_cells)), np.int32) for _ in range(2))
    counts = np.zeros(4, np.int32)
    rc = _lib().oracle_voxel_pooling_prepare_v2(
        B, N, D, H, W, coor.ctypes.data_as(fp), lo.ctypes.data_as(fp), iv.ctypes.data_as(fp),
        sz.ctypes.data_as(fp), rb.ctypes.data_as(ip), rd.ctypes.data_as(ip), rf.ctypes.data_as(ip),
        st.ctypes.data_as(ip), ln.ctypes.data_as(ip), counts.ctypes.data_as(ip), int(threads))
    if rc != 0:
        raise RuntimeError("oracle_voxel_pooling_prepare_v2 failed")
    k, i = int(counts[0]), int(counts[1])
    if k == 0:
        return None, None, None, None, None
    return rb[:k], rd[:k], rf[:k], st[:i], ln[:i]


def canonicalise(ranks_bev, ranks_depth, ranks_feat):
    """Lexicographic (ranks_bev, ranks_depth) order.  Lossless because every
    ranks_depth value is unique; removes the reference's unspecified tie order."""
    order = np.lexsort((np.asarray(ranks_depth), np.asarray(ranks_bev)))
    return (np.asarray(ranks_bev)[order], np.asarray(ranks_depth)[order],
            np.asarray(ranks_feat)[order])


# ---------------------------------------------------------------------------
# Rows F / B: bev_pool_v2 forward / backward  (C restatement, oracle/bevpool_oracle.c)
# ---------------------------------------------------------------------------

_clib = None


def build_c(force=False):
    """gcc-compile oracle/bevpool_oracle.c -> oracle/_build/libbevpool_oracle.so"""
    out_dir = os.path.join(_HERE, "_build")
    so = os.path.join(out_dir, "libbevpool_oracle.so")
    src = os.path.join(_HERE, "bevpool_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        os.makedirs(out_dir, exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC",
                               "-o", so, src, "-lm"])
    return so


def _lib():
    global _clib
    if _clib is None:
        _clib = ctypes.CDLL(build_c())
        fp, ip = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int)
        _clib.oracle_bev_pool_v2_fwd.argtypes = [ctypes.c_int, ctypes.c_int, fp, fp, ip, ip, ip, ip, ip, fp,
                                                 ctypes.c_int]
        _clib.oracle_bev_pool_v2_fwd.restype = None
        _clib.oracle_bev_pool_v2_bwd.argtypes = [ctypes.c_int, ctypes.c_int, fp, fp, fp, ip, ip, ip, ip, ip,
                                                 fp, fp, ctypes.c_int]
        _clib.oracle_bev_pool_v2_bwd.restype = None
        _clib.oracle_voxel_pooling_prepare_v2.argtypes = [ctypes.c_int] * 5 + [fp] * 4 + [ip] * 6 + [ctypes.c_int]
        _clib.oracle_voxel_pooling_prepare_v2.restype = ctypes.c_int
    return _clib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def bev_pool_v2_forward(depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape,
                        interval_starts, interval_lengths, threads=1):
    """bev_pool_v2_kernel (bev_pool_cuda.cu:21-48) on (B,Z,Y,X,C) zeros
    (bev_pool.py:27) -> float32 (B,Z,Y,X,C), channels last, as QuickCumsumCuda.forward
    returns it.  Use to_bczyx() for what bev_pool_v2() returns (bev_pool.py:91)."""
    C = int(bev_feat_shape[-1])
    depth, pd = _f(depth)
    feat, pf = _f(feat)
    assert feat.shape[-1] == C
    rd, prd = _i(ranks_depth)
    rf, prf = _i(ranks_feat)
    rb, prb = _i(ranks_bev)
    st, pst = _i(interval_starts)
    ln, pln = _i(interval_lengths)
    out = np.zeros(tuple(int(s) for s in bev_feat_shape), dtype=np.float32)
    _lib().oracle_bev_pool_v2_fwd(C, st.shape[0], pd, pf, prd, prf, prb, pst, pln,
                                  out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), int(threads))
    return out


def to_bczyx(out_bzyxc):
    """bev_pool.py:91: x.permute(0, 4, 1, 2, 3).contiguous()."""
    return np.ascontiguousarray(np.transpose(out_bzyxc, (0, 4, 1, 2, 3)))


def bev_pool_v2_backward(out_grad, depth, feat, ranks_depth, ranks_feat, ranks_bev, threads=1):
    """QuickCumsumCuda.backward (bev_pool.py:43-83) + bev_pool_grad_kernel
    (bev_pool_cuda.cu:67-121).  out_grad is (B,Z,Y,X,C).  The re-sort by
    ranks_feat (:47-49) is done stably so the accumulation order is defined."""
    out_grad, pg = _f(out_grad)
    C = out_grad.shape[-1]
    depth, pd = _f(depth)
    feat, pf = _f(feat)
    ranks_depth = np.asarray(ranks_depth)
    ranks_feat = np.asarray(ranks_feat)
    ranks_bev = np.asarray(ranks_bev)
    order = np.argsort(ranks_feat, kind="stable")          # :47
    rf_s, rd_s, rb_s = ranks_feat[order], ranks_depth[order], ranks_bev[order]
    first = np.ones(rb_s.shape[0], dtype=bool)             # :50-52
    first[1:] = rf_s[1:] != rf_s[:-1]
    starts = np.nonzero(first)[0].astype(np.int32)         # :53
    lengths = np.empty_like(starts)                        # :54-57
    lengths[:-1] = starts[1:] - starts[:-1]
    lengths[-1] = rb_s.shape[0] - starts[-1]
    rd, prd = _i(rd_s)
    rf, prf = _i(rf_s)
    rb, prb = _i(rb_s)
    st, pst = _i(starts)
    ln, pln = _i(lengths)
    depth_grad = np.zeros_like(depth)                      # :67
    feat_grad = np.zeros_like(feat)                        # :68
    _lib().oracle_bev_pool_v2_bwd(C, st.shape[0], pg, pd, pf, prd, prf, prb, pst, pln,
                                  depth_grad.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                                  feat_grad.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), int(threads))
    return depth_grad, feat_grad


def intervals_from_sorted(ranks_bev):
    """The run-boundary construction used by both the prepare stage
    (view_transformer.py:254-262) and the reference's KAT (bev_pool.py:157-164)."""
    ranks_bev = np.asarray(ranks_bev)
    first = np.ones(ranks_bev.shape[0], dtype=bool)
    first[1:] = ranks_bev[1:] != ranks_bev[:-1]
    starts = np.nonzero(first)[0].astype(np.int32)
    lengths = np.empty_like(starts)
    lengths[:-1] = starts[1:] - starts[:-1]
    lengths[-1] = ranks_bev.shape[0] - starts[-1]
    return starts, lengths


# ---------------------------------------------------------------------------
# Row V (upstream of P): frustum + get_lidar_coor  (view_transformer.py:85-157)
# fp64 restatement used ONLY to make synthetic inputs; not a parity surface.
# ---------------------------------------------------------------------------


def create_frustum(depth_cfg, input_size, downsample):
    """view_transformer.py:85-113 (sid=False): (D, H, W, 3) float32 of (u, v, d)."""
    H_in, W_in = input_size
    Hf, Wf = H_in // downsample, W_in // downsample
    d = np.arange(depth_cfg[0], depth_cfg[1], depth_cfg[2], dtype=np.float32)
    D = d.shape[0]
    xs = np.linspace(0, W_in - 1, Wf, dtype=np.float32)
    ys = np.linspace(0, H_in - 1, Hf, dtype=np.float32)
    fr = np.empty((D, Hf, Wf, 3), dtype=np.float32)
    fr[..., 0] = xs[None, None, :]
    fr[..., 1] = ys[None, :, None]
    fr[..., 2] = d[:, None, None]
    return fr


# ---------------------------------------------------------------------------
# Row R: PointPillarsScatterRCS.forward up to (not including) the two convs
#        (mmdet3d/models/middle_encoders/pillar_scatter.py:64-104, 115-131;
#         mmdet3d/core/utils/gaussian.py:6-23, 26-55, 57-81)
# ---------------------------------------------------------------------------


def gaussian_2d(shape, sigma):
    """gaussian.py:6-23 (float64 numpy, eps-thresholded)."""
    m, n = [(ss - 1.) / 2. for ss in shape]
    y, x = np.ogrid[-m:m + 1, -n:n + 1]
    h = np.exp(-(x * x + y * y) / (2 * sigma * sigma))
    h[h < np.finfo(h.dtype).eps * h.max()] = 0
    return h


def rcs_radius(rcs):
    """pillar_scatter.py:122-126,130: int(relu(rcs[:, -2] * (x^2 + y^2)) + 1), fp32 math,
    Python int() truncation."""
    rcs = np.asarray(rcs, dtype=np.float32)
    r = rcs[:, 0] ** 2 + rcs[:, 1] ** 2
    true_rcs = np.maximum(rcs[:, -2] * r, np.float32(0))
    radius = true_rcs + np.float32(1)
    assert radius.dtype == np.float32
    return [int(v) for v in radius]


def radar_rcs_scatter(point_features, rcs, coors, batch_size, ny, nx):
    """features (B,Cin,ny,nx), heatmap (B,ny,nx), heatmap_feat (B,1,ny,nx), all fp32.
    Sequential in voxel order exactly like the reference loop (:128-131)."""
    point_features = np.asarray(point_features, dtype=np.float32)
    rcs = np.asarray(rcs, dtype=np.float32)
    coors = np.asarray(coors)
    V, Cin = point_features.shape
    # pillar_scatter.py:64-104 (overwrite scatter, per sample)
    features = np.zeros((batch_size, Cin, ny * nx), dtype=np.float32)
    for b in range(batch_size):
        m = coors[:, 0] == b
        idx = coors[m, 2].astype(np.int64) * nx + coors[m, 3].astype(np.int64)
        features[b][:, idx] = point_features[m].T
    features = features.reshape(batch_size, Cin, ny, nx)
    heatmap = np.zeros((batch_size, ny, nx), dtype=np.float32)
    heatmap_feat = np.zeros((batch_size, 1, ny, nx), dtype=np.float32)
    radius = rcs_radius(rcs)
    for i in range(V):
        b, _, y, x = (int(t) for t in coors[i])
        r = radius[i]
        # gaussian.py:26-55
        diameter = 2 * r + 1
        g = gaussian_2d((diameter, diameter), sigma=diameter / 6)
        left, right = min(x, r), min(nx - x, r + 1)
        top, bottom = min(y, r), min(ny - y, r + 1)
        mh = heatmap[b, y - top:y + bottom, x - left:x + right]
        mg = g[r - top:r + bottom, r - left:r + right].astype(np.float32)
        if min(mg.shape) > 0 and min(mh.shape) > 0:
            np.maximum(mh, mg, out=mh)
        # gaussian.py:57-81 (last writer wins)
        heatmap_feat[b, :, y - top:y + bottom, x - left:x + right] = rcs[i, -2]
    return features, heatmap, heatmap_feat
