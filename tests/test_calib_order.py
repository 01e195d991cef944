"""CPU: the fixed operation order of the fused get_lidar_coor (include/rcbevdet_b200.h,
rcb_voxel_pooling_prepare_from_calib) restated in numpy, checked against (a) rig.lidar_coor, which
it must equal bit for bit, and (b) the reference's own get_lidar_coor outputs stored in the goldens:
the coordinates differ in the last bit for 6-22 % of the floats (the reference's batched matmul
rounds in an order its BLAS chooses) but no frustum point changes its BEV cell."""
import hashlib
import json
import os

import numpy as np
import torch

from oracle import oracle
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CALIB_KEYS = ("sensor2ego", "ego2global", "intrin", "post_rot", "post_tran", "bda")


def fixed_order_coor(calib, axes):
    """numpy restatement of k_cells<true> (csrc/prepare.cu): every product and sum separately rounded."""
    import rcbevdet_b200 as rcb
    cam, bda = (t.numpy() for t in rcb.pack_calib(*calib))
    u, v, d = (t.numpy() for t in axes)
    B, N = cam.shape[:2]
    out = np.empty((B, N, d.shape[0], v.shape[0], u.shape[0], 3), np.float32)

    def dot3(m, x, y, z):
        return (m[0] * x + m[1] * y) + m[2] * z

    for b in range(B):
        for n in range(N):
            r, pt, m, t = cam[b, n, 0:9], cam[b, n, 9:12], cam[b, n, 12:21], cam[b, n, 21:24]
            a, bb, c = np.broadcast_arrays(u[None, None, :] - pt[0], v[None, :, None] - pt[1], d[:, None, None] - pt[2])
            q = [(r[3 * i] * a + r[3 * i + 1] * bb) + r[3 * i + 2] * c for i in range(3)]
            px, py, pz = q[0] * q[2], q[1] * q[2], q[2]
            e = [dot3(m[3 * i:3 * i + 3], px, py, pz) + t[i] for i in range(3)]
            for i in range(3):
                out[b, n, ..., i] = dot3(bda[b, 3 * i:3 * i + 3], e[0], e[1], e[2])
    assert out.dtype == np.float32
    return out


def test_fixed_order_equals_rig_lidar_coor_bitwise():
    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import rig
    calib = rig.camera_rig(2, input_size=(128, 352), aug_seed=5)
    axes = rcb.frustum_axes([1.0, 60.0, 2.0], (128, 352), 16)
    got = fixed_order_coor(calib, axes)
    want = rig.lidar_coor(calib, [1.0, 60.0, 2.0], (128, 352), 16).numpy()
    assert np.array_equal(got.view(np.int32), want.view(np.int32))


def test_fixed_order_keeps_every_reference_rank(golden_prepare):
    import rcbevdet_b200 as rcb
    g = golden_prepare
    for name in ("rigA", "augD"):
        calib = tuple(torch.from_numpy(g[f"{name}.{k}"]) for k in CALIB_KEYS)
        axes = rcb.frustum_axes(frustum=torch.from_numpy(g[f"{name}.frustum"]))
        coor = fixed_order_coor(calib, axes)
        ref_coor = g[f"{name}.coor"]
        assert coor.shape == ref_coor.shape
        assert float(np.abs(coor - ref_coor).max()) < 1e-4      # same geometry ...
        lo, iv, sz = g[f"{name}.lower"], g[f"{name}.interval"], g[f"{name}.size"]
        rb, rd, rf, st, ln = oracle.voxel_pooling_prepare_v2(coor, lo, iv, sz)
        assert np.array_equal(rb, g[f"{name}.raw.ranks_bev"])     # ... and identical ranks
        assert np.array_equal(rd, g[f"{name}.canon.ranks_depth"])
        assert np.array_equal(rf, g[f"{name}.canon.ranks_feat"])
        assert np.array_equal(st, g[f"{name}.raw.interval_starts"])
        assert np.array_equal(ln, g[f"{name}.raw.interval_lengths"])


def test_fixed_order_reproduces_full_size_reference_digest():
    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import rig
    with open(os.path.join(GOLDEN, "prepare_full_digest.json")) as f:
        digest = json.load(f)
    coor = fixed_order_coor(rig.camera_rig(1), rcb.frustum_axes(rig.R50_GRID["depth"], rig.R50_INPUT, 16))
    lo, iv, sz = oracle.grid_infos(rig.R50_GRID["x"], rig.R50_GRID["y"], rig.R50_GRID["z"])
    rb, rd, rf, st, ln = oracle.voxel_pooling_prepare_v2_c(coor, lo, iv, sz, threads=4)
    rb, rd, rf = oracle.canonicalise(rb, rd, rf)
    for name, arr in (("ranks_bev", rb), ("ranks_depth", rd), ("ranks_feat", rf), ("interval_starts", st),
                      ("interval_lengths", ln)):
        assert hashlib.sha256(arr.tobytes()).hexdigest() == digest["sha256"][name], name
