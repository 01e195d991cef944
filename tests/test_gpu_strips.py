"""Strip-stationary pooling kernels (csrc/strips.cu) through the public operators, against the CPU
oracle and against the cell-/pixel-stationary kernels.  fp32: rel 1e-5; bf16 context: 1e-2."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import oracle
from test_gpu_pool import RTOL32, _case, _close, _oracle_pool

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _mode():
    from rcbevdet_b200 import strips
    old = strips.MODE
    yield
    strips.set_mode(old)


def _pool(rcb, mode, coor, depth, feat, grid, shape, og=None, dtype=torch.float32, channels_last=False):
    from rcbevdet_b200 import rig, strips, plan as _plan
    strips.set_mode(mode)
    lo, iv, sz = rig.grid_tensors(grid)
    rb, rd, rf, st, ln = rcb.voxel_pooling_prepare_v2(coor.cuda(), lo, iv, sz)
    d = depth.cuda().requires_grad_(True)
    f = feat.cuda().to(dtype).requires_grad_(True)
    bev = rcb.bev_pool_v2(d, f.permute(0, 1, 3, 4, 2), rd, rf, rb, shape, st, ln, channels_last=channels_last)
    plan = _plan.lookup(rd, rf, rb, st, ln, shape[0] * shape[1] * shape[2] * shape[3], depth.numel())
    if og is not None:
        bev.backward(og.cuda())
    return bev.detach(), d.grad, f.grad, plan


@pytest.mark.parametrize("C,depth_cfg,input_size,B,aug", [
    (80, (1.0, 60.0, 2.0), (128, 352), 2, 3),
    (80, (1.0, 60.0, 0.5), (128, 352), 1, None),   # the R50 depth bins (D = 118)
    (64, (1.0, 60.0, 1.0), (112, 304), 2, 4),      # H = 7, W = 19: ragged strips and column groups
    (128, (1.0, 60.0, 2.0), (112, 208), 2, None),  # forward only for C = 128 (backward takes the pixel kernel)
    (80, (1.0, 60.0, 1.0), (288, 352), 1, 5),      # H = 18: two strips per image column
])
def test_strips_against_oracle(C, depth_cfg, input_size, B, aug):
    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import rig, strips
    grid = rig.R50_GRID
    coor, depth, feat = _case(B=B, depth_cfg=depth_cfg, input_size=input_size, C=C, aug=aug, seed=C + B)
    ranks, shape, feat_rows, want = _oracle_pool(coor, depth, feat, grid)
    og = torch.randn((shape[0], shape[4], shape[1], shape[2], shape[3]), generator=torch.Generator().manual_seed(C))
    bev, dg, fg, plan = _pool(rcb, "on", coor, depth, feat, grid, shape, og)
    assert isinstance(plan.strips, strips.StripPlan) and plan.strips.status() == 0
    _close(bev, oracle.to_bczyx(want), RTOL32, "bev")
    og_rows = og.permute(0, 2, 3, 4, 1).contiguous().numpy()
    want_dg, want_fg = oracle.bev_pool_v2_backward(og_rows, depth.numpy(), feat_rows, ranks[1], ranks[2], ranks[0], threads=8)
    _close(dg, want_dg, RTOL32, "depth_grad")
    _close(fg.permute(0, 1, 3, 4, 2), want_fg, RTOL32, "feat_grad")
    # dropped points get exactly zero depth gradient, cells without points exactly zero features
    kept = np.zeros(depth.numel(), bool)
    kept[ranks[1]] = True
    assert float(dg.flatten()[torch.from_numpy(~kept).cuda()].abs().max()) == 0.0
    empty = np.ones(shape[0] * shape[1] * shape[2] * shape[3], bool)
    empty[ranks[0]] = False
    assert float(bev.permute(0, 2, 3, 4, 1).reshape(-1, C)[torch.from_numpy(empty).cuda()].abs().max()) == 0.0


def test_strips_channels_last_bf16_and_reproducible():
    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import rig
    grid = rig.R50_GRID
    coor, depth, feat = _case(B=2, aug=2)
    ranks, shape, feat_rows, want = _oracle_pool(coor, depth, feat, grid)
    og = torch.randn((shape[0], shape[4], shape[1], shape[2], shape[3]), generator=torch.Generator().manual_seed(1))
    a = _pool(rcb, "on", coor, depth, feat, grid, shape, og, channels_last=True)
    b = _pool(rcb, "on", coor, depth, feat, grid, shape, og, channels_last=True)
    assert a[0].permute(0, 2, 3, 4, 1).is_contiguous()
    _close(a[0], oracle.to_bczyx(want), RTOL32, "bev channels-last")
    for x, y in zip(a[:3], b[:3]):
        assert torch.equal(x, y), "strip kernels must be bit-reproducible"
    h = _pool(rcb, "on", coor, depth, feat, grid, shape, og, dtype=torch.bfloat16)
    fb = feat.to(torch.bfloat16).float()
    _, _, rows_b, want_b = _oracle_pool(coor, depth, fb, grid)
    _close(h[0], oracle.to_bczyx(want_b), RTOL32, "bev, bf16 context widened exactly")


def test_auto_mode_switches_on_reuse():
    """Cached ranks (the reference's accelerate mode): the second forward on the same ranks builds the
    strip plan and runs the strip kernels; the first one does not pay for a plan.  One or two samples on
    a small grid stay on the cell-stationary kernel (its small-batch CTA shapes are as fast there)."""
    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import rig, strips, plan as _plan
    strips.set_mode("auto")
    grid = rig.R50_GRID
    for B, switches in ((3, True), (1, False)):
        coor, depth, feat = _case(B=B, depth_cfg=(1.0, 60.0, 0.5), input_size=(128, 352))
        ranks, shape, feat_rows, want = _oracle_pool(coor, depth, feat, grid)
        lo, iv, sz = rig.grid_tensors(grid)
        rb, rd, rf, st, ln = rcb.voxel_pooling_prepare_v2(coor.cuda(), lo, iv, sz)
        plan = _plan.lookup(rd, rf, rb, st, ln, shape[0] * shape[1] * shape[2] * shape[3], depth.numel())
        with torch.no_grad():
            first = rcb.bev_pool_v2(depth.cuda(), feat.cuda().permute(0, 1, 3, 4, 2), rd, rf, rb, shape, st, ln)
            assert plan.strips is None
            second = rcb.bev_pool_v2(depth.cuda(), feat.cuda().permute(0, 1, 3, 4, 2), rd, rf, rb, shape, st, ln)
            assert isinstance(plan.strips, strips.StripPlan) == switches
        _close(first, oracle.to_bczyx(want), RTOL32, "cells kernel")
        _close(second, oracle.to_bczyx(want), RTOL32, "second call")
        assert torch.equal(first, second) != switches


def test_foreign_int64_ranks_take_the_strip_kernels_too():
    """Ranks that did not come from this library's prepare (the oracle's, as int64 like the reference's
    own): validated on the device once, then planned for the strip kernels like any other."""
    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import rig, strips, plan as _plan
    strips.set_mode("on")
    grid = rig.R50_GRID
    coor, depth, feat = _case(B=2, aug=6)
    ranks, shape, feat_rows, want = _oracle_pool(coor, depth, feat, grid)
    rb, rd, rf, st, ln = (torch.from_numpy(np.ascontiguousarray(r)).long().cuda() for r in ranks)
    d = depth.cuda().requires_grad_(True)
    f = feat.cuda().requires_grad_(True)
    bev = rcb.bev_pool_v2(d, f.permute(0, 1, 3, 4, 2), rd, rf, rb, shape, st, ln)
    plan = _plan.lookup(rd, rf, rb, st, ln, shape[0] * shape[1] * shape[2] * shape[3], depth.numel())
    assert isinstance(plan.strips, strips.StripPlan)
    _close(bev, oracle.to_bczyx(want), RTOL32, "bev")
    og = torch.randn(bev.shape, generator=torch.Generator().manual_seed(9))
    bev.backward(og.cuda())
    og_rows = og.permute(0, 2, 3, 4, 1).contiguous().numpy()
    want_dg, want_fg = oracle.bev_pool_v2_backward(og_rows, depth.numpy(), feat_rows, ranks[1], ranks[2], ranks[0], threads=8)
    _close(d.grad, want_dg, RTOL32, "depth_grad")
    _close(f.grad.permute(0, 1, 3, 4, 2), want_fg, RTOL32, "feat_grad")


def test_non_frustum_ranks_fall_back():
    """Points scattered at random (no ray geometry): a strip meets more cells than the plan reserves,
    or a (cell, pixel) pair has several depth runs -> the plan refuses, the general kernels run."""
    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import rig
    grid = rig.R50_GRID
    g = torch.Generator().manual_seed(11)
    B, N, D, H, W, C = 1, 2, 40, 8, 12, 80
    coor = torch.rand(B, N, D, H, W, 3, generator=g) * torch.tensor([110.0, 110.0, 9.0]) - torch.tensor([55.0, 55.0, 5.5])
    depth, feat = rig.pooling_inputs(B, N, D, H, W, C, seed=3)
    ranks, shape, feat_rows, want = _oracle_pool(coor, depth, feat, grid)
    og = torch.randn((shape[0], shape[4], shape[1], shape[2], shape[3]), generator=torch.Generator().manual_seed(2))
    bev, dg, fg, plan = _pool(rcb, "on", coor, depth, feat, grid, shape, og)
    assert plan.strips is False
    _close(bev, oracle.to_bczyx(want), RTOL32, "bev (fallback)")
    og_rows = og.permute(0, 2, 3, 4, 1).contiguous().numpy()
    want_dg, want_fg = oracle.bev_pool_v2_backward(og_rows, depth.numpy(), feat_rows, ranks[1], ranks[2], ranks[0], threads=8)
    _close(dg, want_dg, RTOL32, "depth_grad (fallback)")
    _close(fg.permute(0, 1, 3, 4, 2), want_fg, RTOL32, "feat_grad (fallback)")


@pytest.mark.parametrize("aug", [None, 3])
def test_strips_full_size_r50_against_cell_kernels(aug):
    """BASELINE config 2 at its real size (B = 8): strip kernels against the cell-/pixel-stationary
    kernels (themselves pinned to the oracle and to the reference's own kernels at this size)."""
    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import rig
    grid = rig.R50_GRID
    B = 8
    coor = rig.lidar_coor(rig.camera_rig(B, aug_seed=aug), grid["depth"], rig.R50_INPUT, 16)
    _, N, D, H, W, _ = coor.shape
    depth, feat = rig.pooling_inputs(B, N, D, H, W, 80, seed=5)
    shape = (B, 1, 128, 128, 80)
    og = torch.randn((B, 80, 1, 128, 128), generator=torch.Generator().manual_seed(4))
    ref = _pool(rcb, "off", coor, depth, feat, grid, shape, og)
    got = _pool(rcb, "on", coor, depth, feat, grid, shape, og)
    assert got[3].strips and ref[3].strips is None
    for name, x, y in zip(("bev", "depth_grad", "feat_grad"), got[:3], ref[:3]):
        _close(x, y.cpu().numpy(), RTOL32, name)


def test_strips_hires_against_cell_kernels():
    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import rig
    grid = rig.HIRES_GRID
    coor = rig.lidar_coor(rig.camera_rig(1, input_size=rig.HIRES_INPUT, aug_seed=3), grid["depth"], rig.HIRES_INPUT, 16)
    _, N, D, H, W, _ = coor.shape
    depth, feat = rig.pooling_inputs(1, N, D, H, W, 80, seed=6)
    shape = (1, 1, 256, 256, 80)
    og = torch.randn((1, 80, 1, 256, 256), generator=torch.Generator().manual_seed(4))
    ref = _pool(rcb, "off", coor, depth, feat, grid, shape, og)
    got = _pool(rcb, "on", coor, depth, feat, grid, shape, og)
    if got[3].strips is False:
        pytest.skip("this augmentation puts more segments into a group of strips than the plan reserves")
    for name, x, y in zip(("bev", "depth_grad", "feat_grad"), got[:3], ref[:3]):
        _close(x, y.cpu().numpy(), RTOL32, name)


def test_c_abi_refuses_what_it_cannot_do():
    from rcbevdet_b200 import _lib
    lib = _lib.lib()
    sd = _lib.StripDesc()
    sd.n_img, sd.D, sd.H, sd.W, sd.n_cells = 6, 300, 16, 44, 16384      # D > 256
    assert lib.rcb_strip_plan_bytes(ctypes.byref(sd)) == 0
    sd.D = 118
    assert lib.rcb_strip_plan_bytes(ctypes.byref(sd)) > 0
    d = _lib.PoolDesc()
    d.n_points, d.n_intervals, d.C = 10, 0, 12                          # channel count without a strip kernel
    d.B, d.Z, d.Y, d.X = 1, 1, 128, 128
    d.n_depth, d.n_pixels, d.D, d.HW, d.H = 6 * 118 * 704, 6 * 704, 118, 704, 16
    d.layout, d.feat_dtype, d.flags = _lib.LAYOUT_B_C_CELLS, _lib.DTYPE_F32, _lib.PLAN_ALL
    x = torch.zeros(16, device="cuda")
    rc = lib.rcb_bev_pool_v2_fwd_strips(ctypes.byref(d), ctypes.byref(sd), _lib.ptr(x), _lib.ptr(x), _lib.ptr(x),
                                        _lib.ptr(x), _lib.ptr(x), 64, 0, None)
    assert rc == -3   # RCB_ERR_UNSUPPORTED


@pytest.mark.parametrize("input_size,depth_cfg,B,aug", [((112, 304), (1.0, 60.0, 1.0), 2, 4), ((288, 352), (1.0, 60.0, 2.0), 1, 5)])
def test_plan_and_rows_stay_inside_their_buffers(input_size, depth_cfg, B, aug):
    """Guard bands around the plan buffer and the segment-row scratch of the C-ABI calls (ragged strips
    and column groups): the plan kernels (k_strip_plan, k_seg_prefix / _fill / _assign) and the strip
    forward write nothing outside the extents rcb_strip_plan_bytes / rcb_strip_rows_bytes declare."""
    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import _lib, bev_pool as bp, rig
    from rcbevdet_b200.prepare import prepare_async
    grid = rig.R50_GRID
    coor, depth, feat = _case(B=B, depth_cfg=depth_cfg, input_size=input_size, C=80, aug=aug, seed=3)
    lo, iv, sz = rig.grid_tensors(grid)
    r = prepare_async(coor.cuda(), lo, iv, sz)
    _, N, D, H, W, _ = coor.shape
    sd = _lib.StripDesc()
    sd.n_img, sd.D, sd.H, sd.W, sd.n_cells = B * N, D, H, W, r.n_cells
    lib = _lib.lib()
    G = 4096
    n_plan = lib.rcb_strip_plan_bytes(ctypes.byref(sd))
    n_rows = lib.rcb_strip_rows_bytes(ctypes.byref(sd), 80)
    plan_all = torch.full((n_plan + 2 * G,), 0x5a, dtype=torch.uint8, device="cuda")
    rows_all = torch.full((n_rows + 2 * G,), 0x5a, dtype=torch.uint8, device="cuda")
    plan, rows_ws = plan_all[G:G + n_plan], rows_all[G:G + n_rows]
    _lib.check(lib.rcb_strip_plan_build(ctypes.byref(sd), _lib.ptr(r.point_cell), None, _lib.ptr(plan), n_plan, 0, None), "plan")
    assert int(plan[:4].view(torch.int32).item()) == 0
    dc = depth.cuda()
    frows = feat.permute(0, 1, 3, 4, 2).contiguous().view(-1, 80).cuda()
    d = _lib.PoolDesc()
    d.n_points, d.n_intervals, d.C = r.P, 0, 80
    d.B, d.Z, d.Y, d.X = B, 1, 128, 128
    d.n_depth, d.n_pixels, d.D, d.HW, d.H = dc.numel(), frows.shape[0], D, H * W, H
    d.layout, d.feat_dtype, d.flags = _lib.LAYOUT_B_C_CELLS, _lib.DTYPE_F32, _lib.PLAN_ALL
    out = torch.empty((B, 80, 1, 128, 128), device="cuda")
    _lib.check(lib.rcb_bev_pool_v2_fwd_strips(ctypes.byref(d), ctypes.byref(sd), _lib.ptr(plan), _lib.ptr(dc), _lib.ptr(frows),
                                              _lib.ptr(out), _lib.ptr(rows_ws), n_rows, 0, None), "fwd_strips")
    torch.cuda.synchronize()
    for name, full, n in (("plan", plan_all, n_plan), ("rows", rows_all, n_rows)):
        assert int((full[:G] != 0x5a).sum()) == 0 and int((full[G + n:] != 0x5a).sum()) == 0, name
    ref = torch.empty_like(out)
    bp.pool_forward(d, dc, frows, r.ranks_depth, r.ranks_feat, r.ranks_bev, None, None, r.cell_start, ref)
    _close(out, ref.cpu().numpy(), RTOL32, "strip forward through the guarded buffers")
