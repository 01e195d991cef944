"""The sort-free chain (strips mode "chain", rcbevdet_b200/view_pool.py): voxel_pooling_v2 /
voxel_pooling_v2_from_calib pooled by the strip kernels on a plan built from point_cell alone, with the
sorted pipeline + cell-/pixel-stationary kernels enqueued behind them, gated on the plan's status word.
Checked against the CPU oracle (fp32 rel 1e-5) and against the default chain."""
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


def _chain(rcb, mode, coor, depth, feat, grid, og, calib=None, axes=None, channels_last=False):
    from rcbevdet_b200 import rig, strips
    strips.set_mode(mode)
    lo, iv, sz = rig.grid_tensors(grid)
    d = depth.cuda().requires_grad_(True)
    f = feat.cuda().requires_grad_(True)
    if calib is not None:
        bev = rcb.voxel_pooling_v2_from_calib(calib, axes, d, f, lo, iv, sz, collapse_z=False, channels_last=channels_last)
    else:
        bev = rcb.voxel_pooling_v2(coor.cuda(), d, f, lo, iv, sz, collapse_z=False, channels_last=channels_last)
    bev.backward(og.cuda())
    return bev.detach(), d.grad, f.grad


def _oracle_grads(og, depth, feat_rows, ranks):
    og_rows = og.permute(0, 2, 3, 4, 1).contiguous().numpy()
    return oracle.bev_pool_v2_backward(og_rows, depth.numpy(), feat_rows, ranks[1], ranks[2], ranks[0], threads=8)


@pytest.mark.parametrize("C,depth_cfg,input_size,B,aug", [
    (80, (1.0, 60.0, 2.0), (128, 352), 2, 3),
    (80, (1.0, 60.0, 0.5), (128, 352), 1, None),   # the R50 depth bins (D = 118)
    (64, (1.0, 60.0, 1.0), (112, 304), 2, 4),      # H = 7, W = 19: ragged strips and column groups
    (128, (1.0, 60.0, 2.0), (112, 208), 2, None),
    (80, (1.0, 60.0, 1.0), (288, 352), 1, 5),      # H = 18: two strips per image column
])
def test_chain_against_oracle(C, depth_cfg, input_size, B, aug):
    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import rig
    grid = rig.R50_GRID
    coor, depth, feat = _case(B=B, depth_cfg=depth_cfg, input_size=input_size, C=C, aug=aug, seed=C + B)
    ranks, shape, feat_rows, want = _oracle_pool(coor, depth, feat, grid)
    og = torch.randn((shape[0], shape[4], shape[1], shape[2], shape[3]), generator=torch.Generator().manual_seed(C))
    bev, dg, fg = _chain(rcb, "chain", coor, depth, feat, grid, og)
    _close(bev, oracle.to_bczyx(want), RTOL32, "bev")
    want_dg, want_fg = _oracle_grads(og, depth, feat_rows, ranks)
    _close(dg, want_dg, RTOL32, "depth_grad")
    _close(fg.permute(0, 1, 3, 4, 2), want_fg, RTOL32, "feat_grad")
    # the strip kernels did the forward (same bits as the strip kernels behind bev_pool_v2), not the
    # fallback; the backward is the pixel-stationary kernel's (same bits as the default chain's)
    from test_gpu_strips import _pool
    on = _pool(rcb, "on", coor, depth, feat, grid, shape, og)
    assert on[3].strips and on[3].strips.status() == 0
    assert torch.equal(bev, on[0])
    off = _chain(rcb, "off", coor, depth, feat, grid, og)
    assert torch.equal(dg, off[1]) and torch.equal(fg, off[2])


def test_chain_falls_back_on_the_device():
    """Points scattered at random (no ray geometry): the plan refuses, its status word opens the gate
    and the sorted pipeline + cell-/pixel-stationary kernels -- enqueued behind the strip kernels
    without any read-back -- produce the result: bit-identical to the default chain."""
    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import rig
    grid = rig.R50_GRID
    g = torch.Generator().manual_seed(11)
    B, N, D, H, W, C = 1, 2, 40, 8, 12, 80
    coor = torch.rand(B, N, D, H, W, 3, generator=g) * torch.tensor([110.0, 110.0, 9.0]) - torch.tensor([55.0, 55.0, 5.5])
    depth, feat = rig.pooling_inputs(B, N, D, H, W, C, seed=3)
    ranks, shape, feat_rows, want = _oracle_pool(coor, depth, feat, grid)
    og = torch.randn((shape[0], shape[4], shape[1], shape[2], shape[3]), generator=torch.Generator().manual_seed(2))
    ref = _chain(rcb, "off", coor, depth, feat, grid, og)
    got = _chain(rcb, "chain", coor, depth, feat, grid, og)
    _close(got[0], oracle.to_bczyx(want), RTOL32, "bev (fallback)")
    for name, x, y in zip(("bev", "depth_grad", "feat_grad"), got, ref):
        assert torch.equal(x, y), name


def test_launch_gate_closes_and_opens():
    """rcb_set_launch_gate: with the word at zero the general kernels leave their outputs untouched,
    with a non-zero word they run."""
    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import _lib, rig
    grid = rig.R50_GRID
    coor, depth, feat = _case(B=1)
    lo, iv, sz = rig.grid_tensors(grid)
    word = torch.zeros(4, dtype=torch.int32, device="cuda")
    want = rcb.voxel_pooling_v2(coor.cuda(), depth.cuda(), feat.cuda(), lo, iv, sz)
    torch.cuda.synchronize()
    from rcbevdet_b200 import view_pool
    real_empty = torch.empty

    def poisoned(*a, **k):
        t = real_empty(*a, **k)
        if t.dtype == torch.float32:
            t.fill_(-7.0)
        return t

    for value, ran in ((0, False), (5, True)):
        word[0] = value
        torch.empty = poisoned
        try:
            with _lib.launch_gate(word):
                got = rcb.voxel_pooling_v2(coor.cuda(), depth.cuda(), feat.cuda(), lo, iv, sz)
        finally:
            torch.empty = real_empty
        if ran:
            assert torch.equal(got, want)
        else:
            assert float((got + 7.0).abs().max()) == 0.0, "gated kernels must not write"
    # entry points whose kernels take no gate say so
    d = _lib.PoolDesc()
    d.n_points, d.n_intervals, d.C, d.B, d.Z, d.Y, d.X = 4, 1, 12, 1, 1, 4, 4
    d.n_depth, d.n_pixels, d.layout, d.feat_dtype = 16, 4, _lib.LAYOUT_CELLS_C, _lib.DTYPE_F32
    buf = torch.zeros(1024, device="cuda")
    ib = torch.zeros(64, dtype=torch.int32, device="cuda")
    with _lib.launch_gate(word):
        rc = _lib.lib().rcb_bev_pool_v2_fwd(ctypes.byref(d), _lib.ptr(buf), _lib.ptr(buf), _lib.ptr(ib), _lib.ptr(ib),
                                            _lib.ptr(ib), _lib.ptr(ib), _lib.ptr(ib), None, _lib.ptr(buf), 0, None)
    assert rc == -3, rc   # RCB_ERR_UNSUPPORTED


def test_chain_from_calib_full_size_r50():
    """BASELINE config 2 at its real size (B = 8), calibration-driven: the sort-free chain against the
    default chain (itself pinned to the oracle and to the reference's kernels at this size);
    bit-reproducible; channels-last result."""
    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import rig
    B, C = 8, 80
    grid = rig.R50_GRID
    calib = rig.camera_rig(B, aug_seed=3)
    axes = rcb.frustum_axes(grid["depth"], rig.R50_INPUT, 16)
    depth, feat = rig.pooling_inputs(B, 6, 118, 16, 44, C, seed=4)
    og = torch.randn(B, C, 1, 128, 128, generator=torch.Generator().manual_seed(9))
    ref = _chain(rcb, "off", None, depth, feat, grid, og, calib=calib, axes=axes)
    got = _chain(rcb, "chain", None, depth, feat, grid, og, calib=calib, axes=axes)
    again = _chain(rcb, "chain", None, depth, feat, grid, og, calib=calib, axes=axes)
    for name, x, y, z in zip(("bev", "depth_grad", "feat_grad"), got, ref, again):
        _close(x, y.cpu().numpy(), RTOL32, name)
        assert torch.equal(x, z), f"{name}: the chain must be bit-reproducible"
    assert not torch.equal(got[0], ref[0]), "same bits as the cell kernels: did the strip kernels run at all?"
    cl = _chain(rcb, "chain", None, depth, feat, grid, og, calib=calib, axes=axes, channels_last=True)
    assert cl[0].permute(0, 2, 3, 4, 1).is_contiguous()
    for x, y in zip(cl, got):
        assert torch.equal(x, y)


def test_chain_is_cuda_graph_capturable():
    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import rig, strips
    strips.set_mode("chain")
    B = 1
    axes = rcb.frustum_axes(rig.R50_GRID["depth"], rig.R50_INPUT, 16, device="cuda")
    cam, bda = (t.cuda() for t in rcb.pack_calib(*rig.camera_rig(B)))
    lo, iv, sz = rig.grid_tensors(rig.R50_GRID)
    depth, feat = (t.cuda() for t in rig.pooling_inputs(B, 6, 118, 16, 44, 80, seed=1))
    og = torch.randn(B, 80, 128, 128, device="cuda", generator=torch.Generator("cuda").manual_seed(2))

    def step():
        d = depth.detach().requires_grad_(True)
        f = feat.detach().requires_grad_(True)
        bev = rcb.voxel_pooling_v2_from_calib((cam, bda), axes, d, f, lo, iv, sz)
        bev.backward(og)
        return bev.detach(), d.grad, f.grad

    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(2):
            step()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        captured = step()
    depth2, feat2 = (t.cuda() for t in rig.pooling_inputs(B, 6, 118, 16, 44, 80, seed=5))
    cam2, bda2 = (t.cuda() for t in rcb.pack_calib(*rig.camera_rig(B, aug_seed=3)))
    depth.copy_(depth2), feat.copy_(feat2), cam.copy_(cam2), bda.copy_(bda2)
    graph.replay()
    torch.cuda.synchronize()
    got = [t.clone() for t in captured]
    want = step()
    for a, b in zip(got, want):
        assert torch.equal(a, b)


def test_lss_view_transform_takes_the_chain():
    """lss_view_transform (fused producer -> pool) in mode "chain" against the same call in mode "off":
    values and the gradient w.r.t. the depth-net output, rel 1e-5."""
    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import rig, strips
    B, N, D, C = 2, 6, 118, 80
    calib = rig.camera_rig(B)
    axes = rcb.frustum_axes(rig.R50_GRID["depth"], rig.R50_INPUT, 16)
    lo, iv, sz = rig.grid_tensors(rig.R50_GRID)
    x = torch.randn(B * N, D + C, 16, 44, device="cuda", generator=torch.Generator("cuda").manual_seed(7))
    og = torch.randn(B, C, 128, 128, device="cuda", generator=torch.Generator("cuda").manual_seed(8))
    res = {}
    for mode in ("off", "chain"):
        strips.set_mode(mode)
        xa = x.clone().requires_grad_(True)
        bev, depth = rcb.lss_view_transform(xa, N, D, C, calib, axes, lo, iv, sz)
        bev.backward(og)
        res[mode] = (bev.detach(), xa.grad)
    for a, b, what in zip(res["chain"], res["off"], ("bev", "grad")):
        assert float((a - b).abs().max()) <= 1e-5 * float(b.abs().max()), what
    assert not torch.equal(res["chain"][0], res["off"][0])   # another summation order: the strip kernels ran


def test_point_cells_and_staged_prepare_entries():
    """rcb_frustum_point_cells == the point_cell of the full pipeline; rcb_voxel_pooling_prepare_staged
    (stage 1, then stages 2 | 4) == the one-call pipeline, bit for bit."""
    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import _lib, rig
    from rcbevdet_b200.prepare import _coor_args, prepare_async
    coor, _, _ = _case(B=2, aug=3)
    lo, iv, sz = rig.grid_tensors(rig.R50_GRID)
    ref = prepare_async(coor.cuda(), lo, iv, sz)
    desc, c = _coor_args(coor.cuda(), lo, iv, sz)
    lib = _lib.lib()
    P = coor.numel() // 3
    pc = torch.full((P + 4,), -7, dtype=torch.int32, device="cuda")
    _lib.check(lib.rcb_frustum_point_cells(ctypes.byref(desc), _lib.ptr(c), None, _lib.ptr(pc), 0, None), "point_cells")
    assert torch.equal(pc[:P], ref.point_cell[:P])
    ws = torch.empty(lib.rcb_prepare_workspace_bytes(ctypes.byref(desc)), dtype=torch.uint8, device="cuda")
    i32 = dict(dtype=torch.int32, device="cuda")
    rb, rd, rf = (torch.empty(P, **i32) for _ in range(3))
    n_cells = ref.n_cells
    st, ln = torch.empty(n_cells, **i32), torch.empty(n_cells, **i32)
    cs, counts, pc2 = torch.empty(n_cells + 1, **i32), torch.empty(4, **i32), torch.empty(P + 4, **i32)
    tail = (_lib.ptr(ws), ws.numel(), 0, None)
    _lib.check(lib.rcb_voxel_pooling_prepare_staged(ctypes.byref(desc), _lib.ptr(c), None, 1, None, None, None, None, None,
                                                    _lib.ptr(pc2), None, None, *tail), "stage 1")
    _lib.check(lib.rcb_voxel_pooling_prepare_staged(ctypes.byref(desc), None, None, 6, _lib.ptr(rb), _lib.ptr(rd),
                                                    _lib.ptr(rf), _lib.ptr(st), _lib.ptr(ln), _lib.ptr(pc2), _lib.ptr(cs),
                                                    _lib.ptr(counts), *tail), "stages 2 | 4")
    k, n_iv = (int(v) for v in counts[:2].tolist())
    assert [k, n_iv] == ref.counts[:2].tolist()
    for a, b in ((rb[:k], ref.ranks_bev[:k]), (rd[:k], ref.ranks_depth[:k]), (rf[:k], ref.ranks_feat[:k]),
                 (st[:n_iv], ref.interval_starts[:n_iv]), (ln[:n_iv], ref.interval_lengths[:n_iv]), (cs, ref.cell_start)):
        assert torch.equal(a, b)


@pytest.mark.parametrize("aug", [None, 3])
def test_chain_hires_against_default(aug):
    """BASELINE config 5's geometry (900 x 1600 -> 56 x 100 features, 256^2 BEV), one sample: the sort-free
    chain against the default chain, rel 1e-5.  Without augmentation the strip plan holds (the bits
    differ from the cell kernel's); with it the plan may refuse and the gated fallback answers."""
    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import rig
    grid = rig.HIRES_GRID
    coor = rig.lidar_coor(rig.camera_rig(1, input_size=rig.HIRES_INPUT, aug_seed=aug), grid["depth"], rig.HIRES_INPUT, 16)
    _, N, D, H, W, _ = coor.shape
    depth, feat = rig.pooling_inputs(1, N, D, H, W, 80, seed=6)
    og = torch.randn((1, 80, 1, 256, 256), generator=torch.Generator().manual_seed(4))
    ref = _chain(rcb, "off", coor, depth, feat, grid, og)
    got = _chain(rcb, "chain", coor, depth, feat, grid, og)
    for name, x, y in zip(("bev", "depth_grad", "feat_grad"), got, ref):
        _close(x, y.cpu().numpy(), RTOL32, name)
    if aug is None:
        assert not torch.equal(got[0], ref[0]), "the strip kernels should have pooled this one"
