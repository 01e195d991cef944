"""CUDA bev_pool_v2 forward/backward (through the C ABI) vs the CPU oracle.
fp32: rel 1e-5 (BASELINE.json north_star); bf16/fp16 context: 1e-2."""
import numpy as np
import pytest
import torch

from oracle import oracle
from rank_check import assert_same_ranks

pytestmark = pytest.mark.gpu

RTOL32 = 1e-5


def _close(got, want, rtol, what):
    got = got.detach().cpu().numpy() if isinstance(got, torch.Tensor) else got
    scale = max(float(np.abs(want).max()), 1e-30)
    err = float(np.abs(got - want).max())
    assert err <= rtol * scale, f"{what}: max abs err {err:.3e} > {rtol} * {scale:.3e}"


def _case(B=2, depth_cfg=(1.0, 60.0, 2.0), input_size=(128, 352), C=80, aug=None, seed=1):
    from rcbevdet_b200 import rig
    coor = rig.lidar_coor(rig.camera_rig(B, input_size=input_size, aug_seed=aug), list(depth_cfg), input_size, 16)
    _, N, D, H, W, _ = coor.shape
    depth, feat = rig.pooling_inputs(B, N, D, H, W, C, seed=seed)
    return coor, depth, feat


def _oracle_pool(coor, depth, feat_nchw, grid):
    from rcbevdet_b200 import rig
    lo, iv, sz = (t.numpy() for t in rig.grid_tensors(grid))
    rb, rd, rf, st, ln = oracle.voxel_pooling_prepare_v2(coor.numpy(), lo, iv, sz)
    B = coor.shape[0]
    C = feat_nchw.shape[2]
    shape = (B, int(sz[2]), int(sz[1]), int(sz[0]), C)
    feat_rows = feat_nchw.permute(0, 1, 3, 4, 2).contiguous().float().numpy()
    out = oracle.bev_pool_v2_forward(depth.numpy(), feat_rows, rd, rf, rb, shape, st, ln, threads=8)
    return (rb, rd, rf, st, ln), shape, feat_rows, out


def test_reference_known_answer_test():
    """mmdet3d/ops/bev_pool_v2/bev_pool.py:145-176, verbatim values."""
    import rcbevdet_b200 as rcb
    depth = torch.tensor([0.3, 0.4, 0.2, 0.1, 0.7, 0.6, 0.8, 0.9], device="cuda").view(1, 1, 2, 2, 2)
    depth.requires_grad_(True)
    feat = torch.ones(1, 1, 2, 2, 2, device="cuda", requires_grad=True)
    rd = torch.tensor([0, 4, 1, 6], dtype=torch.int32, device="cuda")
    rf = torch.tensor([0, 0, 1, 2], dtype=torch.int32, device="cuda")
    rb = torch.tensor([0, 0, 1, 1], dtype=torch.int32, device="cuda")
    st = torch.tensor([0, 2], dtype=torch.int32, device="cuda")
    ln = torch.tensor([2, 2], dtype=torch.int32, device="cuda")
    bev = rcb.bev_pool_v2(depth, feat, rd, rf, rb, (1, 1, 2, 2, 2), st, ln)
    assert bev.shape == (1, 2, 1, 2, 2) and bev.is_contiguous()
    loss = bev.sum()
    loss.backward()
    assert loss.item() == pytest.approx(4.4, abs=1e-6)
    assert torch.allclose(depth.grad.flatten().cpu(), torch.tensor([2., 2., 0., 0., 2., 0., 2., 0.]))
    assert torch.allclose(feat.grad.flatten().cpu(), torch.tensor([1.0, 1.0, 0.4, 0.4, 0.8, 0.8, 0., 0.]))


@pytest.mark.parametrize("aug", [None, 3])
def test_prepare_then_pool_forward_backward(aug):
    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import rig
    grid = rig.R50_GRID
    coor, depth, feat = _case(aug=aug)
    ranks, shape, feat_rows, want = _oracle_pool(coor, depth, feat, grid)
    lo, iv, sz = rig.grid_tensors(grid)
    g_ranks = rcb.voxel_pooling_prepare_v2(coor.cuda(), lo, iv, sz)
    assert_same_ranks(g_ranks, ranks)
    rb, rd, rf, st, ln = g_ranks
    d = depth.cuda().requires_grad_(True)
    f = feat.cuda().requires_grad_(True)
    fview = f.permute(0, 1, 3, 4, 2)                 # non-contiguous, as view_transformer.py:195 passes it
    assert not fview.is_contiguous()
    bev = rcb.bev_pool_v2(d, fview, rd, rf, rb, shape, st, ln)
    assert bev.shape == (shape[0], shape[4], shape[1], shape[2], shape[3]) and bev.is_contiguous()
    _close(bev, oracle.to_bczyx(want), RTOL32, "bev")
    # untouched cells are exactly zero
    empty = np.ones(shape[0] * shape[1] * shape[2] * shape[3], bool)
    empty[ranks[0]] = False
    assert float(bev.detach().permute(0, 2, 3, 4, 1).reshape(-1, shape[4])[torch.from_numpy(empty).cuda()].abs().max()) == 0.0
    # backward
    gen = torch.Generator().manual_seed(3)
    og = torch.randn(bev.shape, generator=gen)
    bev.backward(og.cuda())
    og_rows = og.permute(0, 2, 3, 4, 1).contiguous().numpy()
    dg, fg = oracle.bev_pool_v2_backward(og_rows, depth.numpy(), feat_rows, ranks[1], ranks[2], ranks[0], threads=8)
    _close(d.grad, dg, RTOL32, "depth_grad")
    _close(f.grad.permute(0, 1, 3, 4, 2), fg, RTOL32, "feat_grad")
    # dropped points / unseen pixels get exactly zero gradient
    kept = np.zeros(depth.numel(), bool)
    kept[ranks[1]] = True
    assert float(d.grad.flatten()[torch.from_numpy(~kept).cuda()].abs().max()) == 0.0


@pytest.mark.parametrize("C,depth_cfg,input_size,B", [
    (80, (1.0, 60.0, 0.35), (128, 352), 1),   # D = 169: the backward stages its depth columns in two chunks
    (64, (1.0, 60.0, 1.0), (112, 208), 1),    # one quad per lane, 546 pixels: a partial 16-pixel group
    (128, (1.0, 60.0, 2.0), (112, 208), 2),   # two quads per lane
    (96, (1.0, 60.0, 2.0), (128, 352), 1),    # channel count without a half-warp kernel: 32-lane fallback
    (80, (1.0, 8.0, 0.5), (112, 208), 1),     # D = 14 < 16: a single partial batch
])
def test_backward_kernel_variants(C, depth_cfg, input_size, B):
    """Every structured-backward code path (channel templates, chunked depth staging, ragged pixel
    groups) and the forward for the same channel counts, against the oracle."""
    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import rig
    grid = rig.R50_GRID
    coor, depth, feat = _case(B=B, depth_cfg=depth_cfg, input_size=input_size, C=C, seed=C)
    ranks, shape, feat_rows, want = _oracle_pool(coor, depth, feat, grid)
    if ranks[0] is None:
        pytest.skip("geometry keeps no point")
    lo, iv, sz = rig.grid_tensors(grid)
    rb, rd, rf, st, ln = rcb.voxel_pooling_prepare_v2(coor.cuda(), lo, iv, sz)
    d = depth.cuda().requires_grad_(True)
    f = feat.cuda().requires_grad_(True)
    bev = rcb.bev_pool_v2(d, f.permute(0, 1, 3, 4, 2), rd, rf, rb, shape, st, ln)
    _close(bev, oracle.to_bczyx(want), RTOL32, "bev")
    og = torch.randn(bev.shape, generator=torch.Generator().manual_seed(C + 1))
    bev.backward(og.cuda())
    og_rows = og.permute(0, 2, 3, 4, 1).contiguous().numpy()
    dg, fg = oracle.bev_pool_v2_backward(og_rows, depth.numpy(), feat_rows, ranks[1], ranks[2], ranks[0], threads=8)
    _close(d.grad, dg, RTOL32, "depth_grad")
    _close(f.grad.permute(0, 1, 3, 4, 2), fg, RTOL32, "feat_grad")


def test_quickcumsum_channels_last_and_determinism():
    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import rig
    coor, depth, feat = _case(B=1, C=64)
    ranks, shape, feat_rows, want = _oracle_pool(coor, depth, feat, rig.R50_GRID)
    rb, rd, rf, st, ln = (torch.from_numpy(r).cuda() for r in ranks)   # foreign ranks -> validated plan
    f = feat.permute(0, 1, 3, 4, 2).contiguous().cuda()
    out1 = rcb.QuickCumsumCuda.apply(depth.cuda(), f, rd, rf, rb, shape, st, ln)
    out2 = rcb.QuickCumsumCuda.apply(depth.cuda(), f, rd, rf, rb, shape, st, ln)
    assert out1.shape == shape
    _close(out1, want, RTOL32, "channels-last out")
    assert torch.equal(out1, out2)                       # bit-reproducible


def test_int64_ranks_and_fp16_inputs_are_accepted():
    """The reference casts ranks with .int() and inputs with .float() (bev_pool.py:18-25)."""
    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import rig
    coor, depth, feat = _case(B=1, C=32, depth_cfg=(1.0, 60.0, 4.0), input_size=(64, 176))
    ranks, shape, feat_rows, _ = _oracle_pool(coor, depth, feat, rig.R50_GRID)
    rb, rd, rf, st, ln = (torch.from_numpy(r).long().cuda() for r in ranks)
    d16, f16 = depth.half(), feat.half()
    want = oracle.bev_pool_v2_forward(d16.float().numpy(), f16.float().permute(0, 1, 3, 4, 2).contiguous().numpy(),
                                      ranks[1], ranks[2], ranks[0], shape, ranks[3], ranks[4])
    bev = rcb.bev_pool_v2(d16.cuda(), f16.cuda().permute(0, 1, 3, 4, 2), rd, rf, rb, shape, st, ln)
    assert bev.dtype == torch.float32
    _close(bev, oracle.to_bczyx(want), RTOL32, "fp16 inputs, fp32 accumulate")


def test_bf16_context_within_1e2():
    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import rig
    coor, depth, feat = _case(B=1)
    ranks, shape, feat_rows, want = _oracle_pool(coor, depth, feat, rig.R50_GRID)
    rb, rd, rf, st, ln = (torch.from_numpy(r).cuda() for r in ranks)
    f = feat.cuda().bfloat16().requires_grad_(True)
    d = depth.cuda().requires_grad_(True)
    bev = rcb.bev_pool_v2(d, f.permute(0, 1, 3, 4, 2), rd, rf, rb, shape, st, ln)
    _close(bev, oracle.to_bczyx(want), 1e-2, "bf16 context")
    bev.sum().backward()
    assert f.grad.dtype == torch.bfloat16 and d.grad.dtype == torch.float32


def test_general_path_unsorted_unstructured_ranks():
    """Anything the reference accepts must work: intervals in arbitrary cell order, repeated
    ranks_depth, ranks_feat unrelated to ranks_depth."""
    import rcbevdet_b200 as rcb
    rng = np.random.default_rng(0)
    n_cells, n_depth, n_pix, C, K = 4 * 6 * 5, 500, 37, 12, 900
    cells = rng.permutation(n_cells)[:40]
    lens = rng.multinomial(K - 40, np.ones(40) / 40) + 1
    rb = np.repeat(cells, lens).astype(np.int32)
    rd = rng.choice(n_depth, K, replace=True).astype(np.int32)
    rf = rng.integers(0, n_pix, K).astype(np.int32)
    st, ln = oracle.intervals_from_sorted(rb)
    depth = rng.random(n_depth, dtype=np.float32)
    feat = rng.standard_normal((n_pix, C)).astype(np.float32)
    shape = (1, 4, 6, 5, C)
    want = oracle.bev_pool_v2_forward(depth, feat, rd, rf, rb, shape, st, ln)
    t = lambda a: torch.from_numpy(a).cuda()
    d = t(depth).requires_grad_(True)
    f = t(feat).view(1, 1, 1, n_pix, C).requires_grad_(True)
    bev = rcb.bev_pool_v2(d, f, t(rd), t(rf), t(rb), shape, t(st), t(ln))
    _close(bev, oracle.to_bczyx(want), RTOL32, "general fwd")
    og = rng.standard_normal(bev.shape).astype(np.float32)
    bev.backward(t(og))
    # feat_grad accumulates over every point; depth_grad is a per-point write in the reference,
    # so it is only defined where ranks_depth is unique
    og_rows = np.ascontiguousarray(np.transpose(og, (0, 2, 3, 4, 1)))
    dg, fg = oracle.bev_pool_v2_backward(og_rows, depth, feat, rd, rf, rb)
    _close(f.grad.view(n_pix, C), fg, 1e-4, "general feat_grad")
    uniq, cnt = np.unique(rd, return_counts=True)
    once = uniq[cnt == 1]
    _close(d.grad[torch.from_numpy(once).cuda()], dg[once], RTOL32, "general depth_grad")


def test_out_of_range_rank_raises():
    import rcbevdet_b200 as rcb
    t = lambda a: torch.tensor(a, dtype=torch.int32, device="cuda")
    depth = torch.rand(1, 1, 2, 2, 2, device="cuda")
    feat = torch.rand(1, 1, 2, 2, 3, device="cuda")
    with pytest.raises(RuntimeError):
        rcb.bev_pool_v2(depth, feat, t([0, 99]), t([0, 1]), t([0, 1]), (1, 1, 2, 2, 3), t([0, 1]), t([1, 1]))


def test_cpu_tensors_raise():
    import rcbevdet_b200 as rcb
    t = lambda a: torch.tensor(a, dtype=torch.int32)
    with pytest.raises(RuntimeError):
        rcb.bev_pool_v2(torch.rand(1, 1, 2, 2, 2), torch.rand(1, 1, 2, 2, 3), t([0]), t([0]), t([0]),
                        (1, 1, 2, 2, 3), t([0]), t([1]))


def test_trt_wrapper_matches_bev_pool():
    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import rig
    coor, depth, feat = _case(B=1, C=16, depth_cfg=(1.0, 60.0, 4.0), input_size=(64, 176))
    ranks, shape, feat_rows, want = _oracle_pool(coor, depth, feat, rig.R50_GRID)
    rb, rd, rf, st, ln = (torch.from_numpy(r).cuda() for r in ranks)
    out = rcb.TRTBEVPoolv2.apply(depth[0].cuda(), feat[0].permute(0, 2, 3, 1).contiguous().cuda(), rd, rf, rb,
                                 st, ln, 128, 128)
    assert out.shape == (1, 128, 128, 16)
    _close(out, want[:, 0], RTOL32, "TRT wrapper")


def test_fused_voxel_pooling_v2_and_view_transform_flow():
    """Row V: the reference's call sequences (view_transformer.py:180-205 and the accelerate
    branch :277-289) on top of the new ops, plus the sync-free fused entry."""
    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import rig
    grid = rig.R50_GRID
    coor, depth, feat = _case(B=2, C=80)
    ranks, shape, feat_rows, want = _oracle_pool(coor, depth, feat, grid)
    want_bev = oracle.to_bczyx(want)[:, :, 0]            # collapse Z (Z == 1)
    lo, iv, sz = rig.grid_tensors(grid)
    # (1) fused
    d = depth.cuda().requires_grad_(True)
    f = feat.cuda().requires_grad_(True)
    bev = rcb.voxel_pooling_v2(coor.cuda(), d, f, lo, iv, sz)
    assert bev.shape == (2, 80, 128, 128)
    _close(bev, want_bev, RTOL32, "fused voxel_pooling_v2")
    og = torch.randn(bev.shape, generator=torch.Generator().manual_seed(5))
    bev.backward(og.cuda())
    og_rows = og.permute(0, 2, 3, 1).contiguous().view(shape).numpy()
    dg, fg = oracle.bev_pool_v2_backward(og_rows, depth.numpy(), feat_rows, ranks[1], ranks[2], ranks[0], threads=8)
    _close(d.grad, dg, RTOL32, "fused depth_grad")
    _close(f.grad.permute(0, 1, 3, 4, 2), fg, RTOL32, "fused feat_grad")
    # (2) accelerate=True flow: prepare once, .int().contiguous() the ranks, pool many times
    r = rcb.voxel_pooling_prepare_v2(coor.cuda(), lo, iv, sz)
    cached = [t.int().contiguous() for t in r]
    for _ in range(2):
        b2 = rcb.bev_pool_v2(depth.cuda(), feat.cuda().permute(0, 1, 3, 4, 2), cached[1], cached[2], cached[0],
                             shape, cached[3], cached[4]).squeeze(2)
        _close(b2, want_bev, RTOL32, "accelerate flow")
    # (3) nothing inside the grid -> zeros of the right shape (:184-194)
    far = torch.full_like(coor, 500.0).cuda()
    z = rcb.voxel_pooling_v2(far, depth.cuda(), feat.cuda(), lo, iv, sz)
    assert z.shape == (2, 80, 128, 128) and float(z.abs().max()) == 0.0


def test_hires_geometry_small_batch():
    """BASELINE config 5 geometry (56x100 features, 256x256 BEV), B=1: forward vs oracle."""
    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import rig
    grid = rig.HIRES_GRID
    coor = rig.lidar_coor(rig.camera_rig(1, input_size=rig.HIRES_INPUT), [1.0, 60.0, 2.0], rig.HIRES_INPUT, 16)
    _, N, D, H, W, _ = coor.shape
    depth, feat = rig.pooling_inputs(1, N, D, H, W, 80, seed=9)
    ranks, shape, feat_rows, want = _oracle_pool(coor, depth, feat, grid)
    lo, iv, sz = rig.grid_tensors(grid)
    bev = rcb.voxel_pooling_v2(coor.cuda(), depth.cuda(), feat.cuda(), lo, iv, sz, collapse_z=False)
    _close(bev, oracle.to_bczyx(want), RTOL32, "hires fwd")


def test_temporal_frames_folded_into_batch():
    """BASELINE config 3: the frames of a 4D sample are independent poolings (bevdet_rc.py:756-776),
    so T frames can be folded into the batch dimension of ONE launch.  The folded result must equal
    the per-frame results bit for bit, and match the oracle."""
    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import rig
    B, T = 2, 3
    grid = rig.R50_GRID
    motion = rig.temporal_motion(B, T, seed=3)                       # (B*T, 3) ego steps per frame
    calib = rig.camera_rig(B * T, input_size=(128, 352), frame_motion=motion)
    coor = rig.lidar_coor(calib, [1.0, 60.0, 2.0], (128, 352), 16)
    _, N, D, H, W, _ = coor.shape
    depth, feat = rig.pooling_inputs(B * T, N, D, H, W, 80, seed=8)
    lo, iv, sz = rig.grid_tensors(grid)
    folded = rcb.voxel_pooling_v2(coor.cuda(), depth.cuda(), feat.cuda(), lo, iv, sz)
    assert folded.shape == (B * T, 80, 128, 128)
    for f in range(B * T):
        one = rcb.voxel_pooling_v2(coor[f:f + 1].cuda(), depth[f:f + 1].cuda(), feat[f:f + 1].cuda(), lo, iv, sz)
        assert torch.equal(one[0], folded[f]), f"frame {f}"
    ranks, shape, feat_rows, want = _oracle_pool(coor, depth, feat, grid)
    _close(folded, oracle.to_bczyx(want)[:, :, 0], RTOL32, "folded frames vs oracle")
    # frames really differ (ego motion moved the frustum)
    assert not torch.equal(folded[0], folded[1])


def test_hires_full_size_properties():
    """BASELINE config 5 at full size (56x100 features, D=118, 256x256 BEV, B=2; ~8 M points):
    integer outputs against the C oracle, plus size-independent properties of the pooled tensor."""
    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import rig
    grid = rig.HIRES_GRID
    B = 2
    coor = rig.lidar_coor(rig.camera_rig(B, input_size=rig.HIRES_INPUT), grid["depth"], rig.HIRES_INPUT, 16)
    _, N, D, H, W, _ = coor.shape
    assert (D, H, W) == (118, 56, 100)
    lo, iv, sz = rig.grid_tensors(grid)
    got = rcb.voxel_pooling_prepare_v2(coor.cuda(), lo, iv, sz)
    want = oracle.voxel_pooling_prepare_v2_c(coor.numpy(), lo.numpy(), iv.numpy(), sz.numpy(), threads=8)
    assert_same_ranks(got, want)
    rb, rd, rf, st, ln = got
    assert int(ln.max()) > 512                      # exercises the CTA sort tier on real geometry
    assert bool((rb[1:] >= rb[:-1]).all())
    # linearity in depth and a checksum identity: sum over cells of out == sum over kept points of w * feat
    depth, feat = rig.pooling_inputs(B, N, D, H, W, 16, seed=2)
    d, f = depth.cuda(), feat.cuda()
    shape = (B, 1, 256, 256, 16)
    fv = f.permute(0, 1, 3, 4, 2)
    o1 = rcb.bev_pool_v2(d, fv, rd, rf, rb, shape, st, ln)
    o2 = rcb.bev_pool_v2(2.0 * d, fv, rd, rf, rb, shape, st, ln)
    assert torch.allclose(o2, 2.0 * o1, rtol=1e-6, atol=0)
    rows = fv.reshape(-1, 16).double()
    terms = d.flatten()[rd.long()].double()[:, None] * rows[rf.long()]
    total, scale = terms.sum(0), terms.abs().sum(0)
    # fp32 accumulation inside a cell: the error is relative to the sum of magnitudes, not to the
    # (heavily cancelling) signed total
    assert bool(((o1.double().sum(dim=(0, 2, 3, 4)) - total).abs() <= 1e-6 * scale).all())


@pytest.mark.parametrize("grid_xyz,C,dtype,layout", [
    ((50, 37, 2), 80, torch.float32, "bczyx"),     # patches clipped in x and in the folded (Z*Y) rows, Z = 2
    ((50, 37, 2), 80, torch.float32, "cells_c"),   # channels-last output: rows straight out, empty cells zero-filled
    ((13, 9, 1), 16, torch.float32, "bczyx"),      # fewer cells than one row of patches, 4 lanes carry a row
    ((64, 64, 1), 48, torch.float32, "bczyx"),     # lane count known only at run time
    ((64, 40, 1), 128, torch.float32, "cells_c"),  # all 32 lanes carry the row
    ((40, 64, 1), 256, torch.float32, "bczyx"),    # two quads per lane
    ((33, 33, 3), 136, torch.bfloat16, "bczyx"),   # two quads per lane at run time, bf16 rows
    ((128, 128, 1), 64, torch.float16, "cells_c"),
])
def test_forward_cells_kernel_variants(grid_xyz, C, dtype, layout):
    """k_fwd_cells over its template / layout space on ragged synthetic points: a few cells hold
    hundreds of points (several 32-point chunks, merged runs of one pixel), most hold a handful, many
    are empty; grids that are not multiples of the 8 x 4 patch."""
    import rcbevdet_b200 as rcb
    gx, gy, gz = grid_xyz
    B, N, D, H, W = 2, 2, 24, 6, 10
    rng = np.random.default_rng(gx * 1000 + C)
    coor = np.empty((B, N, D, H, W, 3), np.float32)
    coor[..., 0] = rng.random((B, N, D, H, W), dtype=np.float32) * (gx + 6) - 3
    coor[..., 1] = rng.random((B, N, D, H, W), dtype=np.float32) * (gy + 6) - 3
    coor[..., 2] = rng.random((B, N, D, H, W), dtype=np.float32) * (gz + 1) - 0.5
    # one camera's whole frustum into a 2 x 2 block of cells: long cells with every depth bin of a pixel
    # in one cell (runs of D equal rows to merge, in groups of four)
    coor[0, 0, ..., 0] = gx // 2 + rng.random((D, H, W), dtype=np.float32) * 2 - 1
    coor[0, 0, ..., 1] = gy // 2 + rng.random((D, H, W), dtype=np.float32) * 2 - 1
    coor[0, 0, ..., 0] = np.floor(coor[0, 0, :1, ..., 0]) + 0.5            # same (x, y) for all bins of a pixel
    coor[0, 0, ..., 1] = np.floor(coor[0, 0, :1, ..., 1]) + 0.5
    coor[0, 0, ..., 2] = 0.25
    lo, iv, sz = np.float32([0, 0, 0]), np.float32([1, 1, 1]), np.float32([gx, gy, gz])
    depth = rng.random((B, N, D, H, W), dtype=np.float32)
    feat = rng.standard_normal((B, N, C, H, W), dtype=np.float32)
    feat_t = torch.from_numpy(feat).to(dtype)
    rows = feat_t.float().permute(0, 1, 3, 4, 2).contiguous().numpy()      # what the kernel sees after rounding
    ranks = oracle.voxel_pooling_prepare_v2(coor, lo, iv, sz)
    shape = (B, gz, gy, gx, C)
    want = oracle.bev_pool_v2_forward(depth, rows, ranks[1], ranks[2], ranks[0], shape, ranks[3], ranks[4], threads=8)
    assert int(ranks[4].max()) > 100 and (want.reshape(-1, C) == 0).all(1).any()   # long cells and empty cells
    rb, rd, rf, st, ln = rcb.voxel_pooling_prepare_v2(torch.from_numpy(coor).cuda(), lo, iv, sz)
    fview = feat_t.cuda().permute(0, 1, 3, 4, 2)
    tol = RTOL32 if dtype == torch.float32 else 1e-4   # same rounded rows on both sides: only the summation differs
    if layout == "bczyx":
        got = rcb.bev_pool_v2(torch.from_numpy(depth).cuda(), fview, rd, rf, rb, shape, st, ln)
        _close(got, oracle.to_bczyx(want), tol, "bev (B,C,Z,Y,X)")
    else:
        got = rcb.QuickCumsumCuda.apply(torch.from_numpy(depth).cuda(), fview, rd, rf, rb, shape, st, ln)
        _close(got, want, tol, "bev (B,Z,Y,X,C)")
    empty = (want.reshape(-1, C) == 0).all(1)
    got_rows = (got.permute(0, 2, 3, 4, 1) if layout == "bczyx" else got).reshape(-1, C)
    assert float(got_rows[torch.from_numpy(empty).cuda()].abs().max()) == 0.0      # untouched cells are exactly zero


def test_channels_last_result_and_gradient_in_place():
    """bev_pool_v2(..., channels_last=True) / voxel_pooling_v2(..., channels_last=True): same shape and
    values as the reference layout, channels-last memory, and the backward agrees whether the incoming
    gradient is channels-last (used in place as rows) or (B, C, cells)-contiguous (transposed first)."""
    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import rig
    grid = rig.R50_GRID
    coor, depth, feat = _case(B=2, seed=31)
    lo, iv, sz = rig.grid_tensors(grid)
    rb, rd, rf, st, ln = rcb.voxel_pooling_prepare_v2(coor.cuda(), lo, iv, sz)
    shape = (2, 1, 128, 128, 80)
    og = torch.randn(2, 80, 1, 128, 128, device="cuda", generator=torch.Generator("cuda").manual_seed(4))
    results = {}
    for mode in ("ref", "cl", "cl_grad_contig"):
        d = depth.cuda().requires_grad_(True)
        f = feat.cuda().requires_grad_(True)
        bev = rcb.bev_pool_v2(d, f.permute(0, 1, 3, 4, 2), rd, rf, rb, shape, st, ln, channels_last=mode != "ref")
        assert bev.shape == (2, 80, 1, 128, 128)
        if mode == "ref":
            assert bev.is_contiguous()
            bev.backward(og)
        else:
            assert bev.permute(0, 2, 3, 4, 1).is_contiguous() and not bev.is_contiguous()
            g = og if mode == "cl_grad_contig" else og.permute(0, 2, 3, 4, 1).contiguous().permute(0, 4, 1, 2, 3)
            bev.backward(g)
        results[mode] = (bev.detach().contiguous(), d.grad, f.grad)
    for mode in ("cl", "cl_grad_contig"):
        for a, b, what in zip(results[mode], results["ref"], ("bev", "depth_grad", "feat_grad")):
            assert float((a - b).abs().max()) <= 1e-6 * float(b.abs().max()), f"{mode}: {what}"
    # the fused chain, Z == 1: (B, C, Y, X) channels-last, i.e. torch.channels_last
    d = depth.cuda().requires_grad_(True)
    f = feat.cuda().requires_grad_(True)
    bev = rcb.voxel_pooling_v2(coor.cuda(), d, f, lo, iv, sz, channels_last=True)
    assert bev.shape == (2, 80, 128, 128) and bev.is_contiguous(memory_format=torch.channels_last)
    ref = rcb.voxel_pooling_v2(coor.cuda(), depth.cuda(), feat.cuda(), lo, iv, sz)
    assert float((bev - ref).abs().max()) <= 1e-6 * float(ref.abs().max())
    conv = torch.nn.Conv2d(80, 8, 3, padding=1).cuda().to(memory_format=torch.channels_last)
    conv(bev).square().mean().backward()
    d2 = depth.cuda().requires_grad_(True)
    f2 = feat.cuda().requires_grad_(True)
    conv(rcb.voxel_pooling_v2(coor.cuda(), d2, f2, lo, iv, sz)).square().mean().backward()
    assert float((d.grad - d2.grad).abs().max()) <= 1e-4 * float(d2.grad.abs().max())
    assert float((f.grad - f2.grad).abs().max()) <= 1e-4 * float(f2.grad.abs().max())


def test_forward_bits_do_not_depend_on_the_batch_size():
    """The cell-stationary forward picks its CTA shape by the number of samples in the launch (1 / 2-4 /
    more); every shape pools a cell with one warp in point order, so a sample pooled alone, in a batch of
    3 and in a batch of 6 gives the same bits."""
    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import rig
    grid = rig.R50_GRID
    coor, depth, feat = _case(B=6, depth_cfg=(1.0, 60.0, 1.0), input_size=(128, 352), C=80, aug=3, seed=5)
    lo, iv, sz = rig.grid_tensors(grid)
    with torch.no_grad():
        full = rcb.voxel_pooling_v2(coor.cuda(), depth.cuda(), feat.cuda(), lo, iv, sz)
        three = rcb.voxel_pooling_v2(coor[:3].cuda(), depth[:3].cuda(), feat[:3].cuda(), lo, iv, sz)
        one = rcb.voxel_pooling_v2(coor[1:2].cuda(), depth[1:2].cuda(), feat[1:2].cuda(), lo, iv, sz)
    assert torch.equal(full[:3], three) and torch.equal(full[1:2], one)
    assert float(one.abs().max()) > 0
