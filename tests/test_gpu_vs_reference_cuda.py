"""This repo's CUDA kernels vs the reference's OWN CUDA kernels (mmdet3d/ops/bev_pool_v2/src/*,
compiled unmodified into oracle/_ref by oracle/build_ref.py) on identical inputs, and the CPU
oracle vs the same reference kernels (which pins the oracle to the reference itself)."""
import numpy as np
import pytest
import torch

from oracle import oracle, ref_cuda

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_cuda.available(), reason="oracle/_ref/bev_pool_v2_ext.so not built")]


def _rel(a, b):
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30)


def test_reference_kat_on_reference_kernels():
    """The reference's known-answer test (bev_pool.py:145-176) through its own extension."""
    depth = torch.tensor([0.3, 0.4, 0.2, 0.1, 0.7, 0.6, 0.8, 0.9], device="cuda").view(1, 1, 2, 2, 2)
    feat = torch.ones(1, 1, 2, 2, 2, device="cuda")
    t = lambda a: torch.tensor(a, dtype=torch.int32, device="cuda")
    rd, rf, rb, st, ln = t([0, 4, 1, 6]), t([0, 0, 1, 2]), t([0, 0, 1, 1]), t([0, 2]), t([2, 2])
    bev = ref_cuda.bev_pool_v2(depth, feat, rd, rf, rb, (1, 1, 2, 2, 2), st, ln)
    assert bev.sum().item() == pytest.approx(4.4, abs=1e-6)
    dg, fg = ref_cuda.backward(torch.ones(1, 1, 2, 2, 2, device="cuda"), depth, feat, rd, rf, rb)
    assert torch.allclose(dg.flatten().cpu(), torch.tensor([2., 2., 0., 0., 2., 0., 2., 0.]))
    assert torch.allclose(fg.flatten().cpu(), torch.tensor([1.0, 1.0, 0.4, 0.4, 0.8, 0.8, 0., 0.]))


@pytest.mark.parametrize("B,aug", [(2, None), (1, 9)])
def test_full_r50_geometry_against_reference_kernels(B, aug):
    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import rig
    grid = rig.R50_GRID
    coor = rig.lidar_coor(rig.camera_rig(B, aug_seed=aug), grid["depth"], rig.R50_INPUT, 16).cuda()
    _, N, D, H, W, _ = coor.shape
    depth, feat = rig.pooling_inputs(B, N, D, H, W, 80, seed=21)
    depth, feat = depth.cuda(), feat.cuda()
    lo, iv, sz = rig.grid_tensors(grid)
    rb, rd, rf, st, ln = rcb.voxel_pooling_prepare_v2(coor, lo, iv, sz)
    shape = (B, 1, 128, 128, 80)
    fview = feat.permute(0, 1, 3, 4, 2)
    # forward
    want = ref_cuda.bev_pool_v2(depth, fview, rd, rf, rb, shape, st, ln)
    d = depth.clone().requires_grad_(True)
    f = feat.clone().requires_grad_(True)
    got = rcb.bev_pool_v2(d, f.permute(0, 1, 3, 4, 2), rd, rf, rb, shape, st, ln)
    assert _rel(got.detach(), want) <= 1e-5
    # backward
    og = torch.randn(got.shape, device="cuda", generator=torch.Generator("cuda").manual_seed(3))
    got.backward(og)
    dg, fg = ref_cuda.backward(og.permute(0, 2, 3, 4, 1).contiguous(), depth, fview.contiguous(), rd, rf, rb)
    assert _rel(d.grad, dg) <= 1e-5
    assert _rel(f.grad.permute(0, 1, 3, 4, 2), fg) <= 1e-5
    # and the CPU oracle agrees with the reference kernels too (pins the oracle to the reference)
    o = oracle.bev_pool_v2_forward(depth.cpu().numpy(), fview.contiguous().cpu().numpy(), rd.cpu().numpy(),
                                   rf.cpu().numpy(), rb.cpu().numpy(), shape, st.cpu().numpy(), ln.cpu().numpy(),
                                   threads=8)
    assert _rel(torch.from_numpy(oracle.to_bczyx(o)), want.cpu()) <= 1e-6


def test_timing_report_against_reference_kernels(capsys):
    """Not an assertion on speed: records, for profiles/, how long the reference's recompiled
    kernels take on the BASELINE config-2 inputs next to this repo's (run with -s to see it)."""
    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import rig
    B = 8
    grid = rig.R50_GRID
    coor = rig.lidar_coor(rig.camera_rig(B), grid["depth"], rig.R50_INPUT, 16).cuda()
    _, N, D, H, W, _ = coor.shape
    depth, feat = rig.pooling_inputs(B, N, D, H, W, 80, seed=1)
    depth, feat = depth.cuda(), feat.cuda()
    lo, iv, sz = rig.grid_tensors(grid)
    rb, rd, rf, st, ln = rcb.voxel_pooling_prepare_v2(coor, lo, iv, sz)
    shape = (B, 1, 128, 128, 80)
    fview = feat.permute(0, 1, 3, 4, 2)
    og = torch.randn(B, 80, 1, 128, 128, device="cuda")

    def timed(fn, n=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    def ours():
        d = depth.detach().requires_grad_(True)
        f = feat.detach().requires_grad_(True)
        rcb.bev_pool_v2(d, f.permute(0, 1, 3, 4, 2), rd, rf, rb, shape, st, ln).backward(og)

    def ref():
        ref_cuda.bev_pool_v2(depth, fview, rd, rf, rb, shape, st, ln)
        ref_cuda.backward(og.permute(0, 2, 3, 4, 1).contiguous(), depth, fview.contiguous(), rd, rf, rb)

    t_ours, t_ref = timed(ours), timed(ref)
    with capsys.disabled():
        print(f"\n[timing] bev_pool_v2 fwd+bwd, B=8 R50: this repo {t_ours:.3f} ms, reference kernels "
              f"(recompiled for sm_100a, with their host-side prep) {t_ref:.3f} ms, ratio {t_ref / t_ours:.1f}x")
    assert t_ours > 0 and t_ref > 0
