"""the fused producer (channel split + depth softmax + channels-last context) and
the forward()-tail chain built on it, against the reference's formulation in torch
(view_transformer.py:316-320, bev_pool.py:21) -- forward and backward."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_split(x, D, C):
    depth = x[:, :D].float().softmax(dim=1)
    ctx = x[:, D:D + C].float().permute(0, 2, 3, 1).contiguous()
    return depth, ctx


@pytest.mark.parametrize("n_img,D,C,H,W,extra", [(12, 118, 80, 16, 44, 0), (3, 59, 64, 7, 13, 0), (2, 30, 12, 5, 40, 5),
                                                 (1, 200, 128, 3, 33, 0)])
def test_split_forward_backward(n_img, D, C, H, W, extra):
    import rcbevdet_b200 as rcb
    g = torch.Generator("cuda").manual_seed(D + C)
    x = (torch.randn(n_img, D + C + extra, H, W, device="cuda", generator=g) * 3.0)
    xa = x.clone().requires_grad_(True)
    xb = x.clone().requires_grad_(True)
    depth, ctx = rcb.depth_context_split(xa, D, C)
    want_d, want_c = _ref_split(xb, D, C)
    assert depth.shape == (n_img, D, H, W) and ctx.shape == (n_img, H, W, C) and ctx.is_contiguous()
    assert torch.equal(ctx, want_c)                                   # a copy: bit-exact
    assert float((depth.detach() - want_d.detach()).abs().max()) <= 1e-6   # same expf, other summation order (values <= 1)
    assert float((depth.sum(1) - 1).abs().max()) <= 1e-5
    gd = torch.randn(depth.shape, device="cuda", generator=g)
    gc = torch.randn(ctx.shape, device="cuda", generator=g)
    torch.autograd.backward([depth, ctx], [gd, gc])
    torch.autograd.backward([want_d, want_c], [gd, gc])
    assert float((xa.grad - xb.grad).abs().max()) <= 1e-6 * max(1.0, float(xb.grad.abs().max()))
    if extra:
        assert float(xa.grad[:, D + C:].abs().max()) == 0.0


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_split_half_inputs(dtype):
    import rcbevdet_b200 as rcb
    x = torch.randn(4, 118 + 80, 16, 44, device="cuda", generator=torch.Generator("cuda").manual_seed(1)).to(dtype)
    depth, ctx = rcb.depth_context_split(x, 118, 80)
    want_d, want_c = _ref_split(x, 118, 80)
    assert depth.dtype == torch.float32 and torch.equal(ctx, want_c)
    assert float((depth - want_d).abs().max()) <= 1e-6


def test_forward_tail_chain_matches_the_unfused_sequence():
    """lss_view_transform(x, ...) == softmax / slice in torch -> voxel_pooling_v2_from_calib, values and
    gradients w.r.t. the depth-net output."""
    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import rig
    B, N, D, C = 2, 6, 118, 80
    calib = rig.camera_rig(B, aug_seed=4)
    axes = rcb.frustum_axes(rig.R50_GRID["depth"], rig.R50_INPUT, 16)
    lo, iv, sz = rig.grid_tensors(rig.R50_GRID)
    x = torch.randn(B * N, D + C, 16, 44, device="cuda", generator=torch.Generator("cuda").manual_seed(7))
    xa = x.clone().requires_grad_(True)
    xb = x.clone().requires_grad_(True)
    bev, depth = rcb.lss_view_transform(xa, N, D, C, calib, axes, lo, iv, sz)
    want_depth = xb[:, :D].softmax(dim=1)
    want = rcb.voxel_pooling_v2_from_calib(calib, axes, want_depth.view(B, N, D, 16, 44),
                                           xb[:, D:D + C].view(B, N, C, 16, 44), lo, iv, sz)
    assert bev.shape == want.shape == (B, C, 128, 128)
    assert float((bev - want).abs().max()) <= 1e-5 * float(want.abs().max())
    assert float((depth.detach() - want_depth.detach()).abs().max()) <= 1e-6
    og = torch.randn(bev.shape, device="cuda", generator=torch.Generator("cuda").manual_seed(8))
    gd = torch.randn(depth.shape, device="cuda", generator=torch.Generator("cuda").manual_seed(9)) * 0.1
    torch.autograd.backward([bev, depth], [og, gd])          # depth also feeds the depth loss in BEVDepth
    torch.autograd.backward([want, want_depth], [og, gd])
    assert float((xa.grad - xb.grad).abs().max()) <= 1e-5 * float(xb.grad.abs().max())


def test_split_argument_errors():
    import rcbevdet_b200 as rcb
    with pytest.raises(RuntimeError):
        rcb.depth_context_split(torch.zeros(1, 10, 2, 2), 6, 4)
    with pytest.raises(ValueError):
        rcb.depth_context_split(torch.zeros(1, 8, 2, 2, device="cuda"), 6, 4)
    with pytest.raises(RuntimeError):
        rcb.depth_context_split(torch.zeros(1, 400, 2, 2, device="cuda"), 300, 4)   # D > 256
