"""temporal alignment (BEVDepth4D.gen_grid / shift_feature, bevdet_rc.py:585-657).
Golden vectors come from the reference's own two functions, executed from the reference file by
oracle/gen_golden.py.  CPU tests pin the torch restatement (the checker available on the GPU box)
and the product's host-side 3x3 transform; GPU tests compare the CUDA kernel with both."""
import numpy as np
import pytest
import torch

from oracle import refload, torch_ref

CASES = ["t16", "t32", "t128"]


def _case(g, name, device="cpu"):
    t = lambda k: torch.from_numpy(g[f"{name}.{k}"]).to(device)
    bda_adj = t("bda_adj") if f"{name}.bda_adj" in g.files else None
    return t("input"), [t("key"), t("adj")], t("bda"), bda_adj, g[f"{name}.interval"], g[f"{name}.lower"]


@pytest.mark.parametrize("name", CASES)
def test_restatement_matches_reference_golden(golden_temporal, name):
    x, s2k, bda, bda_adj, iv, lo = _case(golden_temporal, name)
    grid = torch_ref.gen_grid(x, s2k, bda, bda_adj, iv, lo)
    out = torch_ref.shift_feature(x, s2k, bda, bda_adj, iv, lo)
    assert np.array_equal(grid.numpy(), golden_temporal[f"{name}.grid"])
    assert np.array_equal(out.numpy(), golden_temporal[f"{name}.output"])


@pytest.mark.skipif(not refload.available(), reason="reference tree not present (GPU box)")
def test_restatement_matches_live_reference():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 4, 20, 28, generator=g)
    eye = torch.eye(4).view(1, 1, 4, 4).repeat(2, 6, 1, 1)
    adj = eye.clone()
    adj[:, :, 0, 3], adj[:, :, 1, 3] = 2.3, -1.1
    bda = torch.eye(3).repeat(2, 1, 1)
    iv, lo = [0.8 * 128 / 28, 0.8 * 128 / 20, 8.0], [-51.2, -51.2, -5.0]
    ref = refload.load_temporal_alignment(iv, lo)
    want = ref.shift_feature(x, [eye, adj], bda)
    assert torch.equal(torch_ref.shift_feature(x, [eye, adj], bda, None, iv, lo), want)


@pytest.mark.parametrize("name", CASES)
def test_host_transform_reproduces_the_reference_grid(golden_temporal, name):
    """rcbevdet_b200.temporal.gen_grid_transform (the 3x3 the kernel consumes) -> the reference's
    normalised grid, up to the rounding of one 3-term dot product."""
    from rcbevdet_b200.temporal import gen_grid_transform
    x, s2k, bda, bda_adj, iv, lo = _case(golden_temporal, name)
    n, _, h, w = x.shape
    tf = gen_grid_transform(n, s2k, bda, bda_adj, iv, lo).double()
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float64), torch.arange(w, dtype=torch.float64), indexing="ij")
    p = torch.stack((xs, ys, torch.ones_like(xs)), -1).view(1, h, w, 3, 1)
    q = tf.view(n, 1, 1, 3, 3).matmul(p)[..., :2, 0]
    grid = q / torch.tensor([w - 1.0, h - 1.0], dtype=torch.float64) * 2.0 - 1.0
    assert float((grid - torch.from_numpy(golden_temporal[f"{name}.grid"]).double()).abs().max()) < 2e-6


# ---------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_shift_feature_matches_reference_golden(golden_temporal, name):
    """The CUDA kernel against the reference's outputs.  Sampling positions differ from the
    reference's by the rounding order of one 3-term dot product (<= 1 ulp of a pixel coordinate),
    i.e. bilinear weights by ~1e-5: tolerance 1e-4 of max |feature| (documented)."""
    import rcbevdet_b200 as rcb
    x, s2k, bda, bda_adj, iv, lo = _case(golden_temporal, name, "cuda")
    got = rcb.shift_feature(x, s2k, bda, bda_adj, iv, lo)
    want = torch.from_numpy(golden_temporal[f"{name}.output"]).cuda()
    assert got.shape == want.shape and got.dtype == torch.float32
    assert float((got - want).abs().max()) <= 1e-4 * float(want.abs().max())


@pytest.mark.gpu
def test_gpu_shift_feature_full_size_forward_backward():
    """BASELINE config 3 size: 8 samples x C=80 x 128x128, against torch's grid_sample fed with the
    restated gen_grid, forward and the gradient w.r.t. the feature map."""
    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import rig
    n, C = 8, 80
    motion = rig.temporal_motion(n, 2, seed=9).view(n, 2, 3)
    key = rig.camera_rig(n, frame_motion=motion[:, 0])[0].cuda()
    adj = rig.camera_rig(n, frame_motion=motion[:, 1] * 2.0)[0].cuda()
    bda = torch.eye(3).repeat(n, 1, 1).cuda()
    lo, iv, _ = rig.grid_tensors(rig.R50_GRID)
    x = torch.randn(n, C, 128, 128, device="cuda", generator=torch.Generator("cuda").manual_seed(2))
    xa = x.clone().requires_grad_(True)
    xb = x.clone().requires_grad_(True)
    got = rcb.shift_feature(xa, [key, adj], bda, None, iv, lo)
    want = torch_ref.shift_feature(xb, [key, adj], bda, None, iv, lo)
    scale = float(want.abs().max())
    assert float((got - want).abs().max()) <= 1e-4 * scale
    og = torch.randn_like(want)
    got.backward(og)
    want.backward(og)
    assert float((xa.grad - xb.grad).abs().max()) <= 1e-4 * float(xb.grad.abs().max())
    # identity motion: the feature map comes back unchanged where it is sampled inside the grid
    same = rcb.shift_feature(x, [key, key], bda, None, iv, lo)
    assert float((same - x).abs().max()) <= 1e-4 * float(x.abs().max())


@pytest.mark.gpu
def test_gpu_shift_feature_argument_errors():
    import rcbevdet_b200 as rcb
    with pytest.raises(RuntimeError):
        rcb.shift_feature(torch.zeros(1, 2, 8, 8), [torch.eye(4).view(1, 1, 4, 4)] * 2, torch.eye(3).view(1, 3, 3),
                          None, [0.8, 0.8, 8.0], [-51.2, -51.2, -5.0])
