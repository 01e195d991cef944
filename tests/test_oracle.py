"""The oracle itself, pinned against everything the reference offers for this path:
its one known-answer test and golden vectors produced by running the reference's own
Python code (oracle/gen_golden.py).  CPU only."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import oracle, refload

PREP_CASES = ["rigA", "randB", "augD", "onecellE", "singleF"]


def test_kat_forward_backward():
    """mmdet3d/ops/bev_pool_v2/bev_pool.py:145-176: loss 4.4, depth.grad, feat.grad."""
    depth = np.array([0.3, 0.4, 0.2, 0.1, 0.7, 0.6, 0.8, 0.9], np.float32).reshape(1, 1, 2, 2, 2)
    feat = np.ones((1, 1, 2, 2, 2), np.float32)
    rd = np.array([0, 4, 1, 6], np.int32)
    rf = np.array([0, 0, 1, 2], np.int32)
    rb = np.array([0, 0, 1, 1], np.int32)
    st, ln = oracle.intervals_from_sorted(rb)
    assert st.tolist() == [0, 2] and ln.tolist() == [2, 2]
    out = oracle.bev_pool_v2_forward(depth, feat, rd, rf, rb, (1, 1, 2, 2, 2), st, ln)
    bev = oracle.to_bczyx(out)
    assert bev.shape == (1, 2, 1, 2, 2)
    assert np.float32(bev.sum()) == np.float32(4.4)
    dg, fg = oracle.bev_pool_v2_backward(np.ones_like(out), depth, feat, rd, rf, rb)
    np.testing.assert_allclose(dg.ravel(), [2., 2., 0., 0., 2., 0., 2., 0.], rtol=0, atol=0)
    np.testing.assert_allclose(fg.ravel(), [1.0, 1.0, 0.4, 0.4, 0.8, 0.8, 0., 0.], rtol=1e-7)


@pytest.mark.parametrize("name", PREP_CASES)
def test_prepare_matches_reference_golden(golden_prepare, name):
    g = golden_prepare
    out = oracle.voxel_pooling_prepare_v2(g[f"{name}.coor"], g[f"{name}.lower"], g[f"{name}.interval"],
                                          g[f"{name}.size"])
    assert int(g[f"{name}.empty"]) == 0
    rb, rd, rf, st, ln = out
    for a in out:
        assert a.dtype == np.int32
    # tie-independent outputs: bit-exact against the reference as-is
    assert np.array_equal(rb, g[f"{name}.raw.ranks_bev"])
    assert np.array_equal(st, g[f"{name}.raw.interval_starts"])
    assert np.array_equal(ln, g[f"{name}.raw.interval_lengths"])
    # tie-dependent outputs: bit-exact against the canonicalised reference
    assert np.array_equal(rd, g[f"{name}.canon.ranks_depth"])
    assert np.array_equal(rf, g[f"{name}.canon.ranks_feat"])
    # the oracle's own output is already canonical
    crb, crd, crf = oracle.canonicalise(rb, rd, rf)
    assert np.array_equal(crd, rd) and np.array_equal(crf, rf) and np.array_equal(crb, rb)


@pytest.mark.parametrize("name", PREP_CASES + ["emptyC"])
@pytest.mark.parametrize("threads", [1, 4])
def test_c_prepare_equals_numpy_prepare(golden_prepare, name, threads):
    """The compiled restatement used as the timed CPU baseline is the same function."""
    g = golden_prepare
    args = (g[f"{name}.coor"], g[f"{name}.lower"], g[f"{name}.interval"], g[f"{name}.size"])
    a, b = oracle.voxel_pooling_prepare_v2(*args), oracle.voxel_pooling_prepare_v2_c(*args, threads=threads)
    for x, y in zip(a, b):
        assert (x is None and y is None) or np.array_equal(x, y)


def test_prepare_empty(golden_prepare):
    g = golden_prepare
    assert int(g["emptyC.empty"]) == 1
    out = oracle.voxel_pooling_prepare_v2(g["emptyC.coor"], g["emptyC.lower"], g["emptyC.interval"],
                                          g["emptyC.size"])
    assert out == (None, None, None, None, None)


def test_truncation_keeps_minus_one_to_zero(golden_prepare):
    """.long() truncates toward zero (view_transformer.py:232): voxel coords in (-1, 0)
    land in cell 0 and are KEPT.  Points 2 and 3 of randB were planted there."""
    g = golden_prepare
    rd = g["randB.canon.ranks_depth"]
    assert 2 in rd and 3 in rd and 0 in rd
    assert 1 not in rd and 4 not in rd


def test_grid_infos_match_reference(golden_prepare):
    g = golden_prepare
    lo, it, sz = oracle.grid_infos([-51.2, 51.2, 0.8], [-51.2, 51.2, 0.8], [-5, 3, 8])
    assert np.array_equal(lo, g["rigA.lower"]) and np.array_equal(it, g["rigA.interval"])
    assert np.array_equal(sz, g["rigA.size"])
    assert np.array_equal(oracle.create_frustum([1.0, 60.0, 5.0], (64, 176), 16), g["rigA.frustum"])


@pytest.mark.parametrize("name", ["r32", "r64", "r128"])
def test_radar_matches_reference_golden(golden_radar, name):
    g = golden_radar
    B, ny, nx, cin = (int(v) for v in g[f"{name}.shape"])
    f, h, hf = oracle.radar_rcs_scatter(g[f"{name}.point_features"], g[f"{name}.rcs"], g[f"{name}.coors"],
                                        B, ny, nx)
    assert np.array_equal(f, g[f"{name}.features"])
    assert np.array_equal(h, g[f"{name}.heatmap"])
    assert np.array_equal(hf, g[f"{name}.heatmap_feat"])


@pytest.mark.skipif(not refload.available(), reason="reference tree not present (GPU box)")
def test_full_size_digest_against_live_reference():
    """Config 1 (B=1, full R50 geometry): oracle vs the digests of the reference's output."""
    import torch
    from rcbevdet_b200 import rig
    with open(os.path.join(os.path.dirname(__file__), "golden", "prepare_full_digest.json")) as f:
        dg = json.load(f)
    vt = refload.load_view_transformer()
    m = vt.LSSViewTransformer(grid_config=dict(rig.R50_GRID), input_size=rig.R50_INPUT, downsample=16,
                              in_channels=8, out_channels=8)
    coor = m.get_lidar_coor(*rig.camera_rig(1)).numpy().astype(np.float32)
    if hashlib.sha256(coor.tobytes()).hexdigest() != dg["sha256"]["coor"]:
        pytest.skip("host BLAS rounds get_lidar_coor differently from the fixture machine")
    out = oracle.voxel_pooling_prepare_v2(coor, m.grid_lower_bound.numpy(), m.grid_interval.numpy(),
                                          m.grid_size.numpy())
    assert out[0].shape[0] == dg["K"] and out[3].shape[0] == dg["I"]
    for k, a in zip(("ranks_bev", "ranks_depth", "ranks_feat", "interval_starts", "interval_lengths"), out):
        assert hashlib.sha256(a.tobytes()).hexdigest() == dg["sha256"][k], k


@pytest.mark.parametrize("name", PREP_CASES + ["emptyC"])
def test_torch_restatement_matches_reference_golden(golden_prepare, name):
    """oracle/torch_ref.py (the formulation bench.py times as "the reference's PyTorch path") gives
    the reference's own outputs: tie-independent arrays as they come, the others canonicalised."""
    import torch
    from oracle import torch_ref
    g = golden_prepare
    out = torch_ref.prepare(torch.from_numpy(g[f"{name}.coor"]), g[f"{name}.lower"], g[f"{name}.interval"],
                            g[f"{name}.size"])
    if int(g[f"{name}.empty"]):
        assert out == (None,) * 5
        return
    rb, rd, rf, st, ln = (t.numpy() for t in out)
    assert np.array_equal(rb, g[f"{name}.raw.ranks_bev"])
    assert np.array_equal(st, g[f"{name}.raw.interval_starts"])
    assert np.array_equal(ln, g[f"{name}.raw.interval_lengths"])
    crb, crd, crf = oracle.canonicalise(rb, rd, rf)
    assert np.array_equal(crd, g[f"{name}.canon.ranks_depth"]) and np.array_equal(crf, g[f"{name}.canon.ranks_feat"])


def test_torch_pool_restatement_kat_and_oracle(golden_prepare):
    """index_add_ restatement: the reference's known-answer test (bev_pool.py:145-176) through
    autograd, and agreement with the C oracle on a golden geometry."""
    import torch
    from oracle import torch_ref
    depth = torch.tensor([0.3, 0.4, 0.2, 0.1, 0.7, 0.6, 0.8, 0.9]).view(1, 1, 2, 2, 2).requires_grad_(True)
    feat = torch.ones(1, 1, 2, 2, 2, requires_grad=True)
    t = lambda a: torch.tensor(a, dtype=torch.int32)
    bev = torch_ref.pool(depth, feat, t([0, 4, 1, 6]), t([0, 0, 1, 2]), t([0, 0, 1, 1]), (1, 1, 2, 2, 2))
    loss = bev.sum()
    loss.backward()
    assert abs(float(loss) - 4.4) < 1e-6
    assert torch.allclose(depth.grad.flatten(), torch.tensor([2., 2., 0., 0., 2., 0., 2., 0.]))
    assert torch.allclose(feat.grad.flatten(), torch.tensor([1.0, 1.0, 0.4, 0.4, 0.8, 0.8, 0., 0.]))
    g = golden_prepare
    coor = g["augD.coor"]
    B, N, D, H, W, _ = coor.shape
    rb, rd, rf, st, ln = oracle.voxel_pooling_prepare_v2(coor, g["augD.lower"], g["augD.interval"], g["augD.size"])
    gx, gy, gz = (int(v) for v in g["augD.size"])
    rng = np.random.default_rng(5)
    dep = rng.random((B, N, D, H, W), dtype=np.float32)
    rows = rng.standard_normal((B, N, H, W, 16), dtype=np.float32)
    want = oracle.to_bczyx(oracle.bev_pool_v2_forward(dep, rows, rd, rf, rb, (B, gz, gy, gx, 16), st, ln))
    got = torch_ref.pool(torch.from_numpy(dep), torch.from_numpy(rows), torch.from_numpy(rd), torch.from_numpy(rf),
                         torch.from_numpy(rb), (B, gz, gy, gx, 16)).numpy()
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()
