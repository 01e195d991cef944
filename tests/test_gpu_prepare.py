"""CUDA voxel_pooling_prepare_v2 (through the C ABI) vs the CPU oracle and the golden vectors
produced by the reference's own code.  Integer outputs: bit-exact."""
import numpy as np
import pytest
import torch

from oracle import oracle
from rank_check import assert_same_ranks

pytestmark = pytest.mark.gpu

NAMES = ("ranks_bev", "ranks_depth", "ranks_feat", "interval_starts", "interval_lengths")


def _gpu_prepare(coor, lower, interval, size):
    import rcbevdet_b200 as rcb
    out = rcb.voxel_pooling_prepare_v2(torch.from_numpy(np.ascontiguousarray(coor)).cuda(),
                                       torch.from_numpy(lower), torch.from_numpy(interval),
                                       torch.from_numpy(size))
    return out


def _check_against_oracle(coor, lower, interval, size):
    got = _gpu_prepare(coor, lower, interval, size)
    want = oracle.voxel_pooling_prepare_v2(coor, lower, interval, size)
    if want[0] is None:
        assert all(g is None for g in got)
        return got
    for name, g in zip(NAMES, got):
        assert g.dtype == torch.int32 and g.is_contiguous() and g.is_cuda, name
    assert_same_ranks(got, want)
    return got


@pytest.mark.parametrize("name", ["rigA", "randB", "augD", "onecellE", "singleF"])
def test_golden_cases(golden_prepare, name):
    g = golden_prepare
    got = _check_against_oracle(g[f"{name}.coor"], g[f"{name}.lower"], g[f"{name}.interval"], g[f"{name}.size"])
    rb, rd, rf, st, ln = (t.cpu().numpy() for t in got)
    # tie-independent outputs: bit-exact against the reference as-is
    assert np.array_equal(rb, g[f"{name}.raw.ranks_bev"])
    assert np.array_equal(st, g[f"{name}.raw.interval_starts"])
    assert np.array_equal(ln, g[f"{name}.raw.interval_lengths"])
    # tie-dependent outputs: bit-exact against the canonicalised reference
    _, rd_c, rf_c = oracle.canonicalise(rb, rd, rf)
    assert np.array_equal(rd_c, g[f"{name}.canon.ranks_depth"])
    assert np.array_equal(rf_c, g[f"{name}.canon.ranks_feat"])


def test_empty_returns_five_nones(golden_prepare):
    g = golden_prepare
    got = _gpu_prepare(g["emptyC.coor"], g["emptyC.lower"], g["emptyC.interval"], g["emptyC.size"])
    assert got == (None, None, None, None, None)


def test_full_r50_batch2_matches_oracle_and_counts():
    """BASELINE config geometry (6 cams, D=118, 16x44, 128x128), B=2, deterministic rig."""
    from rcbevdet_b200 import rig
    coor = rig.lidar_coor(rig.camera_rig(2), rig.R50_GRID["depth"], rig.R50_INPUT, 16).numpy()
    lo, iv, sz = (t.numpy() for t in rig.grid_tensors(rig.R50_GRID))
    got = _check_against_oracle(coor, lo, iv, sz)
    assert got[0].numel() == 2 * 359501 and got[3].numel() == 2 * 12638


def test_train_aug_rig_matches_oracle():
    from rcbevdet_b200 import rig
    coor = rig.lidar_coor(rig.camera_rig(3, aug_seed=5), rig.R50_GRID["depth"], rig.R50_INPUT, 16).numpy()
    lo, iv, sz = (t.numpy() for t in rig.grid_tensors(rig.R50_GRID))
    _check_against_oracle(coor, lo, iv, sz)


@pytest.mark.parametrize("shape", [(1, 1, 5, 3, 7), (3, 1, 1, 1, 3), (2, 3, 7, 5, 9), (1, 2, 33, 5, 5)])
def test_ragged_shapes(shape):
    """P not a multiple of 4, samples smaller than a quad, Z > 1, non-square grids."""
    rng = np.random.default_rng(sum(shape))
    coor = (rng.random(shape + (3,), dtype=np.float32) * np.float32([24, 20, 14]) - np.float32([12, 10, 8]))
    lo = np.float32([-8.0, -6.0, -5.0])
    iv = np.float32([1.0, 0.5, 4.0])
    sz = np.float32([16.0, 24.0, 2.0])
    _check_against_oracle(coor, lo, iv, sz)


@pytest.mark.parametrize("n_cells_hit,P", [(1, 32), (1, 33), (1, 40), (1, 64), (1, 65), (1, 100), (1, 128), (1, 129), (1, 200), (1, 256), (1, 257), (1, 500), (1, 512), (1, 513), (1, 1024), (1, 1025),
                                          (2, 300), (7, 900), (3, 3000), (2, 9000), (5, 20001)])
def test_long_cells_every_sort_tier(n_cells_hit, P):
    """Cells with <= 256 points (register bitonic, 1/2/4/8 values per lane), 257..4096 points
    (shared-memory bitonic) and > 4096 points (in-place global bitonic), incl. non-power-of-two lengths."""
    rng = np.random.default_rng(P)
    coor = np.empty((1, 1, P, 1, 1, 3), np.float32)
    which = rng.integers(0, n_cells_hit, size=P)
    coor[0, 0, :, 0, 0, 0] = which * 1.0 + 0.25 + rng.random(P, dtype=np.float32) * 0.5
    coor[0, 0, :, 0, 0, 1] = 0.1
    coor[0, 0, :, 0, 0, 2] = 0.0
    lo = np.float32([0.0, 0.0, -1.0])
    iv = np.float32([1.0, 1.0, 2.0])
    sz = np.float32([8.0, 4.0, 1.0])
    got = _check_against_oracle(coor, lo, iv, sz)
    assert got[3].numel() == len(np.unique(which))


@pytest.mark.parametrize("grid,B,P,hot", [
    ((64, 32, 1), 1, 30000, 0.0),       # 2^11 cells: buckets of two cells
    ((64, 64, 1), 1, 50000, 0.5),       # exactly 2^12 cells, half the points in one BEV row -> multi-chunk buckets
    ((100, 50, 3), 2, 120000, 0.3),     # 30000 cells, not a power of two, Z > 1
    ((128, 128, 1), 3, 200000, 0.7),    # R50 grid, 140k points in one row of one sample
    ((1024, 1024, 1), 1, 150000, 0.2),  # 2^20 cells: 1024-cell buckets
    ((1500, 1000, 1), 1, 100000, 0.2),  # 1.5 M cells: 2048-cell buckets (two-level up to 2^22 cells)
    ((2048, 2048, 1), 1, 120000, 0.3),  # exactly 2^22 cells: 4096-cell buckets, the last two-level grid
    ((2500, 2000, 1), 1, 100000, 0.2),  # 5 M cells > 2^22: three plain LSD passes (prepare_lsd.cu)
    ((700, 700, 1), 9, 90000, 0.0),     # 4.4 M cells in 9 samples: LSD passes with a batch offset
])
def test_two_level_sort_grids(grid, B, P, hot):
    """Every sort path of the prepare stage (global pass + bucket sort with one and several chunks per
    bucket, every bucket width up to 4096 cells, three LSD passes beyond 2^22 cells), bit-exact
    against the oracle."""
    rng = np.random.default_rng(P + grid[0])
    gx, gy, gz = grid
    per = P // B
    coor = np.empty((B, 1, per, 1, 1, 3), np.float32)
    coor[..., 0] = rng.random((B, 1, per, 1, 1), dtype=np.float32) * (gx + 4) - 2
    coor[..., 1] = rng.random((B, 1, per, 1, 1), dtype=np.float32) * (gy + 4) - 2
    coor[..., 2] = rng.random((B, 1, per, 1, 1), dtype=np.float32) * (gz + 1) - 0.5
    n_hot = int(per * hot)
    if n_hot:  # concentrate points of the last sample in one BEV row (and a few cells of it)
        coor[B - 1, 0, :n_hot, 0, 0, 1] = gy // 2 + 0.5
        coor[B - 1, 0, : n_hot // 2, 0, 0, 0] = rng.integers(0, 3, size=n_hot // 2) + 0.5
    lo = np.float32([0.0, 0.0, 0.0])
    iv = np.float32([1.0, 1.0, 1.0])
    sz = np.float32([gx, gy, gz])
    _check_against_oracle(coor, lo, iv, sz)


def test_truncation_toward_zero_and_boundaries():
    """view_transformer.py:232 `.long()`: voxel coordinates in (-1, 0) are KEPT in cell 0; exactly
    -1.0 and exactly `size` are dropped."""
    lo = np.float32([0.0, 0.0, 0.0])
    iv = np.float32([1.0, 1.0, 1.0])
    sz = np.float32([4.0, 4.0, 2.0])
    pts = np.float32([[-0.5, 0.5, 0.5], [-1.0, 0.5, 0.5], [3.999, 3.999, 1.999], [4.0, 0.0, 0.0],
                      [0.0, -0.999, -0.999], [0.0, 0.0, 2.0], [1.5, 2.5, 1.5], [1.5, 2.5, 1.5]])
    coor = pts.reshape(1, 1, 8, 1, 1, 3)
    got = _check_against_oracle(coor, lo, iv, sz)
    assert got[1].cpu().tolist() == [0, 4, 6, 7, 2]   # cells 0, 0, 25, 25, 31


def test_nan_inf_are_dropped():
    """`.long()` of NaN is INT64_MIN on x86 and on CUDA (cvt.rzi.s64.f32 / __float2ll_rz), +-Inf
    saturates: all three fail the range test (view_transformer.py:238-240) on every platform."""
    lo = np.float32([0.0, 0.0, 0.0])
    iv = np.float32([1.0, 1.0, 1.0])
    sz = np.float32([4.0, 4.0, 1.0])
    pts = np.float32([[np.nan, 0.5, 0.5], [np.inf, 0.5, 0.5], [-np.inf, 0.5, 0.5], [1.5, 0.5, 0.5],
                      [0.5, np.nan, 0.5], [0.5, 0.5, np.nan], [0.5, 0.5, 0.5], [0.5, 0.5, 0.5]])
    rb, rd, rf, st, ln = _check_against_oracle(pts.reshape(1, 1, 8, 1, 1, 3), lo, iv, sz)
    assert rd.cpu().tolist() == [6, 7, 3] and rb.cpu().tolist() == [0, 0, 1]


def test_unsupported_geometry_raises():
    import rcbevdet_b200 as rcb
    coor = torch.zeros(1, 1, 1, 1, 4, 3, device="cuda")
    with pytest.raises(RuntimeError):
        rcb.voxel_pooling_prepare_v2(coor, [0, 0, 0], [1, 1, 1], [4.5, 4, 1])      # non-integral grid
    with pytest.raises(RuntimeError):
        rcb.voxel_pooling_prepare_v2(coor, [0, 0, 0], [1, 1, 1], [4096, 4096, 2])  # > 2^24 cells
    with pytest.raises(RuntimeError):
        rcb.voxel_pooling_prepare_v2(coor.cpu(), [0, 0, 0], [1, 1, 1], [4, 4, 1])  # no CPU fallback


def test_install_on_view_transformer_class():
    """install() gives a class the reference's method signature (self, coor)."""
    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import rig

    class FakeVT:
        def __init__(self):
            self.grid_lower_bound, self.grid_interval, self.grid_size = rig.grid_tensors(rig.R50_GRID)

    rcb.install(FakeVT)
    coor = rig.lidar_coor(rig.camera_rig(1), [1.0, 60.0, 4.0], (64, 176), 16).cuda()
    out = FakeVT().voxel_pooling_prepare_v2(coor)
    want = oracle.voxel_pooling_prepare_v2(coor.cpu().numpy(), *(t.numpy() for t in rig.grid_tensors(rig.R50_GRID)))
    assert_same_ranks(out, want)
