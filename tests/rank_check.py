"""Comparison of prepare outputs with the oracle / reference.

`ranks_bev`, `interval_starts`, `interval_lengths` do not depend on how ties inside a BEV cell are
ordered: bit-exact as they come.  `ranks_depth` / `ranks_feat` are compared after the lossless
canonical ordering by (ranks_bev, ranks_depth) -- the reference's argsort (view_transformer.py:250)
leaves the order inside a cell unspecified; the oracle uses the stable order, this library orders
a cell by (pixel, depth bin).  That library order is checked as a property."""
import numpy as np

from oracle import oracle

NAMES = ("ranks_bev", "ranks_depth", "ranks_feat", "interval_starts", "interval_lengths")


def to_numpy(t):
    return t if isinstance(t, np.ndarray) else t.detach().cpu().numpy()


def assert_same_ranks(got, want, check_order=True):
    """got: the 5-tuple of this library (tensors or arrays); want: the oracle's / reference's."""
    if want[0] is None:
        assert all(g is None for g in got), "expected five Nones (nothing inside the grid)"
        return
    g = [to_numpy(t) for t in got]
    w = [to_numpy(t) for t in want]
    for name, a in zip(NAMES, g):
        assert a.dtype == np.int32, (name, a.dtype)
    assert np.array_equal(g[0], w[0]), "ranks_bev"
    assert np.array_equal(g[3], w[3]), "interval_starts"
    assert np.array_equal(g[4], w[4]), "interval_lengths"
    cg = oracle.canonicalise(g[0], g[1], g[2])
    cw = oracle.canonicalise(w[0], w[1], w[2])
    assert np.array_equal(cg[1], cw[1]), "ranks_depth (canonical order)"
    assert np.array_equal(cg[2], cw[2]), "ranks_feat (canonical order)"
    if check_order and g[0].shape[0] > 1:
        # library order inside a cell: ascending (ranks_feat, ranks_depth)
        same_cell = g[0][1:] == g[0][:-1]
        feat_up = g[2][1:] > g[2][:-1]
        feat_eq = g[2][1:] == g[2][:-1]
        depth_up = g[1][1:] > g[1][:-1]
        ok = ~same_cell | feat_up | (feat_eq & depth_up)
        assert bool(ok.all()), "points of a cell must be ordered by (pixel, depth bin)"
