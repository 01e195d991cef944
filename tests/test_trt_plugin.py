"""the deployment surface of the op.  CPU: the ONNX symbolic of TRTBEVPoolv2 emits
the node the reference emits (bev_pool.py:98-119).  GPU: the plugin-shaped C entry
(rcb_trt_bev_pool_v2_enqueue: arrays of device pointers in ONNX input order, caller workspace, stream)
against the oracle and against the eager TRTBEVPoolv2.forward, for both rank contracts."""
import ctypes

import numpy as np
import pytest
import torch


class _FakeGraph:
    def __init__(self):
        self.calls = []

    def op(self, name, *inputs, **attrs):
        self.calls.append((name, inputs, attrs))
        return "node"


def test_symbolic_emits_the_reference_node():
    from rcbevdet_b200 import TRTBEVPoolv2
    g = _FakeGraph()
    names = ("depth", "feat", "ranks_depth", "ranks_feat", "ranks_bev", "interval_starts", "interval_lengths")
    assert TRTBEVPoolv2.symbolic(g, *names, 200, 176) == "node"
    (name, inputs, attrs), = g.calls
    assert name == "mmdeploy::bev_pool_v2"                      # bev_pool.py:110
    assert inputs == names                                      # :111-117, positional order
    assert attrs == {"out_height_i": 200, "out_width_i": 176}   # :118-119
    g2 = _FakeGraph()
    TRTBEVPoolv2.symbolic(g2, *names)
    assert g2.calls[0][2] == {"out_height_i": 128, "out_width_i": 128}   # defaults :106-107


def _inputs(n_cams=6, input_size=(128, 352), C=80, seed=3):
    import rcbevdet_b200 as rcb
    from rcbevdet_b200 import rig
    coor = rig.lidar_coor(rig.camera_rig(1, input_size=input_size, aug_seed=seed), [1.0, 60.0, 1.0], input_size, 16)
    _, N, D, H, W, _ = coor.shape
    depth, feat = rig.pooling_inputs(1, N, D, H, W, C, seed=seed)
    lo, iv, sz = rig.grid_tensors(rig.R50_GRID)
    ranks = rcb.voxel_pooling_prepare_v2(coor.cuda(), lo, iv, sz)
    return depth[0].cuda(), feat[0].permute(0, 2, 3, 1).contiguous().cuda(), ranks, (N, D, H, W)


def _enqueue(depth, feat_cl, ranks, dims, out_hw, sorted_cells, feat_dtype=0):
    from rcbevdet_b200 import _lib
    lib = _lib.lib()
    rb, rd, rf, st, ln = ranks
    N, D, H, W = dims
    C = feat_cl.shape[-1]
    out = torch.full((1, out_hw[0], out_hw[1], C), float("nan"), device="cuda")
    ws_bytes = lib.rcb_trt_bev_pool_v2_workspace_bytes(out_hw[0], out_hw[1], sorted_cells)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device="cuda")
    ins = (ctypes.c_void_p * 7)(*[t.data_ptr() for t in (depth, feat_cl, rd, rf, rb, st, ln)])
    outs = (ctypes.c_void_p * 1)(out.data_ptr())
    rc = lib.rcb_trt_bev_pool_v2_enqueue(ins, outs, N, D, H, W, C, rd.numel(), st.numel(), out_hw[0], out_hw[1],
                                         feat_dtype, sorted_cells, ctypes.c_void_p(ws.data_ptr()), ws_bytes, 0,
                                         _lib.stream_ptr(out.device))
    torch.cuda.synchronize()
    return rc, out


@pytest.mark.gpu
@pytest.mark.parametrize("sorted_cells", [0, 1])
def test_enqueue_matches_oracle_and_eager_twin(sorted_cells):
    import rcbevdet_b200 as rcb
    from oracle import oracle
    depth, feat_cl, ranks, dims = _inputs()
    rc, out = _enqueue(depth, feat_cl, ranks, dims, (128, 128), sorted_cells)
    assert rc == 0
    rb, rd, rf, st, ln = (t.cpu().numpy() for t in ranks)
    want = oracle.bev_pool_v2_forward(depth.cpu().numpy()[None], feat_cl.cpu().numpy()[None], rd, rf, rb,
                                      (1, 1, 128, 128, feat_cl.shape[-1]), st, ln, threads=8)[:, 0]
    assert out.shape == want.shape
    assert float(np.abs(out.cpu().numpy() - want).max()) <= 1e-5 * float(np.abs(want).max())
    eager = rcb.TRTBEVPoolv2.apply(depth, feat_cl, ranks[1], ranks[2], ranks[0], ranks[3], ranks[4], 128, 128)
    assert float((out - eager).abs().max()) <= 1e-5 * float(eager.abs().max())


@pytest.mark.gpu
def test_enqueue_general_path_takes_shuffled_intervals():
    """sorted_cells = 0 promises nothing about the ranks: intervals in any order give the same BEV."""
    depth, feat_cl, ranks, dims = _inputs(seed=5)
    rc, ref = _enqueue(depth, feat_cl, ranks, dims, (128, 128), 1)
    assert rc == 0
    rb, rd, rf, st, ln = ranks
    perm = torch.randperm(st.numel(), device="cuda", generator=torch.Generator("cuda").manual_seed(1))
    rc, out = _enqueue(depth, feat_cl, (rb, rd, rf, st[perm].contiguous(), ln[perm].contiguous()), dims, (128, 128), 0)
    assert rc == 0 and float((out - ref).abs().max()) <= 1e-5 * float(ref.abs().max())


@pytest.mark.gpu
def test_enqueue_fp16_context_and_errors():
    from rcbevdet_b200 import _lib
    depth, feat_cl, ranks, dims = _inputs(C=64, seed=6)
    rc, ref = _enqueue(depth, feat_cl, ranks, dims, (128, 128), 1)
    rc16, out16 = _enqueue(depth, feat_cl.half(), ranks, dims, (128, 128), 1, feat_dtype=_lib.DTYPE_F16)
    assert rc == 0 and rc16 == 0
    assert float((out16 - ref).abs().max()) <= 1e-2 * float(ref.abs().max())
    lib = _lib.lib()
    assert lib.rcb_trt_bev_pool_v2_enqueue(None, None, 1, 1, 1, 1, 4, 0, 0, 8, 8, 0, 0, None, 0, 0, None) == _lib.RCB_OK - 1
    N, D, H, W = dims
    ins = (ctypes.c_void_p * 7)(*[t.data_ptr() for t in (depth, feat_cl, ranks[1], ranks[2], ranks[0], ranks[3], ranks[4])])
    out = torch.empty((1, 128, 128, 64), device="cuda")
    outs = (ctypes.c_void_p * 1)(out.data_ptr())
    rc = lib.rcb_trt_bev_pool_v2_enqueue(ins, outs, N, D, H, W, 64, ranks[1].numel(), ranks[3].numel(), 128, 128, 0, 1,
                                         None, 0, 0, _lib.stream_ptr(out.device))
    assert rc == -2   # RCB_ERR_WORKSPACE: sorted_cells needs the CSR workspace
