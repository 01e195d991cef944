"""Stage times on the BASELINE config-5 geometry (900x1600 -> 56x100 features, D=118, C=80,
256x256 BEV).  Not a bench line: a sanity check that nothing degenerates at the larger size
(cells of up to ~2000 points, 4x the points per sample).  Run on the GPU box from the repo root."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rcbevdet_b200 as rcb  # noqa: E402
from rcbevdet_b200 import rig  # noqa: E402
from rcbevdet_b200.prepare import prepare_async  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
grid = rig.HIRES_GRID
coor = rig.lidar_coor(rig.camera_rig(B, input_size=rig.HIRES_INPUT), grid["depth"], rig.HIRES_INPUT, 16).cuda()
_, N, D, H, W, _ = coor.shape
depth, feat = rig.pooling_inputs(B, N, D, H, W, 80, seed=1)
depth, feat = depth.cuda(), feat.cuda()
lo, iv, sz = rig.grid_tensors(grid)
og = torch.randn(B, 80, 256, 256, device="cuda")


def timed(fn, n=10):
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


def fwd_bwd():
    d = depth.detach().requires_grad_(True)
    f = feat.detach().requires_grad_(True)
    rcb.voxel_pooling_v2(coor, d, f, lo, iv, sz).backward(og)


t_prep = timed(lambda: prepare_async(coor, lo, iv, sz))
t_all = timed(fwd_bwd)
r = prepare_async(coor, lo, iv, sz)
k, i = r.counts[:2].tolist()
print(f"hires B={B}: P={coor.numel() // 3} K={k} I={i} max interval {int(r.interval_lengths[:i].max())}; "
      f"prepare {t_prep:.3f} ms, prepare+fwd+bwd (autograd API) {t_all:.3f} ms -> {B / t_all * 1e3:.0f} samples/s")
