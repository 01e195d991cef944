"""One B=8 R50 step through the public chain API (rcb.voxel_pooling_v2 + backward), strips mode from
argv[1] (off | chain | ...), argv[2] steps.  Run under ncu for a launch list: tools/chain_times.sh."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rcbevdet_b200 as rcb  # noqa: E402
from rcbevdet_b200 import rig, strips  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "chain"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
from_calib = len(sys.argv) > 3 and sys.argv[3] == "calib"
strips.set_mode(mode)
B, C = int(os.environ.get("RCB_B", "8")), 80
grid = rig.R50_GRID
calib = rig.camera_rig(B, aug_seed=3)
coor = rig.lidar_coor(calib, grid["depth"], rig.R50_INPUT, 16).cuda()
axes = rcb.frustum_axes(grid["depth"], rig.R50_INPUT, 16, device="cuda")
packed = tuple(t.cuda() for t in rcb.pack_calib(*calib))
depth, feat = (t.cuda() for t in rig.pooling_inputs(B, 6, 118, 16, 44, C, seed=4))
og = torch.randn(B, C, 128, 128, device="cuda")
lo, iv, sz = rig.grid_tensors(grid)
for _ in range(steps):
    d = depth.detach().requires_grad_(True)
    f = feat.detach().requires_grad_(True)
    if from_calib:
        bev = rcb.voxel_pooling_v2_from_calib(packed, axes, d, f, lo, iv, sz)
    else:
        bev = rcb.voxel_pooling_v2(coor, d, f, lo, iv, sz)
    bev.backward(og)
torch.cuda.synchronize()
print("checksum", float(bev.double().sum()))
