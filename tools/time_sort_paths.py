"""Prepare on either side of the two-level / LSD boundary (2^22 cells), same 4 M points: states the
cliff DESIGN.md section 3.1 mentions.  Run on the GPU box from the repo root."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rcbevdet_b200.prepare import prepare_async  # noqa: E402

g = torch.Generator("cuda").manual_seed(1)
P = 4_000_000
xy = torch.rand(1, 1, P, 1, 1, 2, device="cuda", generator=g)
for gx, gy in ((128, 128), (1024, 1024), (2048, 2048), (2100, 2000), (4096, 4096)):
    coor = torch.cat((xy[..., :1] * gx, xy[..., 1:] * gy, torch.full_like(xy[..., :1], 0.5)), -1).contiguous()
    lo, iv, sz = [0.0, 0.0, 0.0], [1.0, 1.0, 1.0], [float(gx), float(gy), 1.0]
    for _ in range(3):
        r = prepare_async(coor, lo, iv, sz)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        r = prepare_async(coor, lo, iv, sz)
    e1.record()
    torch.cuda.synchronize()
    k, i = r.counts[:2].tolist()
    print(f"grid {gx}x{gy} = {gx * gy / 2**20:.2f} Mi cells, {P} points (D = P: depth-blocked tiles), kept {k}, "
          f"intervals {i}: {e0.elapsed_time(e1) / 10 * 1e3:.0f} us")
