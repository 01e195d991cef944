"""Development check of the strip kernels against the cell-/pixel-stationary ones (same inputs,
same plan) + timings.  python tools/dev_strips.py [B]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rcbevdet_b200 as rcb  # noqa: E402
from rcbevdet_b200 import _lib, bev_pool as bp, rig, strips  # noqa: E402
from rcbevdet_b200 import plan as _plan  # noqa: E402
from rcbevdet_b200.prepare import prepare_async  # noqa: E402


def timed(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


def run(B, grid, input_size, depth_cfg, C=80, aug=None, layout=_lib.LAYOUT_B_C_CELLS, dtype=torch.float32, time=False):
    dev = torch.device("cuda", 0)
    coor = rig.lidar_coor(rig.camera_rig(B, input_size=input_size, aug_seed=aug), list(depth_cfg), input_size, 16).to(dev)
    _, N, D, H, W, _ = coor.shape
    depth, feat = rig.pooling_inputs(B, N, D, H, W, C, seed=1)
    depth = depth.to(dev)
    rows = feat.permute(0, 1, 3, 4, 2).contiguous().view(-1, C).to(dev).to(dtype)
    lo, iv, sz = rig.grid_tensors(grid)
    r = prepare_async(coor, lo, iv, sz)
    gz, gy, gx = r.grid
    d = _lib.PoolDesc()
    d.n_points, d.n_intervals, d.C = r.P, 0, C
    d.B, d.Z, d.Y, d.X = B, gz, gy, gx
    d.n_depth, d.n_pixels = depth.numel(), rows.shape[0]
    d.D, d.HW, d.H = D, H * W, H
    d.layout, d.feat_dtype, d.flags = layout, bp._DTYPES[dtype], _lib.PLAN_ALL
    shape = (B, C, gz, gy, gx) if layout == _lib.LAYOUT_B_C_CELLS else (B, gz, gy, gx, C)
    out_ref = torch.empty(shape, device=dev)
    bp.pool_forward(d, depth, rows, r.ranks_depth, r.ranks_feat, r.ranks_bev, None, None, r.cell_start, out_ref)
    sp = strips.build(r.point_cell, r.cell_start, B * N, D, H, W, r.n_cells)
    al = lambda v: (v + 255) // 256 * 256
    off = 256 + al(r.n_cells * 4) + al((r.n_cells + 255) // 256 * 4)
    nseg = sp.buf[off:off + 4 * B * N * ((H + 15) // 16) * W].view(torch.int32)
    print(f"B={B} {input_size} D={D} H={H} W={W} C={C} aug={aug} layout={layout} {dtype}: status={sp.status()} "
          f"segments={int(nseg.sum())} (max {int(nseg.max())}) kept={int(r.counts[0])}")
    out = torch.full(shape, float("nan"), device=dev)
    strips.forward(sp, d, depth, rows, out)
    torch.cuda.synchronize()
    scale = float(out_ref.abs().max())
    err = float((out - out_ref).abs().max())
    print(f"   fwd: max abs err {err:.3e} (scale {scale:.3e}, rel {err / scale:.2e}) nan={int(torch.isnan(out).sum())}")
    ok = err <= 1e-5 * scale if dtype == torch.float32 else err <= 1e-2 * scale
    # backward
    g = torch.Generator().manual_seed(7)
    og = torch.randn(shape, generator=g).to(dev)
    plan = _plan.PoolPlan(_lib.PLAN_ALL, r.cell_start, r.point_cell, D, H * W, r.n_cells, r.P)
    saved = (depth, rows, r.ranks_depth, r.ranks_feat, r.ranks_bev)
    dg_ref, fg_ref = bp._backward(og, saved, d, plan, tuple(rows.shape), torch.float32, tuple(depth.shape), torch.float32,
                                  layout == _lib.LAYOUT_B_C_CELLS)
    if C in strips.BWD_CHANNELS:
        dg = torch.full_like(depth, float("nan"))
        fg = torch.full((rows.shape[0], C), float("nan"), device=dev)
        strips.backward(sp, d, og, depth, rows, dg, fg)
        torch.cuda.synchronize()
        e1 = float((dg - dg_ref).abs().max()) / float(dg_ref.abs().max())
        e2 = float((fg - fg_ref).abs().max()) / float(fg_ref.abs().max())
        print(f"   bwd: depth_grad rel {e1:.2e} nan={int(torch.isnan(dg).sum())}; feat_grad rel {e2:.2e} nan={int(torch.isnan(fg).sum())}")
        tol = 1e-5 if dtype == torch.float32 else 1e-2
        ok = ok and e1 <= tol and e2 <= tol
    if time:
        t_plan = timed(lambda: strips.build(r.point_cell, r.cell_start, B * N, D, H, W, r.n_cells))
        t_old = timed(lambda: bp.pool_forward(d, depth, rows, r.ranks_depth, r.ranks_feat, r.ranks_bev, None, None,
                                              r.cell_start, out_ref))
        t_new = timed(lambda: strips.forward(sp, d, depth, rows, out))
        print(f"   time: plan {t_plan:.1f} us; fwd cells {t_old:.1f} us, strips {t_new:.1f} us")
        if C in strips.BWD_CHANNELS:
            t_bo = timed(lambda: bp._backward(og, saved, d, plan, tuple(rows.shape), torch.float32, tuple(depth.shape),
                                              torch.float32, layout == _lib.LAYOUT_B_C_CELLS))
            t_bn = timed(lambda: strips.backward(sp, d, og, depth, rows, dg, fg))
            print(f"   time: bwd pixels (+transpose, +allocs) {t_bo:.1f} us, strips {t_bn:.1f} us")
    return ok


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    if len(sys.argv) > 2 and sys.argv[2] == "one":      # the B-sample R50 case only (for ncu)
        aug = int(sys.argv[3]) if len(sys.argv) > 3 else None
        ok = run(B, rig.R50_GRID, rig.R50_INPUT, rig.R50_GRID["depth"], aug=aug, time=True)
        sys.exit(0 if ok or os.environ.get("RCB_DEV_NOCHECK") else 1)
    if len(sys.argv) > 2 and sys.argv[2] == "hires":    # the 900 x 1600 / 256^2 case only, B samples
        ok = run(B, rig.HIRES_GRID, rig.HIRES_INPUT, rig.HIRES_GRID["depth"], time=True)
        sys.exit(0 if ok or os.environ.get("RCB_DEV_NOCHECK") else 1)
    ok = True
    ok &= run(1, rig.R50_GRID, (64, 176), (1.0, 60.0, 4.0))
    ok &= run(2, rig.R50_GRID, (128, 352), (1.0, 60.0, 2.0), aug=3)
    ok &= run(2, rig.R50_GRID, (128, 352), (1.0, 60.0, 2.0), aug=5, layout=_lib.LAYOUT_CELLS_C)
    ok &= run(2, rig.R50_GRID, (112, 304), (1.0, 60.0, 1.0), aug=4, C=64)       # H = 7, W = 19: ragged strips
    ok &= run(1, rig.R50_GRID, (128, 352), (1.0, 60.0, 2.0), C=128)
    ok &= run(1, rig.R50_GRID, (128, 352), (1.0, 60.0, 2.0), aug=2, dtype=torch.bfloat16)
    ok &= run(2, rig.R50_GRID, rig.R50_INPUT, rig.R50_GRID["depth"], aug=3)
    ok &= run(1, rig.HIRES_GRID, rig.HIRES_INPUT, rig.HIRES_GRID["depth"], time=True)
    ok &= run(B, rig.R50_GRID, rig.R50_INPUT, rig.R50_GRID["depth"], time=True)
    ok &= run(B, rig.R50_GRID, rig.R50_INPUT, rig.R50_GRID["depth"], aug=3, time=True)
    print("ALL OK" if ok else "MISMATCH")
    sys.exit(0 if ok else 1)
