#!/usr/bin/env python
"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN CODE on the CPU.

Run in the build container only (needs /root/reference):
    python oracle/gen_golden.py

The reference publishes exactly one golden vector for this path (the KAT in
mmdet3d/ops/bev_pool_v2/bev_pool.py:145-176, restated in tests/test_oracle.py).
Nothing pins voxel_pooling_prepare_v2 or the RCS radar scatter, so their parity is
pinned here: the unmodified reference classes are imported by path (oracle/refload.py)
and run on seeded synthetic inputs; inputs + outputs are committed as small fixtures.
Outputs whose order the reference leaves unspecified (argsort ties,
view_transformer.py:250) are stored both raw and canonicalised.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import oracle, refload  # noqa: E402
from rcbevdet_b200 import rig  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def _np(t):
    return None if t is None else t.detach().cpu().numpy()


def ref_prepare(vt_mod, grid, input_size, downsample, coor):
    m = vt_mod.LSSViewTransformer(grid_config=grid, input_size=input_size, downsample=downsample,
                                  in_channels=8, out_channels=8)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = m.voxel_pooling_prepare_v2(coor)
    return m, [_np(o) for o in out]


def prepare_case(vt_mod, name, grid, input_size, downsample, coor=None, rig_kwargs=None, batch=1):
    m = vt_mod.LSSViewTransformer(grid_config=grid, input_size=input_size, downsample=downsample,
                                  in_channels=8, out_channels=8)
    case = {}
    if coor is None:
        calib = rig.camera_rig(batch, input_size=input_size, **(rig_kwargs or {}))
        coor = m.get_lidar_coor(*calib)
        for k, t in zip(("sensor2ego", "ego2global", "intrin", "post_rot", "post_tran", "bda"), calib):
            case[f"{name}.{k}"] = _np(t)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = m.voxel_pooling_prepare_v2(coor)
    case[f"{name}.coor"] = _np(coor).astype(np.float32)
    case[f"{name}.lower"] = _np(m.grid_lower_bound)
    case[f"{name}.interval"] = _np(m.grid_interval)
    case[f"{name}.size"] = _np(m.grid_size)
    case[f"{name}.frustum"] = _np(m.frustum)
    names = ("ranks_bev", "ranks_depth", "ranks_feat", "interval_starts", "interval_lengths")
    if out[0] is None:
        case[f"{name}.empty"] = np.array(1)
        return case
    case[f"{name}.empty"] = np.array(0)
    out = [_np(o) for o in out]
    for k, o in zip(names, out):
        assert o.dtype == np.int32, (k, o.dtype)
        case[f"{name}.raw.{k}"] = o
    rb, rd, rf = oracle.canonicalise(out[0], out[1], out[2])
    case[f"{name}.canon.ranks_bev"] = rb
    case[f"{name}.canon.ranks_depth"] = rd
    case[f"{name}.canon.ranks_feat"] = rf
    print(f"  {name}: P={coor.numel() // 3} K={out[0].shape[0]} I={out[3].shape[0]}")
    return case


def gen_prepare():
    vt = refload.load_view_transformer()
    cases = {}
    g128 = dict(rig.R50_GRID)
    # A: deterministic rig, reduced image + coarse depth bins, the R50 BEV grid
    gA = dict(g128, depth=[1.0, 60.0, 5.0])
    cases.update(prepare_case(vt, "rigA", gA, (64, 176), 16, batch=2))
    # B: random coordinates straddling every boundary, Z > 1, non-square grid
    gB = dict(x=[-8.0, 8.0, 1.0], y=[-6.0, 6.0, 0.5], z=[-5.0, 3.0, 4.0], depth=[1.0, 5.0, 1.0])
    gen = torch.Generator().manual_seed(11)
    coor = torch.rand(2, 2, 4, 3, 5, 3, generator=gen)
    coor = coor * torch.tensor([20.0, 16.0, 12.0]) - torch.tensor([10.0, 8.0, 7.0])
    # plant exact-boundary and (-1, 0) voxel-coordinate values (.long() truncation keeps them)
    flat = coor.view(-1, 3)
    flat[0] = torch.tensor([-8.0, -6.0, -5.0])
    flat[1] = torch.tensor([8.0, 0.0, 0.0])          # x == upper bound -> dropped
    flat[2] = torch.tensor([-8.5, 0.0, 0.0])         # voxel x = -0.5 -> truncates to 0 -> kept
    flat[3] = torch.tensor([0.0, -6.4, -8.9])        # voxel y=-0.8, z=-0.975 -> kept
    flat[4] = torch.tensor([0.0, 0.0, -9.0])         # voxel z = -1.0 -> dropped
    flat[5] = torch.tensor([7.999999, 5.999999, 2.999999])
    cases.update(prepare_case(vt, "randB", gB, (48, 80), 16, coor=coor))
    # C: nothing inside the grid -> five Nones
    coorC = torch.full((1, 1, 2, 2, 2, 3), 500.0)
    cases.update(prepare_case(vt, "emptyC", gB, (32, 32), 16, coor=coorC))
    # D: seeded train-time augmentation (image rot/flip/resize + BEV rot/scale/flip)
    gD = dict(g128, depth=[1.0, 60.0, 2.0])
    cases.update(prepare_case(vt, "augD", gD, (128, 352), 16, rig_kwargs=dict(aug_seed=7), batch=1))
    # E: one interval only (all points in one cell), and a single kept point
    coorE = torch.zeros(1, 1, 3, 2, 2, 3)
    coorE[..., 0] = 0.25
    coorE[..., 1] = 0.1
    cases.update(prepare_case(vt, "onecellE", gB, (32, 32), 16, coor=coorE))
    coorF = torch.full((1, 2, 2, 2, 2, 3), 500.0)
    coorF[0, 1, 1, 0, 1] = torch.tensor([1.5, -2.2, 0.3])
    cases.update(prepare_case(vt, "singleF", gB, (32, 32), 16, coor=coorF))
    np.savez_compressed(os.path.join(GOLD, "prepare_ref.npz"), **cases)

    # full-size digest (config 1: B=1 R50 rig) -- too big to commit, so digests only
    m = vt.LSSViewTransformer(grid_config=g128, input_size=rig.R50_INPUT, downsample=16,
                              in_channels=8, out_channels=8)
    calib = rig.camera_rig(1)
    coor = m.get_lidar_coor(*calib)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = [_np(o) for o in m.voxel_pooling_prepare_v2(coor)]
    rb, rd, rf = oracle.canonicalise(out[0], out[1], out[2])
    digest = {
        "config": "B=1 N=6 D=118 H=16 W=44 grid 128x128x1 (SURVEY.md Appendix B rig)",
        "P": int(coor.numel() // 3), "K": int(out[0].shape[0]), "I": int(out[3].shape[0]),
        "sha256": {
            "coor": hashlib.sha256(_np(coor).astype(np.float32).tobytes()).hexdigest(),
            "ranks_bev": hashlib.sha256(rb.tobytes()).hexdigest(),
            "ranks_depth": hashlib.sha256(rd.tobytes()).hexdigest(),
            "ranks_feat": hashlib.sha256(rf.tobytes()).hexdigest(),
            "interval_starts": hashlib.sha256(out[3].tobytes()).hexdigest(),
            "interval_lengths": hashlib.sha256(out[4].tobytes()).hexdigest(),
        },
    }
    print("  full config-1:", digest["P"], digest["K"], digest["I"])
    with open(os.path.join(GOLD, "prepare_full_digest.json"), "w") as f:
        json.dump(digest, f, indent=1)


class _Capture(torch.nn.Module):
    def __init__(self, out_channels):
        super().__init__()
        self.seen = None
        self.out_channels = out_channels

    def forward(self, x):
        self.seen = x.detach().clone()
        return x.new_zeros((x.shape[0], self.out_channels) + tuple(x.shape[2:]))


def radar_case(ps, name, batch, ny, nx, pts, cin, seed, rcs_scale=1.0):
    feats, rcs, coors = rig.radar_pillars(batch, ny, nx, points_per_sample=pts, in_channels=cin, seed=seed)
    rcs = rcs.clone()
    rcs[:, 5] *= rcs_scale
    mod = ps.PointPillarsScatterRCS(in_channels=cin, output_shape=[ny, nx])
    mod.rcs_att = _Capture(cin)       # sees cat[heatmap, heatmap_feat]   (pillar_scatter.py:132)
    mod.compress = _Capture(cin)      # sees cat[features, rcs_att]       (pillar_scatter.py:134)
    with torch.no_grad():
        mod((feats, rcs), coors.long(), batch_size=batch)
    hm = mod.rcs_att.seen
    out = {
        f"{name}.point_features": _np(feats), f"{name}.rcs": _np(rcs), f"{name}.coors": _np(coors),
        f"{name}.shape": np.array([batch, ny, nx, cin]),
        f"{name}.heatmap": _np(hm[:, 0]), f"{name}.heatmap_feat": _np(hm[:, 1:2]),
        f"{name}.features": _np(mod.compress.seen[:, :cin]),
    }
    print(f"  {name}: V={feats.shape[0]} radius max={max(oracle.rcs_radius(_np(rcs)))}")
    return out


def gen_radar():
    ps = refload.load_pillar_scatter()
    cases = {}
    cases.update(radar_case(ps, "r32", 2, 32, 32, 40, 8, seed=5, rcs_scale=0.3))
    cases.update(radar_case(ps, "r64", 2, 48, 64, 150, 16, seed=6))
    cases.update(radar_case(ps, "r128", 1, 128, 128, 400, 4, seed=8))
    np.savez_compressed(os.path.join(GOLD, "radar_ref.npz"), **cases)


def temporal_inputs(n, C, h, w, seed, with_bda_adj):
    """Key / adjacent frame calibration of the deterministic rig (one camera is enough: gen_grid reads
    camera 0 only), seeded BEV augmentation, seeded features."""
    import math
    from rcbevdet_b200 import rig
    g = torch.Generator().manual_seed(seed)
    motion = rig.temporal_motion(n, 2, seed=seed).view(n, 2, 3)
    key = rig.camera_rig(n, frame_motion=motion[:, 0])[0]          # sensor2ego of the key frame (n, 6, 4, 4)
    adj = rig.camera_rig(n, frame_motion=motion[:, 1] * 3.0)[0]    # adjacent frame, exaggerated ego motion
    def bda(k):
        ang = (torch.rand(n, generator=g) - 0.5) * 0.6
        sc = 0.95 + 0.1 * torch.rand(n, generator=g)
        m = torch.zeros(n, 3, 3)
        m[:, 0, 0], m[:, 0, 1], m[:, 1, 0], m[:, 1, 1], m[:, 2, 2] = torch.cos(ang), -torch.sin(ang), torch.sin(ang), torch.cos(ang), 1.0
        m = m * sc.view(n, 1, 1)
        if k:
            m[0, 1] = -m[0, 1]      # a flipped sample
        return m
    feats = torch.randn(n, C, h, w, generator=g)
    return feats, [key, adj], bda(0), (bda(1) if with_bda_adj else None)


def gen_temporal():
    """BEVDepth4D.gen_grid / shift_feature (bevdet_rc.py:585-657) executed from the reference file."""
    lo, iv = [-51.2, -51.2, -5.0], [0.8, 0.8, 8.0]
    cases = {}
    for name, (n, C, h, w, seed, adj) in {"t16": (2, 3, 16, 16, 3, False), "t32": (3, 5, 24, 32, 4, True),
                                          "t128": (1, 2, 128, 128, 5, True)}.items():
        scale = 128.0 / w                       # the 102.4 m grid at this resolution
        iv_c = [iv[0] * scale, iv[1] * 128.0 / h, iv[2]]
        ref = refload.load_temporal_alignment(iv_c, lo)
        feats, s2k, bda, bda_adj = temporal_inputs(n, C, h, w, seed, adj)
        grid = ref.gen_grid(feats, s2k, bda, bda_adj=bda_adj)
        out = ref.shift_feature(feats, s2k, bda, bda_adj=bda_adj)
        cases.update({f"{name}.input": _np(feats), f"{name}.key": _np(s2k[0]), f"{name}.adj": _np(s2k[1]),
                      f"{name}.bda": _np(bda), f"{name}.interval": np.float32(iv_c), f"{name}.lower": np.float32(lo),
                      f"{name}.grid": _np(grid), f"{name}.output": _np(out)})
        if bda_adj is not None:
            cases[f"{name}.bda_adj"] = _np(bda_adj)
        print(name, tuple(out.shape), "covered", float((out.abs().sum(1) > 0).float().mean()))
    np.savez_compressed(os.path.join(GOLD, "temporal_ref.npz"), **cases)


if __name__ == "__main__":
    assert refload.available(), "reference tree not found"
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    print("prepare:")
    gen_prepare()
    print("radar:")
    gen_radar()
    print("temporal:")
    gen_temporal()
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))
