"""Synthetic inputs for the pooling path: a deterministic nuScenes-like 6-camera rig
(SURVEY.md Appendix B), a seeded train-time augmentation variant, per-frame ego motion
for the 4D (temporal) configuration, and synthetic radar pillars.

Everything here is plain PyTorch on whatever device is asked for; it only *produces
inputs* (calibration matrices, depth/context tensors, radar pillars).  There is no
dataset and no checkpoint in this environment, so benchmarks and tests use these.

Shapes follow the reference's config (configs/rcbevdet/rcbevdet-256x704-r50-BEV128-
9kf-depth-cbgs12e-circlelarger.py:21-48): cameras in the order FRONT_LEFT, FRONT,
FRONT_RIGHT, BACK_LEFT, BACK, BACK_RIGHT.
"""
from __future__ import annotations

import math

import torch

CAM_YAW_DEG = (55.0, 0.0, -55.0, 110.0, 180.0, -110.0)

# grid / image configurations named in BASELINE.json
R50_GRID = dict(x=[-51.2, 51.2, 0.8], y=[-51.2, 51.2, 0.8], z=[-5, 3, 8], depth=[1.0, 60.0, 0.5])
R50_INPUT = (256, 704)
HIRES_GRID = dict(x=[-51.2, 51.2, 0.4], y=[-51.2, 51.2, 0.4], z=[-5, 3, 8], depth=[1.0, 60.0, 0.5])
HIRES_INPUT = (900, 1600)


def _rz(yaw):
    c, s = math.cos(yaw), math.sin(yaw)
    return torch.tensor([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]], dtype=torch.float64)


# camera (x right, y down, z forward) -> ego (x forward, y left, z up)
_CAM2EGO_AXES = torch.tensor([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]], dtype=torch.float64)


def camera_rig(batch, input_size=R50_INPUT, src_size=(900, 1600), device="cpu", aug_seed=None,
               frame_motion=None):
    """Calibration for `batch` samples x 6 cameras.

    Returns (sensor2ego (B,6,4,4), ego2global (B,6,4,4), intrin (B,6,3,3),
    post_rot (B,6,3,3), post_tran (B,6,3), bda (B,3,3)), float32 on `device`: the
    argument tuple of LSSViewTransformer.get_lidar_coor (view_transformer.py:115).

    aug_seed=None -> test-time augmentation (loading.py:1672-1684: resize = W_in/1600,
    crop the top).  aug_seed=int -> seeded train-time image + BEV augmentation with the
    reference's ranges (configs/rcbevdet/...:31-35,215-219).
    frame_motion: optional (B,3) tensor (dx, dy, dyaw) applied to the ego pose of each
    sample - used to emulate adjacent temporal frames.
    """
    H_in, W_in = input_size
    src_h, src_w = src_size
    n = len(CAM_YAW_DEG)
    s2e = torch.zeros(batch, n, 4, 4, dtype=torch.float64)
    intrin = torch.zeros(batch, n, 3, 3, dtype=torch.float64)
    post_rot = torch.zeros(batch, n, 3, 3, dtype=torch.float64)
    post_tran = torch.zeros(batch, n, 3, dtype=torch.float64)
    bda = torch.zeros(batch, 3, 3, dtype=torch.float64)
    gen = None
    if aug_seed is not None:
        gen = torch.Generator().manual_seed(int(aug_seed))

    def u(lo, hi):
        return float(torch.rand((), generator=gen, dtype=torch.float64)) * (hi - lo) + lo

    for b in range(batch):
        motion = _rz(0.0)
        shift = torch.zeros(3, dtype=torch.float64)
        if frame_motion is not None:
            dx, dy, dyaw = (float(v) for v in frame_motion[b])
            motion = _rz(dyaw)
            shift = torch.tensor([dx, dy, 0.0], dtype=torch.float64)
        for k, yaw_deg in enumerate(CAM_YAW_DEG):
            yaw = math.radians(yaw_deg)
            rot = motion @ _rz(yaw) @ _CAM2EGO_AXES
            t = motion @ torch.tensor([1.5 * math.cos(yaw), 1.5 * math.sin(yaw), 1.5],
                                      dtype=torch.float64) + shift
            s2e[b, k, :3, :3] = rot
            s2e[b, k, :3, 3] = t
            s2e[b, k, 3, 3] = 1.0
            intrin[b, k] = torch.tensor([[1266.0, 0.0, 816.0], [0.0, 1266.0, 491.0], [0.0, 0.0, 1.0]],
                                        dtype=torch.float64)
            if gen is None:
                resize = float(W_in) / float(src_w)
                new_h = int(src_h * resize)
                crop_h = new_h - H_in
                crop_w = 0
                flip = False
                rotate = 0.0
            else:
                resize = float(W_in) / float(src_w) + u(-0.06, 0.11)
                new_w, new_h = int(src_w * resize), int(src_h * resize)
                crop_h = new_h - H_in
                crop_w = int(u(0.0, 1.0) * max(0, new_w - W_in))
                flip = u(0.0, 1.0) < 0.5
                rotate = u(-5.4, 5.4)
            pr = torch.eye(2, dtype=torch.float64) * resize
            pt = -torch.tensor([float(crop_w), float(crop_h)], dtype=torch.float64)
            if flip:
                a = torch.tensor([[-1.0, 0.0], [0.0, 1.0]], dtype=torch.float64)
                bvec = torch.tensor([float(W_in), 0.0], dtype=torch.float64)
                pr = a @ pr
                pt = a @ pt + bvec
            if rotate != 0.0:
                h = math.radians(rotate)
                a = torch.tensor([[math.cos(h), math.sin(h)], [-math.sin(h), math.cos(h)]],
                                 dtype=torch.float64)
                ctr = torch.tensor([W_in / 2.0, H_in / 2.0], dtype=torch.float64)
                bvec = a @ (-ctr) + ctr
                pr = a @ pr
                pt = a @ pt + bvec
            post_rot[b, k] = torch.eye(3, dtype=torch.float64)
            post_rot[b, k, :2, :2] = pr
            post_tran[b, k, :2] = pt
        if gen is None:
            bda[b] = torch.eye(3, dtype=torch.float64)
        else:
            ang = math.radians(u(-22.5, 22.5))
            scale = u(0.95, 1.05)
            m = _rz(ang) * scale
            if u(0.0, 1.0) < 0.5:
                m = torch.diag(torch.tensor([-1.0, 1.0, 1.0], dtype=torch.float64)) @ m
            if u(0.0, 1.0) < 0.5:
                m = torch.diag(torch.tensor([1.0, -1.0, 1.0], dtype=torch.float64)) @ m
            bda[b] = m
    e2g = torch.eye(4, dtype=torch.float64).expand(batch, n, 4, 4).clone()
    out = (s2e, e2g, intrin, post_rot, post_tran, bda)
    return tuple(t.to(torch.float32).to(device) for t in out)


def temporal_motion(batch, frames, seed=0):
    """Seeded SE(2) ego step per adjacent frame:
    returns (batch*frames, 3) of accumulated (dx, dy, dyaw); frame 0 is the key frame."""
    gen = torch.Generator().manual_seed(int(seed))
    step = torch.rand(batch, frames, 3, generator=gen, dtype=torch.float64)
    step[..., 0] = step[..., 0] * 1.0 - 2.0       # ~ -1..-2 m per frame backwards in time
    step[..., 1] = (step[..., 1] - 0.5) * 0.2
    step[..., 2] = (step[..., 2] - 0.5) * 0.04
    step[:, 0] = 0.0
    return torch.cumsum(step, dim=1).reshape(batch * frames, 3)


def pooling_inputs(batch, n_cams, D, H, W, C, device="cpu", seed=1, dtype=torch.float32):
    """depth = softmax(randn(seed), dim=2) as (B,N,D,H,W);
    context = randn(seed+1) as (B,N,C,H,W) (the caller passes its permuted view, exactly
    like view_transformer.py:195); out_grad = randn(seed+2) is made by the caller."""
    g = torch.Generator().manual_seed(int(seed))
    depth = torch.randn(batch, n_cams, D, H, W, generator=g).softmax(dim=2)
    g2 = torch.Generator().manual_seed(int(seed) + 1)
    feat = torch.randn(batch, n_cams, C, H, W, generator=g2)
    return depth.to(device), feat.to(device=device, dtype=dtype)


def radar_pillars(batch, ny, nx, points_per_sample=3125, in_channels=64, seed=4, device="cpu"):
    """per sample `points_per_sample` radar returns with
    xy ~ U(-51.2, 51.2), rcs ~ U(-5, 30), hard-voxelised to unique (y, x) pillars.

    Returns point_features (V,in_channels) f32, rcs (V,7) f32 with cols 0,1 = x,y
    normalised to [0,1] and col 5 = the RCS value (radar_encoder.py:372-375,450), and
    coors (V,4) int32 = [b, 0, y, x] (the layout mmcv's Voxelization + F.pad produce,
    bevdet_rc.py:170-196)."""
    g = torch.Generator().manual_seed(int(seed))
    feats, rcss, coors = [], [], []
    for b in range(batch):
        xy = torch.rand(points_per_sample, 2, generator=g)
        val = torch.rand(points_per_sample, generator=g) * 35.0 - 5.0
        ix = torch.clamp((xy[:, 0] * nx).long(), 0, nx - 1)
        iy = torch.clamp((xy[:, 1] * ny).long(), 0, ny - 1)
        lin = iy * nx + ix
        # keep the first point of every pillar, in arrival order (hard voxelisation)
        order = torch.argsort(lin, stable=True)
        lin_s = lin[order]
        first = torch.ones_like(lin_s, dtype=torch.bool)
        first[1:] = lin_s[1:] != lin_s[:-1]
        keep = torch.sort(order[first]).values
        v = keep.numel()
        r = torch.zeros(v, 7)
        r[:, 0] = xy[keep, 0]
        r[:, 1] = xy[keep, 1]
        r[:, 2] = 0.5
        r[:, 3] = torch.rand(v, generator=g)
        r[:, 4] = torch.rand(v, generator=g)
        r[:, 5] = val[keep]
        r[:, 6] = torch.rand(v, generator=g)
        c = torch.zeros(v, 4, dtype=torch.int32)
        c[:, 0] = b
        c[:, 2] = iy[keep].int()
        c[:, 3] = ix[keep].int()
        feats.append(torch.randn(v, in_channels, generator=g))
        rcss.append(r)
        coors.append(c)
    return (torch.cat(feats).to(device), torch.cat(rcss).to(device), torch.cat(coors).to(device))


def frustum(depth_cfg, input_size, downsample, device="cpu"):
    """(D, H, W, 3) float32 image-plane frustum (u, v, d): pixel centres on a linspace over the
    input image, depth bins arange(*depth_cfg) -- the template of view_transformer.py:85-113."""
    H_in, W_in = input_size
    Hf, Wf = H_in // downsample, W_in // downsample
    d = torch.arange(*depth_cfg, dtype=torch.float32, device=device)
    u = torch.linspace(0, W_in - 1, Wf, dtype=torch.float32, device=device)
    v = torch.linspace(0, H_in - 1, Hf, dtype=torch.float32, device=device)
    D = d.shape[0]
    return torch.stack((u.view(1, 1, Wf).expand(D, Hf, Wf), v.view(1, Hf, 1).expand(D, Hf, Wf),
                        d.view(D, 1, 1).expand(D, Hf, Wf)), -1)


def _apply3(m, p):
    """(..., 3, 3) applied to (..., D, H, W, 3) with explicit fp32 multiply-adds (no BLAS, so the
    result does not depend on the host's GEMM kernels)."""
    m = m[..., None, None, None, :, :]
    return (m[..., :, 0] * p[..., 0:1] + m[..., :, 1] * p[..., 1:2]) + m[..., :, 2] * p[..., 2:3]


def lidar_coor(calib, depth_cfg, input_size, downsample):
    """Frustum points in the ego/lidar frame, (B, N, D, H, W, 3) float32: undo the image
    augmentation, un-project with the intrinsics, move to the ego frame, apply the BEV
    augmentation (the geometry of view_transformer.py:115-157).  Input synthesis for tests and
    benchmarks; parity of the pooling path is defined on `coor`, whatever produced it."""
    s2e, _e2g, intrin, post_rot, post_tran, bda = calib
    dev = s2e.device
    fr = frustum(depth_cfg, input_size, downsample, device=dev)
    pts = fr[None, None] - post_tran[:, :, None, None, None, :]
    pts = _apply3(torch.linalg.inv(post_rot.double()).float(), pts)
    pts = torch.cat((pts[..., :2] * pts[..., 2:3], pts[..., 2:3]), -1)
    combine = (s2e[:, :, :3, :3].double() @ torch.linalg.inv(intrin.double())).float()
    pts = _apply3(combine, pts) + s2e[:, :, None, None, None, :3, 3]
    return _apply3(bda[:, None], pts).contiguous()


def grid_tensors(grid):
    """(grid_lower_bound, grid_interval, grid_size) as the fp32 tensors the reference builds in
    create_grid_infos (view_transformer.py:80-83)."""
    cfgs = [grid["x"], grid["y"], grid["z"]]
    return (torch.Tensor([c[0] for c in cfgs]), torch.Tensor([c[2] for c in cfgs]),
            torch.Tensor([(c[1] - c[0]) / c[2] for c in cfgs]))
