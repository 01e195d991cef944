"""PyTorch restatement of the reference path, for TIMING the reference's own formulation on any
torch device (the GPU box has no /root/reference to import).  TEST INFRASTRUCTURE ONLY: used by
tests/ (pinned against the reference-generated goldens and the numpy oracle) and by bench.py's
`cpu_baseline` / `reference_gpu` legs; never imported by the product package.

  prepare(coor, lower, interval, size)  -- view_transformer.py:207-265: the same chain of torch
      ops (mask compaction, fp32 rank, argsort, run boundaries), so its cost on a device is the
      reference's cost.
  pool(depth, feat, ranks..., shape)    -- the reference has NO CPU pool kernel (its extension
      passes data_ptr straight to a CUDA launch); this is the index_add_ restatement of
      bev_pool_cuda.cu:21-48 + the permute of bev_pool.py:91 that SURVEY.md section 8(c)
      validated against the reference's known-answer test.  Autograd of it is the backward.
"""
from __future__ import annotations

import torch


def prepare(coor, lower, interval, size):
    """-> (ranks_bev, ranks_depth, ranks_feat, interval_starts, interval_lengths) int32, or 5 x None."""
    B, N, D, H, W, _ = coor.shape
    n = B * N * D * H * W
    dev = coor.device
    lower, interval, size = (torch.as_tensor(t, dtype=torch.float32) for t in (lower, interval, size))
    ranks_depth = torch.arange(n, dtype=torch.int32, device=dev)                                   # :223-224
    ranks_feat = torch.arange(n // D, dtype=torch.int32, device=dev).reshape(B, N, 1, H, W)        # :225-228
    ranks_feat = ranks_feat.expand(B, N, D, H, W).flatten()
    vox = ((coor - lower.to(coor)) / interval.to(coor)).long().view(n, 3)                          # :230-232
    batch = torch.arange(B, device=dev).reshape(B, 1).expand(B, n // B).reshape(n, 1).to(vox)      # :233-234
    vox = torch.cat((vox, batch), 1)                                                               # :235
    keep = (vox[:, 0] >= 0) & (vox[:, 0] < size[0]) & (vox[:, 1] >= 0) & (vox[:, 1] < size[1]) & \
           (vox[:, 2] >= 0) & (vox[:, 2] < size[2])                                                # :238-240
    vox, ranks_depth, ranks_feat = vox[keep], ranks_depth[keep], ranks_feat[keep]                  # :243-244
    ranks_bev = vox[:, 3] * (size[2] * size[1] * size[0])                                          # :246-247 (fp32)
    ranks_bev = ranks_bev + vox[:, 2] * (size[1] * size[0])                                        # :248
    ranks_bev = ranks_bev + (vox[:, 1] * size[0] + vox[:, 0])                                      # :249
    order = ranks_bev.argsort()                                                                    # :250
    ranks_bev, ranks_depth, ranks_feat = ranks_bev[order], ranks_depth[order], ranks_feat[order]   # :251-252
    first = torch.ones(ranks_bev.shape[0], device=dev, dtype=torch.bool)                           # :254-256
    first[1:] = ranks_bev[1:] != ranks_bev[:-1]
    starts = torch.where(first)[0].int()                                                           # :257
    if len(starts) == 0:                                                                           # :258-259
        return None, None, None, None, None
    lengths = torch.zeros_like(starts)                                                             # :260-262
    lengths[:-1] = starts[1:] - starts[:-1]
    lengths[-1] = ranks_bev.shape[0] - starts[-1]
    return (ranks_bev.int().contiguous(), ranks_depth.int().contiguous(), ranks_feat.int().contiguous(),
            starts.int().contiguous(), lengths.int().contiguous())


def pool(depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape):
    """depth (B,N,D,H,W), feat (B,N,H,W,C) any strides -> (B,C,Z,Y,X) contiguous (bev_pool.py:86-92)."""
    B, Z, Y, X, C = (int(s) for s in bev_feat_shape)
    rows = feat.contiguous().view(-1, C)                                                           # bev_pool.py:21
    w = depth.contiguous().view(-1)[ranks_depth.long()]
    out = rows.new_zeros((B * Z * Y * X, C))                                                       # bev_pool.py:27
    out = out.index_add(0, ranks_bev.long(), w[:, None] * rows[ranks_feat.long()])                 # bev_pool_cuda.cu:21-48
    return out.view(B, Z, Y, X, C).permute(0, 4, 1, 2, 3).contiguous()                             # bev_pool.py:91


def gen_grid(input, sensor2keyegos, bda, bda_adj, grid_interval, grid_lower_bound):
    """BEVDepth4D.gen_grid (bevdet_rc.py:585-652) restated op for op: normalised sampling grid
    (n, h, w, 2) that maps the key frame's BEV pixels into the adjacent frame's BEV."""
    n, c, h, w = input.shape
    xs = torch.linspace(0, w - 1, w, dtype=input.dtype, device=input.device).view(1, w).expand(h, w)   # :590-592
    ys = torch.linspace(0, h - 1, h, dtype=input.dtype, device=input.device).view(h, 1).expand(h, w)   # :593-595
    grid = torch.stack((xs, ys, torch.ones_like(xs)), -1)                                              # :596
    grid = grid.view(1, h, w, 3).expand(n, h, w, 3).view(n, h, w, 3, 1)                                # :600
    c02l0 = sensor2keyegos[0][:, 0:1, :, :]                                                            # :604
    c12l0 = sensor2keyegos[1][:, 0:1, :, :]                                                            # :607
    bda_ = torch.zeros((n, 1, 4, 4), dtype=grid.dtype).to(grid)                                        # :610-612
    bda_[:, :, :3, :3] = bda.unsqueeze(1)
    bda_[:, :, 3, 3] = 1
    c02l0 = bda_.matmul(c02l0)                                                                         # :613
    if bda_adj is not None:                                                                            # :614-617
        bda_ = torch.zeros((n, 1, 4, 4), dtype=grid.dtype).to(grid)
        bda_[:, :, :3, :3] = bda_adj.unsqueeze(1)
        bda_[:, :, 3, 3] = 1
    c12l0 = bda_.matmul(c12l0)                                                                         # :618
    l02l1 = c02l0.matmul(torch.inverse(c12l0))[:, 0, :, :].view(n, 1, 1, 4, 4)                         # :622-623
    keep = [True, True, False, True]
    l02l1 = l02l1[:, :, :, keep, :][:, :, :, :, keep]                                                  # :631-633
    feat2bev = torch.zeros((3, 3), dtype=grid.dtype).to(grid)                                          # :635-641
    feat2bev[0, 0] = float(grid_interval[0])
    feat2bev[1, 1] = float(grid_interval[1])
    feat2bev[0, 2] = float(grid_lower_bound[0])
    feat2bev[1, 2] = float(grid_lower_bound[1])
    feat2bev[2, 2] = 1
    feat2bev = feat2bev.view(1, 3, 3)
    tf = torch.inverse(feat2bev).matmul(l02l1).matmul(feat2bev)                                        # :642
    grid = tf.matmul(grid)                                                                             # :645
    norm = torch.tensor([w - 1.0, h - 1.0], dtype=input.dtype, device=input.device)                    # :646-648
    return grid[:, :, :, :2, 0] / norm.view(1, 1, 1, 2) * 2.0 - 1.0                                    # :649-650


def shift_feature(input, sensor2keyegos, bda, bda_adj, grid_interval, grid_lower_bound):
    """bevdet_rc.py:654-657."""
    grid = gen_grid(input, sensor2keyegos, bda, bda_adj, grid_interval, grid_lower_bound)
    return torch.nn.functional.grid_sample(input, grid.to(input.dtype), align_corners=True)
