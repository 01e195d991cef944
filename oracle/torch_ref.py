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
