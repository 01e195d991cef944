"""Host-side logic of the N > 1 path on CPU: world_size 2, gloo.  Each rank pools its shard of
the batch (with the CPU oracle standing in for the GPU kernels -- this test is about the
partitioning and the verification gather, not the arithmetic) and the gathered result must
equal the unsharded one."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_samples, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import oracle
    from rcbevdet_b200 import rig, shard
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo_s, hi_s = shard.shard_range(n_samples, rank, world)
        grid = rig.R50_GRID
        lo, iv, sz = (t.numpy() for t in rig.grid_tensors(grid))
        coor = rig.lidar_coor(rig.camera_rig(n_samples, input_size=(64, 176), aug_seed=3), [1.0, 60.0, 4.0],
                              (64, 176), 16)
        _, N, D, H, W, _ = coor.shape
        depth, feat = rig.pooling_inputs(n_samples, N, D, H, W, 8, seed=4)

        def pool(c, d, f):
            b = c.shape[0]
            rb, rd, rf, st, ln = oracle.voxel_pooling_prepare_v2(c.numpy(), lo, iv, sz)
            rows = f.permute(0, 1, 3, 4, 2).contiguous().numpy()
            return torch.from_numpy(oracle.to_bczyx(oracle.bev_pool_v2_forward(
                d.numpy(), rows, rd, rf, rb, (b, 1, 128, 128, 8), st, ln)))

        local = pool(coor[lo_s:hi_s], depth[lo_s:hi_s], feat[lo_s:hi_s])
        full = shard.gather_samples(local, n_samples)
        times = shard.max_over_ranks([float(rank + 1), 5.0 - rank])
        if rank == 0:
            want = pool(coor, depth, feat)
            q.put((bool(torch.equal(full, want)), times, (lo_s, hi_s)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_samples", [3, 4])
def test_two_rank_sharding_and_gather(n_samples):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_samples, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    same, times, rng = q.get(timeout=10)
    assert same, "gathered shards differ from the unsharded result"
    assert times == [2.0, 5.0]
    assert rng == (0, n_samples - n_samples // 2)


def test_shard_ranges_cover_everything():
    from rcbevdet_b200 import shard
    for n in (0, 1, 7, 8, 9, 64):
        for world in (1, 2, 3, 8):
            parts = [shard.shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1
