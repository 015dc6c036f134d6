"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: frame partitioning and the point-sharded
density (bbox MAX all-reduce + integer grid SUM all-reduce).  The per-rank kernels are replaced by the
CPU oracle here — in the product they are the CUDA ops (tests/test_gpu_sharding.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_frame_ranges_partition_everything():
    from lidar_ai_recommendation_software_b200.sharding import frame_range, frame_range_with_halo
    for n in (0, 1, 7, 300, 301):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                seen += list(frame_range(n, r, world))
            assert seen == list(range(n))
            sizes = [len(frame_range(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    r, halo = frame_range_with_halo(300, 2, 8)
    assert halo == r.start - 1 and frame_range_with_halo(300, 0, 8)[1] is None


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, grid, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lidar_ai_recommendation_software_b200 import synth
        from lidar_ai_recommendation_software_b200.sharding import frame_range, gather_results, sharded_grid_density
        from oracle import np_semantics as nps
        pts = synth.crowd_frame(n, seed=7, extent=40.0, extent_y=25.0)
        sl = frame_range(n, rank, world)
        shard = torch.from_numpy(pts[sl.start:sl.stop] if rank != 1 or n > 10 else pts[:0])

        def bbox(p):
            if p.shape[0] == 0:
                return torch.full((2,), float("inf"), dtype=torch.float64), torch.full((2,), float("-inf"), dtype=torch.float64)
            a = p[:, :2].double()
            return a.min(0).values, a.max(0).values

        def hist(p, ex, ey):
            a = p.numpy().astype(np.float64)
            return torch.from_numpy(nps.histogram2d_counts(a[:, 0], a[:, 1], ex, ey).astype(np.int32))

        gx, gy, dens = sharded_grid_density(shard, grid, local_bbox=bbox, local_hist=hist)
        res = gather_results({rank: int(round(dens.sum() * grid * grid))})
        if rank == 0:
            q.put((gx, gy, dens, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [5000, 3])
def test_point_sharded_density_matches_single_rank(n):
    from lidar_ai_recommendation_software_b200 import synth
    from oracle import ref_path
    world, grid = 2, 0.5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, grid, q)) for r in range(world)]
    for p in procs:
        p.start()
    gx, gy, dens, res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pts = synth.crowd_frame(n, seed=7, extent=40.0, extent_y=25.0)
    if n <= 10:
        pts = pts[: (n + 1) // 2]        # rank 1 contributed an empty shard in that case
    xyz = pts[:, :3].astype(np.float64)
    wx, wy, wd = ref_path.calculate_grid_density(xyz[:, :2], (xyz[:, 0].min(), xyz[:, 0].max()),
                                                 (xyz[:, 1].min(), xyz[:, 1].max()), grid)
    assert np.array_equal(gx, wx) and np.array_equal(gy, wy) and np.array_equal(dens, wd)
    assert sorted(res) == [0, 1] and res[0] == len(pts)
