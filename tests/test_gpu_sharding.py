"""GPU tests of the sharded paths with "virtual ranks" on one device: the per-shard CUDA results must
combine to the single-shard result bit-exactly (the collective itself is an integer sum / max)."""
import numpy as np
import pytest
import torch

from oracle import ref_path

pytestmark = pytest.mark.gpu


def test_sum_of_shard_grids_equals_whole_grid():
    from lidar_ai_recommendation_software_b200 import ops, synth
    from lidar_ai_recommendation_software_b200.sharding import frame_range, sharded_grid_density
    n, world, g = 400_000, 4, 0.5
    pts = synth.crowd_frame(n, seed=7, extent=200.0, extent_y=150.0, people_frac=0.4)   # C.4 venue, scaled down
    d = torch.from_numpy(pts).cuda()
    # single "rank": the product path end to end
    gx, gy, dens = sharded_grid_density(d, g)
    xyz = pts[:, :3].astype(np.float64)
    wx, wy, wd = ref_path.calculate_grid_density(xyz[:, :2], (xyz[:, 0].min(), xyz[:, 0].max()),
                                                 (xyz[:, 1].min(), xyz[:, 1].max()), g)
    assert np.array_equal(gx, wx) and np.array_equal(gy, wy) and np.array_equal(dens, wd)
    # virtual ranks: bbox max-merge, same edges everywhere, integer grid sum
    los, his = [], []
    for r in range(world):
        sl = frame_range(n, r, world)
        bb = ops.bbox(d[sl.start:sl.stop])
        los.append(bb[:2]); his.append(bb[4:6])
    lo = torch.stack(los).min(0).values.cpu().numpy()
    hi = torch.stack(his).max(0).values.cpu().numpy()
    ex, ey = ops.arange_edges(lo[0], hi[0], g), ops.arange_edges(lo[1], hi[1], g)
    total = torch.zeros((len(ex) - 1, len(ey) - 1), dtype=torch.int32, device="cuda")
    for r in range(world):
        sl = frame_range(n, r, world)
        total += ops.hist2d_points_counts(d[sl.start:sl.stop], ex, ey)
    assert np.array_equal(total.cpu().numpy().astype(np.float64) / (g * g), wd)
    # accumulate-into-one-grid form (counts are ADDED by the kernel)
    acc = torch.zeros_like(total)
    for r in range(world):
        sl = frame_range(n, r, world)
        ops.hist2d_points_counts(d[sl.start:sl.stop], ex, ey, out=acc)
    assert torch.equal(acc, total)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_nccl_point_sharded_density_two_gpus(tmp_path):
    import subprocess
    import sys
    script = tmp_path / "w.py"
    script.write_text(
        "import os, sys, numpy as np, torch, torch.distributed as dist\n"
        "sys.path.insert(0, os.environ['LIDAR_ROOT'])\n"
        "from lidar_ai_recommendation_software_b200 import synth\n"
        "from lidar_ai_recommendation_software_b200.sharding import frame_range, sharded_grid_density\n"
        "r = int(os.environ['RANK']); w = int(os.environ['WORLD_SIZE']); torch.cuda.set_device(r)\n"
        "dist.init_process_group('nccl', device_id=torch.device('cuda', r))\n"
        "pts = synth.crowd_frame(300000, seed=7, extent=100.0)\n"
        "sl = frame_range(len(pts), r, w)\n"
        "gx, gy, d = sharded_grid_density(torch.from_numpy(pts[sl.start:sl.stop]).cuda(), 0.5)\n"
        "if r == 0: np.save(os.environ['OUT'], d)\n"
        "dist.destroy_process_group()\n")
    import os
    env = dict(os.environ, LIDAR_ROOT=str(__import__('pathlib').Path(__file__).resolve().parent.parent),
               OUT=str(tmp_path / "d.npy"))
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                    "--master-addr", "127.0.0.1", "--master-port", "29617", str(script)], check=True, env=env,
                   timeout=300)
    from lidar_ai_recommendation_software_b200 import synth
    pts = synth.crowd_frame(300000, seed=7, extent=100.0)
    xyz = pts[:, :3].astype(np.float64)
    _, _, wd = ref_path.calculate_grid_density(xyz[:, :2], (xyz[:, 0].min(), xyz[:, 0].max()),
                                               (xyz[:, 1].min(), xyz[:, 1].max()), 0.5)
    assert np.array_equal(np.load(tmp_path / "d.npy"), wd)
