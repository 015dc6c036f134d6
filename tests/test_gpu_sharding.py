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


def test_scan_density_context_single_rank_matches_reference_and_three_enqueue_form():
    """world = 1: the fused cooperative kernel (device-side edges, host-mapped descriptor) against the reference's
    calculate_grid_density restatement, for both point layouts, an empty cloud and a capacity overflow."""
    from lidar_ai_recommendation_software_b200 import _capi, synth
    from lidar_ai_recommendation_software_b200.sharding import ScanDensity
    ctx = ScanDensity(torch.device("cuda", 0), cap_cells=1 << 18, max_nx=1024, max_ny=1024)
    for n, ext, g in [(250_000, 60.0, 0.5), (1000, 5.0, 1.0), (1, 1.0, 0.25), (40_000, 30.0, 0.3)]:
        pts = synth.crowd_frame(n, seed=n % 17, extent=ext, extent_y=ext * 0.6)
        xyz = pts[:, :3].astype(np.float64)
        wx, wy, wd = ref_path.calculate_grid_density(xyz[:, :2], (xyz[:, 0].min(), xyz[:, 0].max()),
                                                     (xyz[:, 1].min(), xyz[:, 1].max()), g)
        for dev_pts in (torch.from_numpy(pts).cuda(), torch.from_numpy(np.ascontiguousarray(xyz)).cuda()):
            gx, gy, dens = ctx(dev_pts, g)
            assert np.array_equal(gx, wx) and np.array_equal(gy, wy) and np.array_equal(dens, wd)
    assert ctx(torch.empty((0, 4), dtype=torch.float32, device="cuda"), 0.5) == (None, None, None)
    big = torch.from_numpy(synth.crowd_frame(1000, seed=1, extent=400.0)).cuda()
    with pytest.raises(_capi.LidarError):
        ctx(big, 0.25)                       # 3 200 x 3 200 cells > cap_cells
    gx, gy, dens = ctx(torch.from_numpy(synth.crowd_frame(5000, seed=2, extent=10.0)).cuda(), 0.5)   # still usable
    assert int(round(dens.sum() * 0.25)) == 5000
    ctx.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_point_sharded_density_two_gpus_fused_and_nccl(tmp_path):
    """Real multi-GPU run (torchrun, one rank per GPU): the fused NVLink kernel and the lidar_nccl_* fallback must
    both reproduce the single-GPU reference grid bit for bit, several calls in a row, with an empty shard too."""
    import os
    import subprocess
    import sys
    world = min(torch.cuda.device_count(), 8)
    script = tmp_path / "w.py"
    script.write_text(
        "import os, sys, json, numpy as np, torch, torch.distributed as dist\n"
        "sys.path.insert(0, os.environ['LIDAR_ROOT'])\n"
        "from lidar_ai_recommendation_software_b200 import synth\n"
        "from lidar_ai_recommendation_software_b200.sharding import frame_range, ScanDensity\n"
        "r = int(os.environ['RANK']); w = int(os.environ['WORLD_SIZE']); torch.cuda.set_device(r)\n"
        "dev = torch.device('cuda', r)\n"
        "dist.init_process_group('nccl', device_id=dev)\n"
        "info = {}\n"
        "for backend in ('fused', 'nccl'):\n"
        "    ctx = ScanDensity(dev, backend=backend, cap_cells=1 << 18, max_nx=1024, max_ny=1024)\n"
        "    info[backend] = {'backend': ctx.backend, 'multicast': bool(getattr(ctx, 'multicast', False))}\n"
        "    for k, (n, ext) in enumerate([(300000, 100.0), (50000, 20.0), (300000, 100.0), (3, 2.0)]):\n"
        "        pts = synth.crowd_frame(n, seed=7 + k, extent=ext)\n"
        "        sl = frame_range(len(pts), r, w)\n"
        "        shard = pts[sl.start:sl.stop] if n != 3 else (pts if r == 0 else pts[:0])\n"
        "        gx, gy, d = ctx(torch.from_numpy(shard).to(dev), 0.5)\n"
        "        if r == 0: np.savez(os.path.join(os.environ['OUT'], f'{backend}_{k}.npz'), gx=gx, gy=gy, d=d)\n"
        "    ctx.close()\n"
        "if r == 0: json.dump(info, open(os.path.join(os.environ['OUT'], 'info.json'), 'w'))\n"
        "dist.destroy_process_group()\n")
    env = dict(os.environ, LIDAR_ROOT=str(__import__('pathlib').Path(__file__).resolve().parent.parent), OUT=str(tmp_path))
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                    "--master-addr", "127.0.0.1", "--master-port", "29617", str(script)], check=True, env=env,
                   timeout=600)
    import json
    from lidar_ai_recommendation_software_b200 import synth
    info = json.load(open(tmp_path / "info.json"))
    print("sharded density back ends:", info)
    assert info["fused"]["backend"] == "fused" and info["nccl"]["backend"] == "nccl"
    for backend in ("fused", "nccl"):
        for k, (n, ext) in enumerate([(300000, 100.0), (50000, 20.0), (300000, 100.0), (3, 2.0)]):
            pts = synth.crowd_frame(n, seed=7 + k, extent=ext)
            xyz = pts[:, :3].astype(np.float64)
            wx, wy, wd = ref_path.calculate_grid_density(xyz[:, :2], (xyz[:, 0].min(), xyz[:, 0].max()),
                                                         (xyz[:, 1].min(), xyz[:, 1].max()), 0.5)
            got = np.load(tmp_path / f"{backend}_{k}.npz")
            assert np.array_equal(got["gx"], wx) and np.array_equal(got["gy"], wy), (backend, k)
            assert np.array_equal(got["d"], wd), (backend, k)
