"""GPU parity tests (first slice): bbox, density histogram, ROI crop, voxel downsample and the fused
frame pipeline — CUDA path through the C ABI vs the CPU oracle, bit-exact for every integer output.
"""
import numpy as np
import pytest
import torch

from oracle import new_ops, np_semantics as nps, ref_path

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from lidar_ai_recommendation_software_b200 import ops as _ops
    return _ops


@pytest.fixture(scope="module")
def synth():
    from lidar_ai_recommendation_software_b200 import synth as _s
    return _s


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ---- K1 ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 31, 1000, 100003])
def test_bbox_exact(ops, synth, n):
    pts = synth.crowd_frame(n, seed=n)
    bb = ops.bbox(dev(pts)).cpu().numpy()
    p64 = pts.astype(np.float64)
    assert np.array_equal(bb[:4], p64.min(0)) and np.array_equal(bb[4:], p64.max(0))
    xyz = p64[:, :3].copy()
    bb = ops.bbox(dev(xyz)).cpu().numpy()
    assert np.array_equal(bb[:3], xyz.min(0)) and np.array_equal(bb[4:7], xyz.max(0))


def test_bbox_empty(ops):
    bb = ops.bbox(torch.empty((0, 4), dtype=torch.float32, device="cuda")).cpu().numpy()
    assert np.all(np.isposinf(bb[:4])) and np.all(np.isneginf(bb[4:]))


def test_moments_close_to_numpy(ops, synth):
    xyz = synth.crowd_frame(50000, seed=2)[:, :3].astype(np.float64)
    m = ops.moments(dev(xyz)).cpu().numpy()
    mean = m[:3] / len(xyz)
    assert np.allclose(mean, xyz.mean(0), rtol=1e-12, atol=1e-13)
    m2 = ops.moments(dev(xyz), center=mean).cpu().numpy()
    assert np.allclose(np.sqrt(m2[3:] / len(xyz)), xyz.std(0), rtol=1e-12)


# ---- K6 ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("g", [1.0, 0.5, 0.3])
def test_grid_density_counts_bit_exact(ops, synth, mode, g):
    xyz = synth.crowd_frame(200000, seed=5)[:, :3].astype(np.float64)
    xr = (xyz[:, 0].min(), xyz[:, 0].max())
    yr = (xyz[:, 1].min(), xyz[:, 1].max())
    want, ex, ey = ref_path.grid_density_counts(xyz[:, :2], xr, yr, g)
    if mode == 2 and want.size * 4 > 200 * 1024:
        pytest.skip("grid does not fit a CTA's shared memory: SHARED mode is refused by design")
    d = dev(xyz)
    got = ops.hist2d_counts(d[:, 0], d[:, 1], ops.arange_edges(*xr, g), ops.arange_edges(*yr, g), mode=mode)
    assert np.array_equal(ops.arange_edges(*xr, g), ex)
    assert np.array_equal(got.cpu().numpy(), want)
    assert int(got.sum()) == len(xyz)


def test_hist2d_edge_cases(ops):
    ex = np.arange(-3.0, 3.0 + 0.5, 0.5)
    ey = np.linspace(-2.0, 2.0, 9)
    rng = np.random.default_rng(0)
    u = rng.uniform(-4, 4, 5000)
    v = rng.uniform(-3, 3, 5000)
    u[:20] = np.repeat(ex[[0, 3, -1, -2]], 5)
    v[:20] = np.tile(ey[[0, 2, -1, 4, -2]], 4)
    u[20] = np.nan
    want = nps.histogram2d_counts(u, v, ex, ey)
    for mode in (1, 2):
        got = ops.hist2d_counts(dev(u), dev(v), ex, ey, mode=mode).cpu().numpy()
        assert np.array_equal(got, want)
    # non-uniform edges take the binary-search path
    exn = np.array([-3.0, -2.9, -1.0, 0.0, 0.001, 2.5, 3.0])
    want = nps.histogram2d_counts(u[21:], v[21:], exn, ey)
    got = ops.hist2d_counts(dev(u[21:]), dev(v[21:]), exn, ey).cpu().numpy()
    assert np.array_equal(got, want)
    # empty input
    got = ops.hist2d_counts(dev(u[:0]), dev(v[:0]), ex, ey)
    assert int(got.sum()) == 0


def test_hist2d_big_scan_equals_numpy(ops, synth):
    """5 M points of the venue scan on a grid too large for shared memory (the GLOBAL path of config 5):
    counts must equal numpy's, accumulate into a pre-filled output, and sum to n."""
    pts = synth.venue_scan_shard(5_000_000, 7, 0, 1)
    xy = pts[:, :2].astype(np.float64)
    ex = ops.arange_edges(xy[:, 0].min(), xy[:, 0].max(), 0.5)
    ey = ops.arange_edges(xy[:, 1].min(), xy[:, 1].max(), 0.5)
    assert (len(ex) - 1) * (len(ey) - 1) * 4 > 227 * 1024
    want = nps.histogram2d_counts(xy[:, 0], xy[:, 1], ex, ey)
    d = dev(pts)
    got = ops.hist2d_points_counts(d, ex, ey)
    assert np.array_equal(got.cpu().numpy(), want) and int(got.sum()) == len(pts)
    again = ops.hist2d_points_counts(d, ex, ey, out=got)       # accumulates on top
    assert np.array_equal(again.cpu().numpy(), 2 * want)


def test_heatmap_linspace_edges_float4(ops, synth):
    pts = synth.crowd_frame(300000, seed=9)
    xyz = pts[:, :3].astype(np.float64)
    pd = {"points": xyz, "dimensions": {"x_range": (xyz[:, 0].min(), xyz[:, 0].max()),
                                        "y_range": (xyz[:, 1].min(), xyz[:, 1].max()),
                                        "z_range": (xyz[:, 2].min(), xyz[:, 2].max())}}
    want, ex, ey = ref_path.heatmap_counts(pd, bins=100)
    got = ops.hist2d_points_counts(dev(pts), ops.linspace_edges(*pd["dimensions"]["x_range"], 100),
                                   ops.linspace_edges(*pd["dimensions"]["y_range"], 100))
    assert np.array_equal(got.cpu().numpy(), want)
    assert int(got.sum()) == len(pts)  # max point lands in the closed last bin


# ---- a5 ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [0, 1, 2047, 2048, 2049, 150001])
def test_roi_crop_bit_exact(ops, synth, n):
    pts = synth.crowd_frame(max(n, 1), seed=3, extent=20.0)[:n]
    lo, hi = (-5.0, -7.5, 0.05), (12.25, 9.0, 1.5)
    want, wmask = new_ops.roi_crop(pts, lo, hi)
    got, mask = ops.roi_crop(dev(pts) if n else torch.empty((0, 4), dtype=torch.float32, device="cuda"), lo, hi)
    assert np.array_equal(mask.cpu().numpy().astype(bool), wmask)
    assert np.array_equal(got.cpu().numpy(), want)
    xyz = pts[:, :3].astype(np.float64)
    want, wmask = new_ops.roi_crop(xyz, lo, hi)
    got, mask = ops.roi_crop(dev(xyz) if n else torch.empty((0, 3), dtype=torch.float64, device="cuda"), lo, hi)
    assert np.array_equal(got.cpu().numpy(), want) and np.array_equal(mask.cpu().numpy().astype(bool), wmask)


# ---- K5 ------------------------------------------------------------------------------------------
def check_voxels(res, want, n):
    assert tuple(res.dims) == tuple(want["dims"])
    assert np.array_equal(res.voxel_key.cpu().numpy(), want["voxel_key"])
    assert np.array_equal(res.inverse.cpu().numpy(), want["inverse"])
    assert res.n_voxels == len(want["unique_keys"])
    assert np.array_equal(res.unique_keys.cpu().numpy(), want["unique_keys"])
    assert np.array_equal(res.counts.cpu().numpy(), want["counts"])
    got_c = res.centroids.cpu().numpy()
    assert np.allclose(got_c, want["centroids"], rtol=1e-6, atol=1e-7)
    # in practice the fixed-point mean rounds to the same fp32 as the fp64 mean
    assert np.mean(got_c == want["centroids"]) > 0.9999


@pytest.mark.parametrize("n,voxel,extent", [(1, 0.05, 5.0), (1000, 0.05, 5.0), (100000, 0.05, 50.0),
                                            (100000, 0.5, 50.0), (250000, 0.2, 10.0)])
def test_voxel_downsample_bit_exact(ops, synth, n, voxel, extent):
    pts = synth.crowd_frame(n, seed=11, extent=extent)
    want = new_ops.voxel_downsample(pts, voxel)
    res = ops.voxel_downsample(dev(pts), voxel)
    check_voxels(res, want, n)


def test_voxel_downsample_with_origin_and_duplicates(ops, synth):
    pts = synth.crowd_frame(5000, seed=1, extent=3.0)
    pts = np.concatenate([pts, pts[:1000]])  # exact duplicates share a voxel
    org = (-4.0, -4.0, -1.0)
    want = new_ops.voxel_downsample(pts, 0.1, origin=org)
    res = ops.voxel_downsample(dev(pts), 0.1, origin=org)
    check_voxels(res, want, len(pts))


def test_frame_pipeline_fused_density_and_reuse(ops, synth):
    """bbox -> voxelise + calculate_grid_density counts with no host round trip; the pipeline object
    is reused for several frames (exercises the all-zero workspace invariant)."""
    pipe = ops.FramePipeline(max_points=120000, voxel_size=0.05, grid_size=0.5, max_key_space=1 << 28,
                             max_nx=256, max_ny=256)
    for seed, n in ((0, 100000), (1, 120000), (2, 777), (0, 100000)):
        pts = synth.crowd_frame(n, seed=seed, extent=50.0)
        pipe.enqueue(dev(pts))
        res = pipe.result()
        want = new_ops.voxel_downsample(pts, 0.05)
        check_voxels(res, want, n)
        xyz = pts[:, :3].astype(np.float64)
        xr = (xyz[:, 0].min(), xyz[:, 0].max())
        yr = (xyz[:, 1].min(), xyz[:, 1].max())
        wc, ex, ey = ref_path.grid_density_counts(xyz[:, :2], xr, yr, 0.5)
        assert res.grid_counts.shape == wc.shape
        assert np.array_equal(res.grid_counts.cpu().numpy(), wc)
        gx, gy = res.grid_edges()
        assert np.array_equal(gx, ex) and np.array_equal(gy, ey)


def test_frame_pipeline_determinism(ops, synth):
    pts = dev(synth.crowd_frame(200000, seed=4, extent=20.0))
    pipe = ops.FramePipeline(max_points=200000, voxel_size=0.05, grid_size=0.5, max_nx=128, max_ny=128)
    outs = []
    for _ in range(3):
        pipe.enqueue(pts)
        r = pipe.result()
        outs.append((r.centroids.clone(), r.counts.clone(), r.inverse.clone(), r.grid_counts.clone()))
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            assert torch.equal(a, b)


def test_frame_pipeline_capacity_error(ops, synth):
    from lidar_ai_recommendation_software_b200._capi import LidarError
    pts = dev(synth.crowd_frame(10000, seed=4, extent=50.0))
    pipe = ops.FramePipeline(max_points=10000, voxel_size=0.05, max_key_space=1 << 20)
    pipe.enqueue(pts)
    with pytest.raises(LidarError):
        pipe.result()
    # the pipeline recovers: a frame that fits still works
    small = synth.crowd_frame(10000, seed=4, extent=1.0)
    pipe.enqueue(dev(small))
    check_voxels(pipe.result(), new_ops.voxel_downsample(small, 0.05), 10000)
    with pytest.raises(LidarError):
        pipe.enqueue(dev(synth.crowd_frame(10001, seed=1)))


def test_full_size_properties_1m(ops, synth):
    """BASELINE config 2 at full size (1 M points): size-independent properties."""
    pts = synth.crowd_frame(1_000_000, seed=0, extent=50.0)
    d = dev(pts)
    pipe = ops.FramePipeline(max_points=1_000_000, voxel_size=0.05, grid_size=0.5, max_nx=256, max_ny=256)
    pipe.enqueue(d)
    r = pipe.result()
    assert int(r.counts.sum()) == len(pts) and int(r.grid_counts.sum()) == len(pts)
    uk = r.unique_keys.cpu().numpy()
    assert np.all(np.diff(uk) > 0)
    assert torch.equal(r.unique_keys[r.inverse.long()], r.voxel_key)
    # idempotence: downsampling the centroids with the same origin keeps every voxel (count 1 each)
    pipe2 = ops.FramePipeline(max_points=r.n_voxels, voxel_size=0.05)
    pipe2.enqueue(r.centroids.contiguous(), origin=r.origin)
    r2 = pipe2.result()
    assert r2.n_voxels >= int(0.999 * r.n_voxels)
    # oracle comparison on the integer outputs at full size (numpy unique: a few seconds)
    want = new_ops.voxel_downsample(pts, 0.05)
    assert np.array_equal(r.inverse.cpu().numpy(), want["inverse"])
    assert np.array_equal(r.counts.cpu().numpy(), want["counts"])
