"""The 64-bit-key sort path of voxel downsample (`lidar_voxel_downsample_sorted`) against the int64 oracle
(oracle/new_ops.py) and against the occupancy-bitmap frame kernel where both apply."""
import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from oracle import new_ops

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from lidar_ai_recommendation_software_b200 import ops as _ops
    return _ops


def check(ops, pts, voxel, origin=None, roi=None):
    d = torch.from_numpy(np.ascontiguousarray(pts, dtype=np.float32)).cuda()
    r = ops.voxel_downsample_sorted(d, voxel, origin=origin, roi=roi)
    keep = np.ones(len(pts), dtype=bool)
    if roi is not None:
        _, keep = new_ops.roi_crop(pts.astype(np.float32), roi[0], roi[1])
    want = new_ops.voxel_downsample(pts[keep], voxel, origin=origin)
    assert r.n_kept == int(keep.sum()) and r.n_voxels == len(want["unique_keys"])
    key, inv = r.voxel_key.cpu().numpy(), r.inverse.cpu().numpy()
    assert key.dtype == np.int64 and np.array_equal(key[keep], want["voxel_key"]) and np.all(key[~keep] == -1)
    assert np.array_equal(inv[keep], want["inverse"]) and np.all(inv[~keep] == -1)
    assert np.array_equal(r.unique_keys.cpu().numpy(), want["unique_keys"])
    assert np.array_equal(r.counts.cpu().numpy(), want["counts"])
    assert np.allclose(r.centroids.cpu().numpy(), want["centroids"], rtol=1e-6, atol=1e-6 * max(1.0, float(np.abs(pts[:, :3]).max())))
    if len(want["unique_keys"]):
        assert r.dims == want["dims"]
    return r


def test_kilometre_venue_needs_more_than_31_bits(ops):
    """1 km x 1 km x 30 m at 0.05 m = 2.4e11 cells; the bitmap kernels stop at 2^31."""
    rng = np.random.default_rng(0)
    n = 400_000
    pts = np.column_stack([rng.uniform(0, 1000, n), rng.uniform(0, 1000, n), rng.uniform(0, 30, n), rng.uniform(0, 1, n)]).astype(np.float32)
    pts[: n // 4] = pts[n // 4: n // 2] + np.float32(0.004)              # plenty of multi-member voxels
    r = check(ops, pts, 0.05)
    assert r.desc.key_space > (1 << 31) and r.desc.passes == 5
    auto = ops.voxel_downsample(torch.from_numpy(pts).cuda(), 0.05)      # the one-shot op picks this path by itself
    assert torch.equal(auto.unique_keys, r.unique_keys) and torch.equal(auto.inverse, r.inverse)


def test_one_far_outlier(ops, ):
    from lidar_ai_recommendation_software_b200 import synth
    pts = synth.crowd_frame(50_000, seed=3, extent=20.0)
    pts[777, :3] = (9.0e4, -7.5e4, 3.0e3)
    r = check(ops, pts, 0.05)
    assert r.desc.key_space > (1 << 31)


def test_equals_the_bitmap_kernel_where_both_apply(ops):
    from lidar_ai_recommendation_software_b200 import synth
    for n, ext in ((1, 1.0), (33, 2.0), (20_000, 10.0), (300_000, 50.0)):
        pts = synth.crowd_frame(n, seed=n % 7, extent=ext)
        d = torch.from_numpy(pts).cuda()
        r = check(ops, pts, 0.05)
        pipe = ops.FramePipeline(max_points=n, voxel_size=0.05, max_key_space=1 << 28)
        pipe.enqueue(d)
        f = pipe.result()
        assert f.n_voxels == r.n_voxels
        assert torch.equal(f.inverse, r.inverse) and torch.equal(f.counts, r.counts)
        assert torch.equal(f.unique_keys.long(), r.unique_keys) and torch.equal(f.voxel_key.long(), r.voxel_key)
        assert torch.equal(f.centroids, r.centroids)                      # same exact sums, same rounding
    empty = ops.voxel_downsample_sorted(torch.empty((0, 4), dtype=torch.float32, device="cuda"), 0.05)
    assert empty.n_voxels == 0 and empty.inverse.numel() == 0


def test_fused_roi_crop_and_explicit_origin(ops):
    from lidar_ai_recommendation_software_b200 import synth
    pts = synth.crowd_frame(120_000, seed=5, extent=40.0)
    roi = ((-10.0, -12.5, 0.0), (25.0, 30.0, 1.5))
    r = check(ops, pts, 0.1, roi=roi)
    assert 0 < r.n_kept < len(pts)
    check(ops, pts, 0.1, origin=(-41.0, -41.0, -1.0), roi=roi)
    none = check(ops, pts, 0.1, roi=((500.0, 500.0, 0.0), (600.0, 600.0, 1.0)))      # nothing inside the box
    assert none.n_kept == 0 and none.n_voxels == 0
    dup = np.tile(pts[:1], (5000, 1))
    assert check(ops, dup, 0.05).n_voxels == 1
    from lidar_ai_recommendation_software_b200 import _capi
    with pytest.raises(_capi.LidarError):
        ops.voxel_downsample_sorted(torch.from_numpy(pts).cuda(), 0.1, origin=(0.0, 0.0, 0.0))   # points below the origin


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 3000), st.integers(0, 2**31 - 1), st.sampled_from([0.05, 0.1, 0.37, 1.0]),
       st.sampled_from([0.0, 1.0e3, 3.0e4, 1.0e6]))
def test_property_random_clouds_with_an_outlier(n, seed, voxel, far):
    from lidar_ai_recommendation_software_b200 import ops
    rng = np.random.default_rng(seed)
    pts = np.column_stack([rng.normal(0, 3, n), rng.normal(0, 3, n), rng.uniform(0, 2, n), rng.uniform(0, 1, n)]).astype(np.float32)
    pts[rng.integers(0, n), :2] += np.float32(far)                    # x and y: up to 2e7 x 2e7 x 40 cells, still < 2^63
    if n > 4:
        pts[: n // 3] = pts[n // 3: 2 * (n // 3)]                        # exact duplicates
    check(ops, pts, voxel)
