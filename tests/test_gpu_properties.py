"""Property tests (hypothesis) of the CUDA path against numpy / scikit-learn / the oracle on generated inputs:
ragged sizes, duplicated values, values exactly on edges and thresholds, signed zeros, single points.
SURVEY.md §4: histogram sum = #in-range points, voxel inverse map round-trips, FPS indices unique, ball-query
indices within the radius, DBSCAN labels identical to scikit-learn."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

pytestmark = pytest.mark.gpu

CFG = dict(deadline=None, max_examples=80, derandomize=True,
           suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])


def _lattice_cloud(draw, n_max, dims=3, span=4.0):
    """Points on a coarse lattice plus jitter: plenty of exact duplicates and exact edge hits."""
    n = draw(st.integers(1, n_max))
    seed = draw(st.integers(0, 2**31 - 1))
    mode = draw(st.sampled_from(["lattice", "uniform", "mixed", "line"]))
    rng = np.random.default_rng(seed)
    if mode == "lattice":
        p = rng.integers(-8, 9, size=(n, dims)) * 0.25
    elif mode == "uniform":
        p = rng.uniform(-span, span, size=(n, dims))
    elif mode == "line":
        p = np.zeros((n, dims))
        p[:, 0] = rng.integers(-40, 41, size=n) * 0.125
    else:
        p = np.where(rng.random((n, dims)) < 0.5, rng.integers(-8, 9, size=(n, dims)) * 0.25,
                     rng.uniform(-span, span, size=(n, dims)))
    p = p.astype(np.float64)
    p[rng.random(n) < 0.05] *= -0.0 + 1.0          # keep values, exercise the code path with mixed signs
    return np.ascontiguousarray(p)


@st.composite
def clouds3(draw, n_max=600):
    return _lattice_cloud(draw, n_max, 3)


@settings(**CFG)
@given(clouds3(), st.integers(1, 9), st.integers(1, 9), st.sampled_from([1, 2]))
def test_hist2d_is_numpy_histogram2d(p, nx, ny, mode):
    from lidar_ai_recommendation_software_b200 import ops
    ex = np.linspace(-2.0, 2.0, nx + 1)
    ey = np.linspace(-1.5, 2.0, ny + 1)
    d = torch.from_numpy(p).cuda()
    got = ops.hist2d_counts(d[:, 0], d[:, 1], ex, ey, mode=mode).cpu().numpy()
    want, _, _ = np.histogram2d(p[:, 0], p[:, 1], bins=[ex, ey])
    np.testing.assert_array_equal(got, want.astype(np.int64))
    inside = ((p[:, 0] >= ex[0]) & (p[:, 0] <= ex[-1]) & (p[:, 1] >= ey[0]) & (p[:, 1] <= ey[-1])).sum()
    assert got.sum() == inside


@settings(**CFG)
@given(clouds3(), st.sampled_from([0.05, 0.25, 0.3, 1.0]))
def test_voxel_downsample_round_trips(p, voxel):
    from lidar_ai_recommendation_software_b200 import ops
    from oracle import new_ops
    pts = np.zeros((p.shape[0], 4), dtype=np.float32)
    pts[:, :3] = p
    pts[:, 3] = np.linspace(0, 1, p.shape[0], dtype=np.float32)
    res = ops.voxel_downsample(torch.from_numpy(pts).cuda(), voxel)
    o = new_ops.voxel_downsample(pts, voxel)
    key, inv = res.voxel_key.cpu().numpy(), res.inverse.cpu().numpy()
    np.testing.assert_array_equal(key, o["voxel_key"])
    np.testing.assert_array_equal(inv, o["inverse"])
    uk = res.unique_keys.cpu().numpy()
    assert res.n_voxels == len(uk) == len(o["counts"]) and res.dims == o["dims"]
    np.testing.assert_array_equal(uk, np.unique(key))                 # ascending, no duplicates
    np.testing.assert_array_equal(uk[inv], key)                       # the inverse map round-trips
    cnt = res.counts.cpu().numpy()
    np.testing.assert_array_equal(cnt, o["counts"])
    assert cnt.sum() == p.shape[0]
    np.testing.assert_allclose(res.centroids.cpu().numpy(), o["centroids"], rtol=1e-6, atol=1e-6)
    # every member lies in the voxel its key names
    ijk = np.floor((pts[:, :3].astype(np.float64) - np.array(res.origin)) / voxel).astype(np.int64)
    np.testing.assert_array_equal((ijk[:, 0] * res.dims[1] + ijk[:, 1]) * res.dims[2] + ijk[:, 2], key)


@settings(**CFG)
@given(clouds3(), st.floats(-2.0, 0.5), st.floats(0.0, 2.5))
def test_roi_crop_is_the_numpy_mask(p, lo, width):
    from lidar_ai_recommendation_software_b200 import ops
    lo3 = np.array([lo, lo, lo])
    hi3 = lo3 + width
    d = torch.from_numpy(p).cuda()
    out, mask = ops.roi_crop(d, lo3, hi3)
    want = ((p >= lo3) & (p <= hi3)).all(axis=1)
    np.testing.assert_array_equal(mask.cpu().numpy().astype(bool), want)
    np.testing.assert_array_equal(out.cpu().numpy(), p[want])


@settings(**CFG)
@given(st.integers(1, 3000), st.integers(0, 2**31 - 1), st.sampled_from(["dup", "uniform", "special"]))
def test_select_kth_is_np_partition(n, seed, mode):
    from lidar_ai_recommendation_software_b200 import ops
    rng = np.random.default_rng(seed)
    if mode == "dup":
        x = rng.integers(-3, 4, size=n).astype(np.float64) * 0.5
    elif mode == "uniform":
        x = rng.normal(size=n)
    else:
        x = rng.choice(np.array([0.0, -0.0, 5e-324, -5e-324, 1e308, -1e308, np.inf, -np.inf, 1.0, -1.0]), size=n)
    k = int(rng.integers(0, n))
    col = torch.from_numpy(np.ascontiguousarray(np.stack([x, x, x], 1))).cuda()[:, 2]
    a, b = ops.select_kth(col, k)
    s = np.sort(x)
    assert a == s[k] and b == s[min(k + 1, n - 1)]


@settings(**CFG)
@given(clouds3(400), st.sampled_from([0.26, 0.3, 0.5, 0.75]), st.integers(1, 6))
def test_dbscan_labels_are_sklearns(p, eps, min_samples):
    from sklearn.cluster import DBSCAN
    from lidar_ai_recommendation_software_b200 import ops
    d = torch.from_numpy(p).cuda()
    want = DBSCAN(eps=eps, min_samples=min_samples).fit(p).labels_
    for dense in (True, False):
        ops.set_dbscan_dense(dense)
        try:
            labels, nc, guard = ops.dbscan(d, eps, min_samples, tol=0.0)
        finally:
            ops.set_dbscan_dense(True)
        np.testing.assert_array_equal(labels.cpu().numpy(), want)
        assert nc == (want.max() + 1 if want.size else 0) and guard == 0


@settings(**CFG)
@given(clouds3(500), st.sampled_from([0.25, 0.5, 1.0]))
def test_ball_count_is_kdtree_count(p, r):
    from sklearn.neighbors import KDTree
    from lidar_ai_recommendation_software_b200 import ops
    got = ops.ball_count(torch.from_numpy(p).cuda(), r).cpu().numpy()
    want = KDTree(p).query_radius(p, r=r, count_only=True)
    np.testing.assert_array_equal(got, want)


@settings(**CFG)
@given(st.integers(1, 3), st.integers(8, 700), st.integers(0, 2**31 - 1), st.sampled_from([0.3, 0.8]))
def test_fps_and_ball_query_invariants(batch, n, seed, radius):
    from lidar_ai_recommendation_software_b200 import pointnet2
    from oracle import new_ops
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(-1, 1, size=(batch, n, 3)).astype(np.float32)
    xyz[:, n // 2:] = np.round(xyz[:, n // 2:] * 4) / 4               # duplicates: ties in the argmax
    m = max(1, n // 4)
    d = torch.from_numpy(xyz).cuda()
    idx = pointnet2.furthest_point_sample(d, m).cpu().numpy()
    np.testing.assert_array_equal(idx, new_ops.furthest_point_sample(xyz, m))
    assert (idx[:, 0] == 0).all()
    new_xyz = np.take_along_axis(xyz, idx[..., None].astype(np.int64), axis=1)
    k = 8
    bq = pointnet2.ball_query(d, torch.from_numpy(new_xyz).cuda(), radius, k).cpu().numpy()
    np.testing.assert_array_equal(bq, new_ops.ball_query(xyz, new_xyz, radius, k))
    for b in range(batch):
        pts = xyz[b][bq[b]]                                           # (m, k, 3)
        d2 = ((pts - new_xyz[b][:, None, :]).astype(np.float32) ** 2).sum(-1)
        assert (d2 < np.float32(radius) ** 2 + 1e-6).all()            # the centre itself is always a hit


@settings(**CFG)
@given(clouds3(800))
def test_chained_preprocess_front_is_numpy(p):
    from lidar_ai_recommendation_software_b200 import ops
    n = p.shape[0]
    desc, inl, col, ng, idx, _ = ops.preprocess_front(torch.from_numpy(p).cuda(), want_colors=True)
    mean, std = p.mean(axis=0), p.std(axis=0)
    np.testing.assert_allclose(np.array(desc.mean), mean, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(np.array(desc.std), std, rtol=1e-12, atol=1e-13)
    if int(desc.guard_sigma):
        return                                       # a point sits on the 3-sigma knife edge: certificate raised, no claim
    mask = (np.abs(p - mean) < 3 * std).all(axis=1)
    np.testing.assert_array_equal(inl.cpu().numpy(), p[mask])
    kept = p[mask]
    if kept.shape[0] == 0:
        assert int(desc.n_in) == 0
        return
    z = kept[:, 2]
    assert desc.z_thr == float(np.percentile(z, 30))
    ground = z <= desc.z_thr
    np.testing.assert_array_equal(ng.cpu().numpy(), kept[~ground])
    np.testing.assert_array_equal(idx.cpu().numpy(), np.nonzero(~ground)[0])
    assert int(round(desc.plane[0])) == int(ground.sum())
    np.testing.assert_array_equal(np.array(desc.bbox_in), np.concatenate([kept.min(axis=0), kept.max(axis=0)]))
    h = (kept[:, 2] - p[:, 2].min()) / (p[:, 2].max() - p[:, 2].min() + 1e-10)
    np.testing.assert_array_equal(col.cpu().numpy(), np.stack([h, 0.5 * (1 - h), np.full_like(h, 0.5)], 1))


@pytest.mark.parametrize("rings,azimuth", [(24, 6000), (40, 4096)])
def test_dbscan_on_a_ring_scan_is_sklearns(rings, azimuth):
    """Scan-ordered sensor frame: cells of hundreds of returns and neighbouring ground rings that never merge — the
    case the heavy-cell separation certificate of db_union_dense exists for."""
    from sklearn.cluster import DBSCAN
    from lidar_ai_recommendation_software_b200 import ops, synth
    f = synth.ring_sequence_frame(1, rings=rings, azimuth_steps=azimuth, n_people=150)[:, :3].astype(np.float64)
    f = f[f[:, 2] > np.percentile(f[:, 2], 30)]
    f = np.ascontiguousarray(f[np.linalg.norm(f[:, :2], axis=1) < 40.0])
    assert 20_000 < f.shape[0] < 200_000
    want = DBSCAN(eps=0.3, min_samples=5).fit(f).labels_
    d = torch.from_numpy(f).cuda()
    for dense in (True, False):
        ops.set_dbscan_dense(dense)
        try:
            labels, nc, guard = ops.dbscan(d, 0.3, 5, tol=0.0)
        finally:
            ops.set_dbscan_dense(True)
        np.testing.assert_array_equal(labels.cpu().numpy(), want)
        assert nc == want.max() + 1 and guard == 0
