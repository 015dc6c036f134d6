"""CPU tests: the oracle (oracle/) pinned against the unmodified reference through tests/golden/.

The golden vectors were produced by tests/golden/make_golden.py, which imports /root/reference in the
build container; here only the stored outputs are used, so this runs anywhere.
"""
import hashlib

import numpy as np
import pytest

from oracle import new_ops, np_semantics as nps, ref_path

CASE_NAMES = ["ref_sample_10k", "crowd_20k", "crowd_100k", "tiny_14", "sparse_300"]


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


# ---- third-party semantics restated vs the real libraries ------------------------------------
def test_arange_fill_rule_matches_numpy():
    rng = np.random.default_rng(0)
    for _ in range(500):
        a = rng.uniform(-100, 100)
        step = rng.choice([0.05, 0.1, 0.25, 0.3, 0.5, 1.0, 1.7])
        b = a + rng.uniform(0.1, 150)
        assert np.array_equal(nps.arange_f64(a, b, step), np.arange(a, b, step))


def test_linspace_edges_match_numpy():
    rng = np.random.default_rng(1)
    for _ in range(200):
        lo = rng.uniform(-100, 100)
        hi = lo + rng.uniform(0.1, 200)
        bins = int(rng.integers(1, 300))
        assert np.array_equal(nps.linspace_edges(lo, hi, bins), np.linspace(lo, hi, bins + 1))


def test_histogram2d_counts_match_numpy_including_edges():
    rng = np.random.default_rng(2)
    ex = np.arange(-3.0, 3.0 + 0.5, 0.5)
    ey = np.linspace(-2.0, 2.0, 9)
    u = rng.uniform(-4, 4, 5000)
    v = rng.uniform(-3, 3, 5000)
    # points exactly on edges, on the last edge, outside, and NaN-free extremes
    u[:20] = np.repeat(ex[[0, 3, -1, -2]], 5)
    v[:20] = np.tile(ey[[0, 2, -1, 4, -2]], 4)
    want, _, _ = np.histogram2d(u, v, bins=[ex, ey])
    assert np.array_equal(nps.histogram2d_counts(u, v, ex, ey), want.astype(np.int64))


def test_percentile_linear_matches_numpy():
    rng = np.random.default_rng(3)
    for n in (1, 2, 3, 10, 11, 1000, 1001, 7919):
        x = rng.normal(size=n)
        for q in (0, 30, 50, 77.7, 100):
            assert nps.percentile_linear(x, q) == float(np.percentile(x, q))


def test_mean_std_axis0_is_sequential():
    rng = np.random.default_rng(4)
    p = rng.uniform(-50, 50, (3000, 3))
    m, s = nps.mean_std_axis0(p)
    assert np.array_equal(m, np.mean(p, axis=0))
    assert np.allclose(s, np.std(p, axis=0), rtol=1e-14, atol=0)


def test_dbscan_closed_form_matches_sklearn():
    DBSCAN = pytest.importorskip("sklearn.cluster").DBSCAN
    rng = np.random.default_rng(5)
    for seed, eps in ((0, 0.3), (1, 0.5), (2, 0.15)):
        r = np.random.default_rng(seed)
        centres = r.uniform(-5, 5, (30, 3))
        X = np.concatenate([centres[r.integers(0, 30, 3000)] + r.normal(0, 0.15, (3000, 3)),
                            r.uniform(-6, 6, (600, 3))])
        X = X[r.permutation(len(X))]
        want = DBSCAN(eps=eps, min_samples=5).fit(X).labels_
        assert np.array_equal(nps.dbscan_labels(X, eps, 5), want)


def test_standard_scale_matches_sklearn():
    SS = pytest.importorskip("sklearn.preprocessing").StandardScaler
    rng = np.random.default_rng(6)
    X = rng.normal(3.0, [1.0, 20.0, 0.3], (5000, 3))
    got, mean, scale = ref_path.standard_scale(X)
    sc = SS()
    want = sc.fit_transform(X)
    assert np.array_equal(mean, sc.mean_) and np.array_equal(scale, sc.scale_)
    assert np.array_equal(got, want)


def test_radius_count_matches_kdtree():
    KDTree = pytest.importorskip("sklearn.neighbors").KDTree
    rng = np.random.default_rng(7)
    c = rng.uniform(-10, 10, (300, 2))
    q = rng.uniform(-10, 10, (500, 2))
    q[:10] = c[:10] + np.array([2.0, 0.0])  # exactly on the radius
    want = KDTree(c).query_radius(q, r=2.0, count_only=True)
    assert np.array_equal(nps.radius_count(c, q, 2.0), want)


def test_local_density_counts_match_kdtree():
    """visualisation local density (utils/visualization.py:43-45, 164-168): 3-D and 2-D, points on the radius."""
    KDTree = pytest.importorskip("sklearn.neighbors").KDTree
    rng = np.random.default_rng(11)
    p3 = np.concatenate([rng.uniform(-3, 3, (1500, 3)), rng.normal(0, 0.2, (1500, 3))])
    p3[:20] = p3[20:40] + np.array([0.5, 0.0, 0.0])      # exactly r apart
    assert np.array_equal(nps.local_density_counts(p3, 0.5), KDTree(p3).query_radius(p3, r=0.5, count_only=True))
    p2 = p3[:, :2]
    assert np.array_equal(nps.local_density_counts(p2, 0.5), KDTree(p2).query_radius(p2, r=0.5, count_only=True))


# ---- REF rows vs golden vectors of the unmodified reference ------------------------------------
@pytest.mark.parametrize("name", CASE_NAMES)
def test_preprocess_variant_a(name, golden, case_points):
    g, pts = golden(name), case_points(name)
    assert sha(pts) == str(g["input_sha"])
    out = ref_path.preprocess_lidar_data(pts)
    assert len(out["points"]) == int(g["a_n_inliers"])
    assert sha(out["points"]) == str(g["a_points_sha"])
    assert sha(out["colors"]) == str(g["a_colors_sha"])
    assert np.array_equal(out["clusters"], g["a_clusters"])
    assert np.allclose(out["ground_plane"], g["a_plane"], rtol=1e-9, atol=1e-12)
    d = out["dimensions"]
    dims = np.array([*d["x_range"], *d["y_range"], *d["z_range"], d["width"], d["length"], d["height"]])
    assert np.array_equal(dims, g["a_dims"])
    assert np.allclose(ref_path.extract_people_positions(out), g["a_people"], rtol=1e-12, atol=0)


@pytest.mark.parametrize("name", CASE_NAMES)
def test_preprocess_variant_b(name, golden, case_points):
    g, pts = golden(name), case_points(name)
    out = ref_path.preprocess_point_cloud(pts)
    assert len(out["points"]) == int(g["b_n_inliers"])
    assert sha(out["colors"]) == str(g["b_colors_sha"])
    assert np.array_equal(out["clusters"], g["b_clusters"])


@pytest.mark.parametrize("name", CASE_NAMES)
def test_grid_density_and_heatmap(name, golden, case_points):
    g, pts = golden(name), case_points(name)
    pa = ref_path.preprocess_lidar_data(pts)
    d = pa["dimensions"]
    for gs in (1.0, 0.5):
        gx, gy, dens = ref_path.calculate_grid_density(pa["points"][:, :2], d["x_range"], d["y_range"], gs)
        assert np.array_equal(np.rint(dens * gs * gs).astype(np.int32), g[f"a_grid_counts_g{gs}"])
        assert np.array_equal(gx, g[f"a_grid_x_g{gs}"]) and np.array_equal(gy, g[f"a_grid_y_g{gs}"])
    counts, ex, ey = ref_path.heatmap_counts(pa)
    assert np.array_equal(counts, g["heat_counts"])
    assert np.array_equal(ex, g["heat_ex"]) and np.array_equal(ey, g["heat_ey"])
    assert ref_path.calculate_grid_density(np.zeros((0, 2)), (0, 1), (0, 1)) == (None, None, None)


def _hot(hs, key="density"):
    return np.array([[h["x"], h["y"], h[key]] for h in hs], dtype=np.float64).reshape(-1, 3)


@pytest.mark.parametrize("name", CASE_NAMES)
def test_density_models(name, golden, case_points):
    g, pts = golden(name), case_points(name)
    ra = ref_path.density_analyze(ref_path.preprocess_lidar_data(pts))
    assert np.array_equal([ra["total_people"], ra["avg_density"], ra["max_density"]], g["a_density_scalars"])
    assert np.array_equal(ra["density_map"], g["a_density_map"])
    assert np.array_equal(_hot(ra["hotspots"]), g["a_hotspots"])
    rb = ref_path.analyze_crowd_density_b(ref_path.preprocess_point_cloud(pts))
    assert np.allclose([rb["total_people"], rb["avg_density"], rb["max_density"]], g["b_density_scalars"], rtol=1e-12)
    assert np.array_equal(rb["density_grid"], g["b_density_grid"])
    assert np.allclose(_hot(rb["hotspots"]), g["b_hotspots"], rtol=1e-12)


@pytest.mark.parametrize("name", CASE_NAMES)
def test_flow_models(name, golden, case_points):
    g, pts = golden(name), case_points(name)
    fa = ref_path.flow_analyze(ref_path.preprocess_lidar_data(pts))
    assert sha(fa["flow_vectors"]["positions"]) == str(g["a_flow_positions_sha"])
    assert np.allclose(fa["flow_vectors"]["vectors"], g["a_flow_vectors"], rtol=1e-12, atol=1e-15)
    assert np.allclose(fa["flow_vectors"]["magnitudes"], g["a_flow_magnitudes"], rtol=1e-12)
    assert np.isclose(fa["avg_speed"], g["a_flow_scalars"][0], rtol=1e-12)
    assert fa["dominant_direction"] == str(g["a_flow_direction"])
    assert np.allclose(_hot(fa["bottlenecks"], "severity"), g["a_bottlenecks"], rtol=1e-12)
    fb = ref_path.analyze_crowd_flow_b(ref_path.preprocess_point_cloud(pts))
    assert np.allclose(fb["flow_vectors"]["vectors"], g["b_flow_vectors"], rtol=1e-12, atol=1e-15)
    assert np.isclose(fb["avg_speed"], g["b_flow_scalars"][0], rtol=1e-12)
    assert fb["dominant_direction"] == str(g["b_flow_direction"])
    assert np.allclose(_hot(fb["bottlenecks"], "severity"), g["b_bottlenecks"], rtol=1e-12)


def test_risk_levels():
    assert [ref_path.risk_level(d) for d in (0.0, 0.99, 1.0, 2.49, 2.5, 3.99, 4.0, 9.0)] == \
        ["Low", "Low", "Moderate", "Moderate", "High", "High", "Critical", "Critical"]


# ---- NEW ops: self-consistency of the frozen definitions ---------------------------------------
def test_voxel_downsample_properties():
    from lidar_ai_recommendation_software_b200 import synth
    pts = synth.crowd_frame(20000, seed=1, extent=10.0)
    r = new_ops.voxel_downsample(pts, 0.05)
    assert r["counts"].sum() == len(pts)
    assert np.all(np.diff(r["unique_keys"]) > 0)
    assert np.array_equal(r["unique_keys"][r["inverse"]], r["voxel_key"])
    # every centroid lies inside its voxel (up to fp32 rounding of the mean)
    org = np.array(r["origin"])
    ijk = np.floor((r["centroids"][:, :3].astype(np.float64) - org) / 0.05 + 1e-6)
    dims = r["dims"]
    key = (ijk[:, 0] * dims[1] + ijk[:, 1]) * dims[2] + ijk[:, 2]
    assert np.mean(key == r["unique_keys"]) > 0.999


def test_roi_crop_and_fps_ballquery_small():
    rng = np.random.default_rng(0)
    pts = rng.uniform(-1, 1, (500, 4)).astype(np.float32)
    out, m = new_ops.roi_crop(pts, (-0.5, -0.5, -0.5), (0.5, 0.5, 0.5))
    assert out.shape[0] == m.sum() and np.array_equal(out, pts[m])
    xyz = rng.uniform(-1, 1, (2, 256, 3)).astype(np.float32)
    idx = new_ops.furthest_point_sample(xyz, 32)
    assert idx.shape == (2, 32) and all(len(set(r)) == 32 for r in idx) and np.all(idx[:, 0] == 0)
    new_xyz = np.stack([xyz[b][idx[b]] for b in range(2)])
    bq = new_ops.ball_query(xyz, new_xyz, 0.4, 8)
    assert bq.shape == (2, 32, 8)
    g = new_ops.group_points(xyz, None, bq, new_xyz)
    assert g.shape == (2, 3, 32, 8)
    assert np.all((g ** 2).sum(1) < 0.4 ** 2 + 1e-6)
