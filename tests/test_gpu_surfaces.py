"""GPU parity tests of the two drop-in call surfaces against the golden outputs of the UNMODIFIED
reference (tests/golden/*.npz, produced by make_golden.py from /root/reference) and the CPU oracle.

Surface A: utils.data_processing + models.crowd_density_model + models.crowd_flow_model
Surface B: the functions inlined in app_simplified.py / app_with_db.py
"""
import hashlib

import numpy as np
import pytest
import torch

from oracle import new_ops, np_semantics as nps, ref_path

pytestmark = pytest.mark.gpu

CASE_NAMES = ["ref_sample_10k", "crowd_20k", "crowd_100k", "tiny_14", "sparse_300"]


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def _hot(hs, key="density"):
    return np.array([[h["x"], h["y"], h[key]] for h in hs], dtype=np.float64).reshape(-1, 3)


@pytest.fixture(scope="module")
def pkg():
    import lidar_ai_recommendation_software_b200 as p
    from lidar_ai_recommendation_software_b200 import apps, ops, preprocess, synth
    from lidar_ai_recommendation_software_b200.models.crowd_density_model import CrowdDensityModel
    from lidar_ai_recommendation_software_b200.models.crowd_flow_model import CrowdFlowModel
    from lidar_ai_recommendation_software_b200.utils import data_processing

    class NS:
        pass

    ns = NS()
    ns.apps, ns.ops, ns.pre, ns.dp, ns.synth = apps, ops, preprocess, data_processing, synth
    ns.CDM, ns.CFM = CrowdDensityModel, CrowdFlowModel
    return ns


@pytest.fixture(scope="module")
def processed_a(pkg, case_points):
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = pkg.dp.preprocess_lidar_data(case_points(name))
        return cache[name]

    return get


@pytest.fixture(scope="module")
def processed_b(pkg, case_points):
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = pkg.apps.preprocess_point_cloud(case_points(name))
        return cache[name]

    return get


# ---- surface A ---------------------------------------------------------------------------------
@pytest.mark.parametrize("name", CASE_NAMES)
def test_preprocess_lidar_data_matches_reference(name, pkg, golden, processed_a):
    g, out = golden(name), processed_a(name)
    guards = out[pkg.pre.DEVICE_KEY].guards
    assert guards["sigma"] == 0 and guards["dbscan"] == 0, f"knife-edge certificate failed: {guards}"
    assert list(k for k in out if not k.startswith("_")) == ["points", "colors", "normals", "clusters",
                                                              "ground_plane", "dimensions"]
    assert out["points"].dtype == np.float64 and out["clusters"].dtype == np.int64
    assert len(out["points"]) == int(g["a_n_inliers"])
    assert sha(out["points"]) == str(g["a_points_sha"])                      # mask + compaction bit-exact
    assert sha(out["colors"]) == str(g["a_colors_sha"])                      # colours bit-exact
    assert np.array_equal(out["clusters"], g["a_clusters"])                  # DBSCAN labels bit-exact
    assert np.allclose(out["ground_plane"], g["a_plane"], rtol=1e-9, atol=1e-12)
    d = out["dimensions"]
    dims = np.array([*d["x_range"], *d["y_range"], *d["z_range"], d["width"], d["length"], d["height"]])
    assert np.array_equal(dims, g["a_dims"])
    assert np.array_equal(out["normals"][:, 2], np.ones(len(out["points"]))) and not out["normals"][:, :2].any()


@pytest.mark.parametrize("name", CASE_NAMES)
def test_people_positions_and_grid_density(name, pkg, golden, processed_a):
    g, pa = golden(name), processed_a(name)
    pos = pkg.dp.extract_people_positions(pa)
    assert pos.shape == g["a_people"].shape
    assert np.allclose(pos, g["a_people"], rtol=1e-12, atol=0)
    # a dict rebuilt by a caller (no device cache) goes through the upload path and agrees
    plain = {"points": pa["points"].copy(), "clusters": pa["clusters"].copy()}
    assert np.array_equal(pkg.dp.extract_people_positions(plain), pos)
    d = pa["dimensions"]
    for gs in (1.0, 0.5):
        gx, gy, dens = pkg.dp.calculate_grid_density(pa["points"][:, :2], d["x_range"], d["y_range"], gs)
        assert np.array_equal(np.rint(dens * gs * gs).astype(np.int32), g[f"a_grid_counts_g{gs}"])
        assert np.array_equal(gx, g[f"a_grid_x_g{gs}"]) and np.array_equal(gy, g[f"a_grid_y_g{gs}"])
        assert dens.dtype == np.float64
    assert pkg.dp.calculate_grid_density(np.zeros((0, 2)), (0, 1), (0, 1)) == (None, None, None)
    hist, ex, ey = pkg.apps.density_heatmap_counts(pa, bins=100)
    assert np.array_equal(hist.astype(np.int32), g["heat_counts"])
    assert np.array_equal(ex, g["heat_ex"]) and np.array_equal(ey, g["heat_ey"])


def test_create_density_heatmap_under_its_own_name(pkg, golden, processed_a, monkeypatch):
    """Surface B swaps `create_density_heatmap` in by name (app_simplified.py:198-232, called at :1074): same
    go.Figure(go.Heatmap(z=hist.T, x=centres, y=centres, ...)) + layout.  plotly is not in this image, so a recording
    stand-in of plotly.graph_objects shows what the figure is built from."""
    import sys
    import types

    class _Rec:
        def __init__(self, *a, **k):
            self.args, self.kwargs, self.layout = a, k, {}

        def update_layout(self, **k):
            self.layout.update(k)

    go = types.ModuleType("plotly.graph_objects")
    go.Figure = type("Figure", (_Rec,), {})
    go.Heatmap = type("Heatmap", (_Rec,), {})
    plotly = types.ModuleType("plotly")
    plotly.graph_objects = go
    monkeypatch.setitem(sys.modules, "plotly", plotly)
    monkeypatch.setitem(sys.modules, "plotly.graph_objects", go)
    g, pa = golden("crowd_20k"), processed_a("crowd_20k")
    fig = pkg.apps.create_density_heatmap(pa)
    hm = fig.kwargs["data"]
    assert isinstance(fig, go.Figure) and isinstance(hm, go.Heatmap)
    assert np.array_equal(hm.kwargs["z"], g["heat_counts"].T.astype(np.float64)) and hm.kwargs["z"].dtype == np.float64
    ex, ey = g["heat_ex"], g["heat_ey"]
    assert np.array_equal(hm.kwargs["x"], (ex[:-1] + ex[1:]) / 2) and np.array_equal(hm.kwargs["y"], (ey[:-1] + ey[1:]) / 2)
    assert hm.kwargs["colorscale"] == "Viridis" and hm.kwargs["colorbar"] == dict(title="Point Density")
    assert fig.layout == dict(xaxis_title="X (m)", yaxis_title="Y (m)", title="Point Density Heatmap", height=500)


@pytest.mark.parametrize("name", CASE_NAMES)
def test_crowd_density_model(name, pkg, golden, processed_a):
    g = golden(name)
    model = pkg.CDM()
    ra = model.analyze(processed_a(name))
    assert list(ra) == ["total_people", "avg_density", "max_density", "density_map", "grid_coordinates",
                        "density_values", "hotspots"]
    assert np.array_equal([ra["total_people"], ra["avg_density"], ra["max_density"]], g["a_density_scalars"])
    assert np.array_equal(ra["density_map"], g["a_density_map"])
    assert np.array_equal(_hot(ra["hotspots"]), g["a_hotspots"])
    assert [model.calculate_risk_level(v) for v in (0.5, 1.0, 2.5, 4.0)] == ["Low", "Moderate", "High", "Critical"]
    empty = model.analyze({"points": np.zeros((5, 3)), "clusters": -np.ones(5, dtype=np.int64),
                           "dimensions": {"x_range": (0, 1), "y_range": (0, 1)}})
    assert empty["total_people"] == 0 and empty["density_map"].shape == (1, 1) and empty["hotspots"] == []


@pytest.mark.parametrize("name", CASE_NAMES)
def test_crowd_flow_model(name, pkg, golden, processed_a):
    g = golden(name)
    np.random.seed(123)
    fa = pkg.CFM().analyze(processed_a(name))
    # the reference re-seeds the GLOBAL legacy RNG with 42 and draws 6 uniforms: observable side effect
    state_after = np.random.uniform()
    np.random.seed(42)
    [np.random.uniform() for _ in range(6)]
    assert state_after == np.random.uniform()
    assert sha(fa["flow_vectors"]["positions"]) == str(g["a_flow_positions_sha"])
    assert np.allclose(fa["flow_vectors"]["vectors"], g["a_flow_vectors"], rtol=1e-9, atol=1e-12)
    assert np.allclose(fa["flow_vectors"]["magnitudes"], g["a_flow_magnitudes"], rtol=1e-9)
    assert np.isclose(fa["avg_speed"], g["a_flow_scalars"][0], rtol=1e-9)
    assert fa["dominant_direction"] == str(g["a_flow_direction"])
    assert np.allclose(_hot(fa["bottlenecks"], "severity"), g["a_bottlenecks"], rtol=1e-12)
    empty = pkg.CFM().analyze({"points": np.zeros((5, 3)), "clusters": -np.ones(5, dtype=np.int64),
                               "dimensions": {"x_range": (0, 1), "y_range": (0, 1)}})
    assert empty["dominant_direction"] == "N/A" and empty["flow_vectors"]["positions"].shape == (0, 2)


# ---- surface B ---------------------------------------------------------------------------------
@pytest.mark.parametrize("name", CASE_NAMES)
def test_preprocess_point_cloud_matches_reference(name, pkg, golden, processed_b):
    g, out = golden(name), processed_b(name)
    assert list(k for k in out if not k.startswith("_")) == ["points", "colors", "clusters", "dimensions"]
    assert out[pkg.pre.DEVICE_KEY].guards["sigma"] == 0
    assert len(out["points"]) == int(g["b_n_inliers"])
    assert sha(out["colors"]) == str(g["b_colors_sha"])
    assert np.array_equal(out["clusters"], g["b_clusters"])                  # 100s of clusters, bit-exact


@pytest.mark.parametrize("name", CASE_NAMES)
def test_analyze_crowd_density_b(name, pkg, golden, processed_b):
    g = golden(name)
    rb = pkg.apps.analyze_crowd_density(processed_b(name))
    assert list(rb) == ["total_people", "avg_density", "max_density", "density_grid", "hotspots"]
    assert np.allclose([rb["total_people"], rb["avg_density"], rb["max_density"]], g["b_density_scalars"], rtol=1e-12)
    assert np.array_equal(rb["density_grid"], g["b_density_grid"])          # integer radius counts / 4
    assert np.allclose(_hot(rb["hotspots"]), g["b_hotspots"], rtol=1e-12)


@pytest.mark.parametrize("name", CASE_NAMES)
def test_analyze_crowd_flow_b(name, pkg, golden, processed_b):
    g = golden(name)
    fb = pkg.apps.analyze_crowd_flow(processed_b(name))
    assert np.allclose(fb["flow_vectors"]["vectors"], g["b_flow_vectors"], rtol=1e-9, atol=1e-12)
    assert np.allclose(fb["flow_vectors"]["magnitudes"], g["b_flow_magnitudes"], rtol=1e-9)
    assert np.isclose(fb["avg_speed"], g["b_flow_scalars"][0], rtol=1e-9)
    assert fb["dominant_direction"] == str(g["b_flow_direction"])
    assert np.allclose(_hot(fb["bottlenecks"], "severity"), g["b_bottlenecks"], rtol=1e-12)


# ---- stage-level parity vs the oracle ------------------------------------------------------------
def test_select_kth_matches_sort(pkg):
    rng = np.random.default_rng(0)
    for n in (1, 2, 7, 1000, 100001):
        x = rng.normal(size=(n, 3))
        x[: n // 3, 2] = x[0, 2]          # ties
        x[n // 2:, 2] *= -1.0             # both signs
        d = torch.from_numpy(x).cuda()
        s = np.sort(x[:, 2])
        for k in sorted({0, n // 3, max(n - 2, 0), n - 1}):
            a, b = pkg.ops.select_kth(d[:, 2], k)
            assert a == s[k] and b == s[min(k + 1, n - 1)]
        thr = pkg.pre.percentile_from_order_stats(*pkg.ops.select_kth(d[:, 2], int(np.floor((n - 1) * 0.3))), n, 30)
        assert thr == float(np.percentile(x[:, 2], 30))
    # the buffered path (n >= 65536: keys with the winning 16-bit prefix are compacted after two passes) on degenerate
    # columns: every value equal (the buffer holds all of them), two values, a height column with heavy ties
    for col in (np.full(70_000, 1.25), np.repeat([0.5, -0.5], 40_000), np.round(rng.uniform(0, 2, 300_000), 2)):
        d = torch.from_numpy(np.column_stack([col, col, col])).cuda()
        s = np.sort(col)
        n = len(col)
        for k in sorted({0, n // 3, n // 2, n - 2, n - 1}):
            a, b = pkg.ops.select_kth(d[:, 2], k)
            assert a == s[k] and b == s[min(k + 1, n - 1)]


@pytest.mark.parametrize("eps,seed", [(0.3, 0), (0.5, 1), (0.15, 2), (1.5, 3)])
def test_dbscan_labels_match_oracle(pkg, eps, seed):
    r = np.random.default_rng(seed)
    centres = r.uniform(-5, 5, (40, 3))
    X = np.concatenate([centres[r.integers(0, 40, 6000)] + r.normal(0, 0.15, (6000, 3)), r.uniform(-6, 6, (1500, 3))])
    X = X[r.permutation(len(X))]
    want = nps.dbscan_labels(X, eps, 5)
    labels, nc, guard = pkg.ops.dbscan(torch.from_numpy(X).cuda(), eps, 5)
    assert guard == 0
    assert np.array_equal(labels.cpu().numpy().astype(np.int64), want)
    assert nc == (want.max() + 1 if (want >= 0).any() else 0)


@pytest.mark.parametrize("eps,min_samples", [(0.3, 5), (0.12, 3), (0.6, 8)])
def test_dbscan_dense_grid_equals_general_grid_and_oracle(pkg, eps, min_samples):
    """Scan-ordered, very dense returns (hundreds of points per eps-ball, like a 128-beam frame near the
    sensor) mixed with sparse clutter: the dense cell grid (full cells are core without a distance test,
    merged cells are skipped) must give sklearn's labels exactly, like the general grid."""
    r = np.random.default_rng(7)
    th = np.linspace(0, 2 * np.pi, 3000, endpoint=False)
    rings = [np.stack([rad * np.cos(th), rad * np.sin(th), 0.02 * np.sin(7 * th) + r.normal(0, 0.004, th.size)], 1)
             for rad in (1.0, 1.1, 1.25, 2.0)]
    blobs = r.uniform(-3, 3, (25, 3)) * [1, 1, 0.3]
    people = blobs[r.integers(0, 25, 3000)] + r.normal(0, 0.08, (3000, 3))
    clutter = r.uniform(-3.5, 3.5, (1000, 3)) * [1, 1, 0.5]
    X = np.concatenate(rings + [people, clutter])
    want = nps.dbscan_labels(X, eps, min_samples)
    d = torch.from_numpy(X).cuda()
    try:
        pkg.ops.set_dbscan_dense(True)
        lab_d, nc_d, _ = pkg.ops.dbscan(d, eps, min_samples)
        pkg.ops.set_dbscan_dense(False)
        lab_g, nc_g, _ = pkg.ops.dbscan(d, eps, min_samples)
    finally:
        pkg.ops.set_dbscan_dense(True)
    assert np.array_equal(lab_g.cpu().numpy().astype(np.int64), want)
    assert np.array_equal(lab_d.cpu().numpy().astype(np.int64), want)
    assert nc_d == nc_g == (want.max() + 1 if (want >= 0).any() else 0)


def test_dbscan_dense_grid_ring_frame_equals_general_grid(pkg):
    """A quarter-resolution 128-beam frame (too big for the CPU oracle in seconds): both grids must agree."""
    f = pkg.synth.ring_sequence_frame(3, rings=64, azimuth_steps=8192)
    X = np.ascontiguousarray(f[f[:, 2] > 0.15][:, :3], dtype=np.float64)
    d = torch.from_numpy(X).cuda()
    try:
        pkg.ops.set_dbscan_dense(True)
        lab_d, nc_d, _ = pkg.ops.dbscan(d, 0.3, 5)
        pkg.ops.set_dbscan_dense(False)
        lab_g, nc_g, _ = pkg.ops.dbscan(d, 0.3, 5)
    finally:
        pkg.ops.set_dbscan_dense(True)
    assert nc_d == nc_g and nc_d > 10
    assert torch.equal(lab_d, lab_g)


def test_dbscan_degenerate_inputs(pkg):
    # all points identical, fewer than min_samples, exactly on the eps boundary (inclusive)
    same = np.zeros((10, 3))
    lab, nc, _ = pkg.ops.dbscan(torch.from_numpy(same).cuda(), 0.3, 5)
    assert nc == 1 and np.array_equal(lab.cpu().numpy(), np.zeros(10))
    few = np.arange(12, dtype=np.float64).reshape(4, 3)
    lab, nc, _ = pkg.ops.dbscan(torch.from_numpy(few).cuda(), 0.3, 5)
    assert nc == 0 and np.all(lab.cpu().numpy() == -1)
    line = np.zeros((6, 3))
    line[:, 0] = np.arange(6) * 0.25       # neighbours exactly 0.25 and 0.5 apart, eps = 0.5 inclusive
    want = nps.dbscan_labels(line, 0.5, 5)
    lab, _, _ = pkg.ops.dbscan(torch.from_numpy(line).cuda(), 0.5, 5)
    assert np.array_equal(lab.cpu().numpy().astype(np.int64), want)


@pytest.mark.parametrize("r", [0.5, 0.05, 2.0])
def test_local_point_density_matches_oracle(pkg, r):
    """KDTree.query_radius(count_only=True) of the visualisation paths, 3-D and 2-D projection, bit-exact counts:
    dense rings (cells entirely inside the ball are counted wholesale), clutter, and pairs exactly r apart."""
    from lidar_ai_recommendation_software_b200.utils import visualization as viz
    rng = np.random.default_rng(5)
    th = np.linspace(0, 2 * np.pi, 4000, endpoint=False)
    ring = np.stack([1.5 * np.cos(th), 1.5 * np.sin(th), rng.normal(0, 0.01, th.size)], 1)
    pts = np.concatenate([ring, rng.uniform(-3, 3, (3000, 3)), rng.normal(0, 0.15, (3000, 3))])
    pts[:30] = pts[30:60] + np.array([r, 0.0, 0.0])
    got = viz.local_point_density(pts, r)
    assert got.dtype == np.int64 and np.array_equal(got, nps.local_density_counts(pts, r))
    got2 = viz.local_point_density(pts[:, :2], r)
    assert np.array_equal(got2, nps.local_density_counts(pts[:, :2], r))
    assert viz.local_point_density(np.zeros((0, 3)), r).shape == (0,)


def test_distance_from_center_on_the_device(pkg, case_points):
    """utils/visualization.py:50-54 (np.mean + np.sqrt(np.sum((p - c)**2, axis=1))) through the device kernels."""
    from lidar_ai_recommendation_software_b200.utils import visualization as viz
    pts = case_points("crowd_20k")
    want = np.sqrt(np.sum((pts - np.mean(pts, axis=0)) ** 2, axis=1))
    got = viz.distance_from_center(pts)
    assert got.dtype == np.float64 and got.shape == want.shape
    assert np.allclose(got, want, rtol=1e-12, atol=1e-12)
    far = pts + np.array([4.0e5, 5.0e6, 0.0])        # UTM-sized offsets: the centroid must not lose the small scale
    assert np.allclose(viz.distance_from_center(far), np.sqrt(np.sum((far - np.mean(far, axis=0)) ** 2, axis=1)),
                       rtol=1e-9, atol=1e-7)


def test_projection_histogram_matches_numpy_semantics(pkg, processed_b, case_points):
    from lidar_ai_recommendation_software_b200.utils import visualization as viz
    pd = processed_b("crowd_20k")
    for dims in (("x", "y"), ("x", "z"), ("y", "z")):
        heat, xc, yc = viz.projection_histogram(pd, dims, 64)
        want, ex, ey = ref_path.heatmap_counts(pd, bins=64, dims=tuple("xyz".index(d) for d in dims))
        assert np.array_equal(heat, np.asarray(want, dtype=np.float64).T)
        assert np.array_equal(xc, (ex[:-1] + ex[1:]) / 2) and np.array_equal(yc, (ey[:-1] + ey[1:]) / 2)


def test_windows_core_run_analysis_socket(pkg, case_points):
    """ProjectManager.run_analysis keys (project_manager.py:338-348), JSON-serialisable, consistent with the models."""
    import json
    from lidar_ai_recommendation_software_b200.windows_core import convert_numpy, run_analysis
    pts = case_points("crowd_20k")
    res = run_analysis(pts, {"grid_size": 1.0})
    assert list(res) == ["total_people", "avg_density", "max_density", "density_map", "hotspots", "avg_speed",
                         "dominant_direction", "bottlenecks", "timestamp"]
    json.dumps(convert_numpy(res))
    pd = pkg.apps.preprocess_point_cloud(pts)
    dens = pkg.CDM(grid_size=1.0).analyze(pd)
    flow = pkg.CFM().analyze(pd)
    assert res["total_people"] == dens["total_people"] and np.array_equal(res["density_map"], dens["density_map"])
    assert res["max_density"] == dens["max_density"] and res["dominant_direction"] == flow["dominant_direction"]
    assert [b["severity"] for b in res["bottlenecks"]] == [b["severity"] for b in flow["bottlenecks"]]
    res_a = run_analysis(pts, {"variant": "A"})
    assert res_a["total_people"] == 1       # variant A merges the venue into one cluster, like the reference
    with pytest.raises(ValueError):
        run_analysis(pts, {"variant": "C"})


def test_cluster_centroids_exact(pkg):
    rng = np.random.default_rng(1)
    pts = rng.uniform(-80, 80, (50000, 3))
    lab = rng.integers(-1, 37, 50000).astype(np.int64)
    cent, counts = pkg.ops.cluster_centroids(torch.from_numpy(pts).cuda(), torch.from_numpy(lab).cuda(), 37)
    want = np.array([pts[lab == c].mean(0) for c in range(37)])
    assert np.allclose(cent.cpu().numpy(), want, rtol=1e-13, atol=1e-13)
    assert np.array_equal(counts.cpu().numpy(), np.bincount(lab[lab >= 0], minlength=37))
    # deterministic run to run
    c2, _ = pkg.ops.cluster_centroids(torch.from_numpy(pts).cuda(), torch.from_numpy(lab).cuda(), 37)
    assert torch.equal(cent, c2)


def test_downsample_point_cloud_contract(pkg):
    pts = np.random.default_rng(0).normal(size=(1000, 3))
    out = pkg.dp.downsample_point_cloud(pts, 0.1)
    assert out.shape == (100, 3) and out.dtype == pts.dtype
    rows = {tuple(r) for r in pts}
    assert all(tuple(r) in rows for r in out) and len({tuple(r) for r in out}) == 100
    assert pkg.dp.downsample_point_cloud(pts, 1.0) is pts
    assert pkg.dp.downsample_point_cloud(pts, 1e-9).shape == (1, 3)
    # the draw is the reference's own expression on the GLOBAL legacy RNG (utils/data_processing.py:245-249): with
    # the same seed the result is the reference's, row for row, and the RNG is left in the same state
    big = np.random.default_rng(1).normal(size=(200_000, 3))
    np.random.seed(123)
    got = pkg.dp.downsample_point_cloud(big, 0.1)
    after = np.random.random()
    np.random.seed(123)
    want = big[np.random.choice(len(big), max(1, int(len(big) * 0.1)), replace=False)]
    assert np.array_equal(got, want) and after == np.random.random()


def test_frame_flow_matches_oracle(pkg):
    rng = np.random.default_rng(2)
    prev = rng.uniform(-20, 20, (300, 2))
    cur = prev[rng.permutation(300)[:280]] + rng.normal(0.08, 0.03, (280, 2))
    cur = np.concatenate([cur, rng.uniform(-20, 20, (15, 2))])
    flow, match, vel = pkg.CFM.__module__ and __import__(
        "lidar_ai_recommendation_software_b200.flow", fromlist=["frame_flow"]).frame_flow(
        prev, cur, 0.1, (-20.0, 20.0), (-20.0, 20.0))
    wm, wv = new_ops.frame_flow_match(prev, cur, 0.1, 1.5)
    assert np.array_equal(match, wm)                                        # indices bit-exact
    assert np.array_equal(vel, wv)
    wvec, wmag = new_ops.frame_flow_field(flow["positions"], cur, wm, wv, 3.0)
    assert np.allclose(flow["vectors"], wvec, rtol=1e-3, atol=1e-9)
    assert np.allclose(flow["magnitudes"], wmag, rtol=1e-3, atol=1e-9)


def test_sequence_model_uses_previous_frame(pkg, processed_b):
    pb = processed_b("crowd_20k")
    model = pkg.CFM()
    first = model.analyze_sequence_frame(pb)
    assert first["dominant_direction"] != "N/A" and model.prev_positions is not None
    moved = dict(pb)
    moved.pop(pkg.pre.DEVICE_KEY)
    moved["points"] = pb["points"] + np.array([0.1, 0.0, 0.0])              # everybody walks +x at 1 m/s
    second = model.analyze_sequence_frame(moved, dt=0.1)
    assert second["dominant_direction"] == "E"
    m = second["flow_vectors"]["magnitudes"]
    assert np.allclose(m[m > 0], 1.0, rtol=1e-3)


def test_sequence_runner_equals_serial_loop(pkg):
    """Worker threads / streams of SequenceRunner change nothing: same clusters, matches and flow vectors."""
    from lidar_ai_recommendation_software_b200.sequence import SequenceRunner
    frames = [np.ascontiguousarray(pkg.synth.ring_sequence_frame(i, rings=48, azimuth_steps=4096)[:, :3], dtype=np.float64)
              for i in range(4)]
    serial_model = pkg.CFM()
    want = []
    for f in frames:
        pd = pkg.pre.run(f, variant="B", host_arrays=False)
        want.append((pd[pkg.pre.DEVICE_KEY].n_clusters, serial_model.analyze_sequence_frame(pd, dt=0.1)))
    runner = SequenceRunner(variant="B", workers=3, dt=0.1)
    got = [(pd[pkg.pre.DEVICE_KEY].n_clusters, res) for pd, res in runner.run(iter(frames))]
    runner.close()
    assert len(got) == len(want)
    for (nc_g, r_g), (nc_w, r_w) in zip(got, want):
        assert nc_g == nc_w and r_g["dominant_direction"] == r_w["dominant_direction"]
        assert np.array_equal(r_g["flow_vectors"]["vectors"], r_w["flow_vectors"]["vectors"])
        assert ("matches" in r_g) == ("matches" in r_w)
        if "matches" in r_w:
            assert np.array_equal(r_g["matches"], r_w["matches"])


@pytest.mark.gpu
def test_sequence_runner_state_and_empty_frames(pkg):
    """The concurrent flow step (a frame needs its predecessor's POSITIONS only) against the serial, stateful
    `analyze_sequence_frame`: every key of every result, across a frame without any cluster (the next frame falls back to
    the simulated field, as the first one does), across two `run` calls (the model's `prev_positions` carries over), and
    the state the model is left in."""
    from lidar_ai_recommendation_software_b200.sequence import SequenceRunner
    ring = [np.ascontiguousarray(pkg.synth.ring_sequence_frame(i, rings=48, azimuth_steps=4096)[:, :3], dtype=np.float64)
            for i in range(4)]
    noise = np.random.default_rng(3).uniform(-50.0, 50.0, (3000, 3))           # nothing within eps = 0.3 of anything
    frames = [ring[0], ring[1], noise, ring[2], ring[3], ring[0]]
    serial = pkg.CFM()
    want = [serial.analyze_sequence_frame(pkg.pre.run(f, variant="B", host_arrays=False), dt=0.1) for f in frames]
    assert want[2]["dominant_direction"] == "N/A" and "matches" not in want[3] and "matches" in want[4]
    for workers in (2, 4):
        runner = SequenceRunner(variant="B", workers=workers, dt=0.1)
        got = [res for _, res in runner.run(iter(frames[:3]))] + [res for _, res in runner.run(iter(frames[3:]))]
        runner.close()
        assert len(got) == len(want)
        for g, w in zip(got, want):
            assert sorted(g) == sorted(w)
            assert g["dominant_direction"] == w["dominant_direction"] and g["bottlenecks"] == w["bottlenecks"]
            if "matches" in w:        # measured field: np.mean on the host; the simulated one sums with device atomics
                assert np.array_equal(np.asarray(g["avg_speed"]), np.asarray(w["avg_speed"]))
            else:
                assert np.isclose(g["avg_speed"], w["avg_speed"], rtol=1e-12, atol=0.0)
            for k in ("positions", "vectors", "magnitudes"):
                assert np.array_equal(g["flow_vectors"][k], w["flow_vectors"][k]), k
            if "matches" in w:
                assert np.array_equal(g["matches"], w["matches"])
        assert np.array_equal(runner.model.prev_positions, serial.prev_positions)
        assert np.array_equal(runner.model.flow_vectors["vectors"], serial.flow_vectors["vectors"])


def test_run_sequence_frame_equals_the_separate_calls(pkg):
    """`preprocess.run_sequence_frame` (one C call: front, DBSCAN, labels, centroids, the waits between them in C) against
    `run(variant="B", host_arrays=False)` + `people_positions`: inliers, labels, cluster count, dimensions, guards and
    people positions, on ring frames, on a frame without clusters, on a frame with fewer than 11 non-ground points and
    with a centroid capacity that is too small (the positions then come from the ordinary path)."""
    import torch
    frames = [np.ascontiguousarray(pkg.synth.ring_sequence_frame(i, rings=48, azimuth_steps=4096)[:, :3], dtype=np.float64)
              for i in range(2)]
    frames.append(np.random.default_rng(3).uniform(-50.0, 50.0, (3000, 3)))
    tiny = np.random.default_rng(4).uniform(-1.0, 1.0, (30, 3))
    tiny[:21, 2] -= 10.0                                   # 30th percentile leaves 9 non-ground points: one cluster
    frames.append(tiny)
    for f in frames:
        want = pkg.pre.run(f, variant="B", host_arrays=False)
        want_pos = pkg.pre.people_positions(want)
        for cap in (8192, 3):
            got = pkg.pre.run_sequence_frame(torch.from_numpy(f).cuda(), centroid_cap=cap)
            cw, cg = want[pkg.pre.DEVICE_KEY], got[pkg.pre.DEVICE_KEY]
            assert got["dimensions"] == want["dimensions"] and sorted(got) == sorted(want)
            assert cg.n_clusters == cw.n_clusters and cg.guards == cw.guards
            assert torch.equal(cg.points, cw.points) and torch.equal(cg.clusters, cw.clusters)
            assert (cg.positions is not None) == (cg.n_clusters <= cap)
            got_pos = pkg.pre.people_positions(got)
            assert got_pos.shape == want_pos.shape and np.array_equal(got_pos, want_pos)


def test_frame_flow_step_equals_separate_entries(pkg):
    """`ops.frame_flow_step` (one copy in, `lidar_frame_flow`, one copy out) against the separate match / field entries fed
    with the numpy lattice of models/crowd_flow_model.py:107-111 -- host and device positions, bit for bit."""
    import torch
    rng = np.random.default_rng(11)
    prev = rng.uniform(-20.0, 20.0, (300, 2))
    cur = prev[rng.permutation(300)[:260]] + rng.normal(0.0, 0.05, (260, 2))
    x_grid, y_grid = np.arange(-21.5, 22.0, 1.0), np.arange(-20.25, 21.0, 1.0)
    X, Y = np.meshgrid(x_grid, y_grid)
    lattice = np.vstack([X.ravel(), Y.ravel()]).T
    match, vel, cur32 = pkg.ops.frame_flow_match(prev, cur, 0.1, 1.5)
    vec, mag = pkg.ops.frame_flow_field(lattice, cur32, match, vel, 3.0)
    want = (lattice, vec.cpu().numpy(), mag.cpu().numpy(), match.cpu().numpy(), vel.cpu().numpy())
    assert (want[3] >= 0).sum() > 200 and np.count_nonzero(want[2]) > 100
    for positions in (cur, torch.from_numpy(cur).cuda()):
        got = pkg.ops.frame_flow_step(prev, positions, x_grid, y_grid, 0.1, 1.5, 3.0)
        assert len(got) == 5
        for g, w in zip(got, want):
            assert g.dtype == w.dtype and np.array_equal(g, w)
    extra = torch.arange(7, dtype=torch.int64, device="cuda")
    got = pkg.ops.frame_flow_step(prev, cur, x_grid, y_grid, 0.1, 1.5, 3.0, extra=(extra,))
    assert np.array_equal(got[5], np.arange(7)) and np.array_equal(got[1], want[1])


def test_errors_are_python_exceptions(pkg):
    with pytest.raises(Exception):
        pkg.dp.preprocess_lidar_data(np.zeros((0, 3)))
    with pytest.raises(Exception):
        pkg.dp.load_lidar_data("/nonexistent/file.xyz")


def test_crowd_metrics_table_matches_ckdtree_join(pkg, processed_b):
    """plot_crowd_metrics' join (utils/visualization.py:295-323): nearest density cell per flow node, congestion risk and
    its 0-10 normalisation against scipy's cKDTree + the reference's pandas expressions.  Indices are compared where
    the nearest cell is unique (jittered nodes); on the reference's own lattice (every node exactly half-way between
    two centres) the nearest DISTANCE must agree and the chosen cell must be one of the equally near ones."""
    from scipy.spatial import cKDTree
    from lidar_ai_recommendation_software_b200.utils import visualization as viz
    pd_ = processed_b("crowd_20k")
    dres = pkg.CDM(grid_size=1.0).analyze(pd_)
    fres = pkg.CFM().analyze(pd_)
    dpts = np.column_stack(dres["grid_coordinates"])
    tree = cKDTree(dpts)
    # (1) unique nearest cells: jitter the lattice nodes off the half-way lines
    rng = np.random.default_rng(0)
    jit = {"flow_vectors": dict(fres["flow_vectors"])}
    jit["flow_vectors"]["positions"] = fres["flow_vectors"]["positions"] + rng.uniform(0.05, 0.45, fres["flow_vectors"]["positions"].shape)
    got = viz.crowd_metrics_table(dres, jit)
    dist, idx = tree.query(jit["flow_vectors"]["positions"], k=1)
    assert np.array_equal(got["nearest_index"], idx)
    assert np.allclose(got["nearest_distance"], dist, rtol=1e-15, atol=0)
    dens = dres["density_values"][idx]
    risk = dens / (jit["flow_vectors"]["magnitudes"] + 0.1)
    assert np.array_equal(got["density"], dens) and np.array_equal(got["congestion_risk"], risk)
    assert np.array_equal(got["congestion_risk_normalized"], risk / risk.max() * 10)
    assert np.array_equal(got["speed"], jit["flow_vectors"]["magnitudes"])
    # (2) the reference's own lattice: ties everywhere
    tie = viz.crowd_metrics_table(dres, fres)
    dist, idx = tree.query(fres["flow_vectors"]["positions"], k=1)
    assert np.allclose(tie["nearest_distance"], dist, rtol=1e-15, atol=0)
    chosen = dpts[tie["nearest_index"]]
    d_chosen = np.sqrt(((chosen - fres["flow_vectors"]["positions"]) ** 2).sum(1))
    assert np.allclose(d_chosen, dist, rtol=1e-12, atol=1e-12)
