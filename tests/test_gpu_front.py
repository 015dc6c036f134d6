"""The chained preprocess (`lidar_preprocess_front`: one enqueue, scalars stay on the device) against the
stage-by-stage ops it replaces and against numpy on the same cloud."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cloud(n, seed):
    from lidar_ai_recommendation_software_b200 import synth
    return np.ascontiguousarray(synth.add_outliers(synth.crowd_frame(n, seed=seed))[:, :3], dtype=np.float64)


@pytest.mark.parametrize("n,seed", [(14, 3), (5000, 0), (100_000, 1), (300_001, 2)])
def test_front_matches_staged_ops(n, seed):
    from lidar_ai_recommendation_software_b200 import ops, preprocess as pre
    pts = _cloud(n, seed)
    d = torch.from_numpy(pts).cuda()
    desc, inl, col, ng, idx, X = ops.preprocess_front(d, want_colors=True, scaler=True)

    bb = ops.bbox(d).cpu().numpy()
    np.testing.assert_array_equal(np.array(desc.bbox_raw)[[0, 1, 2, 4, 5, 6]], bb[[0, 1, 2, 4, 5, 6]])
    s1 = ops.moments(d).cpu().numpy()
    mean = s1[:3] / n
    std = np.sqrt(ops.moments(d, center=mean).cpu().numpy()[3:] / n)
    np.testing.assert_array_equal(np.array(desc.mean), mean)          # same kernel, same launch shape
    np.testing.assert_array_equal(np.array(desc.std), std)
    np.testing.assert_allclose(mean, pts.mean(axis=0), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(std, pts.std(axis=0), rtol=1e-12)

    zden = bb[6] - bb[2] + 1e-10
    inl2, col2, _, g2 = ops.sigma_filter(d, mean, 3 * std, 1e-9 * std, bb[2], zden)
    assert inl.shape == inl2.shape and int(desc.guard_sigma) == g2
    assert torch.equal(inl, inl2) and torch.equal(col, col2)
    mask = (np.abs(pts - pts.mean(axis=0)) < 3 * pts.std(axis=0)).all(axis=1)
    if g2 == 0:
        np.testing.assert_array_equal(inl.cpu().numpy(), pts[mask])

    n_in = inl.shape[0]
    z = inl[:, 2].cpu().numpy()
    lo = math.floor((n_in - 1) * 0.3)
    a, b = ops.select_kth(inl[:, 2], lo)
    assert (a, b) == (desc.kth[0], desc.kth[1])
    assert desc.z_thr == pre.percentile_from_order_stats(a, b, n_in, 30) == float(np.percentile(z, 30))

    ng2, idx2, sums2, _ = ops.ground_split(inl, desc.z_thr, mean)
    assert torch.equal(ng, ng2) and torch.equal(idx, idx2)
    np.testing.assert_allclose(np.array(desc.plane), sums2, rtol=1e-12, atol=1e-9)
    hin = inl.cpu().numpy()
    np.testing.assert_array_equal(np.array(desc.bbox_in), np.concatenate([hin.min(axis=0), hin.max(axis=0)]))
    m = ng.shape[0]
    if m:
        hng = ng.cpu().numpy()
        np.testing.assert_array_equal(np.array(desc.bbox_ng), np.concatenate([hng.min(axis=0), hng.max(axis=0)]))
        # StandardScaler + eps as preprocess_lidar_data derives them (utils/data_processing.py:190-196)
        from sklearn.preprocessing import StandardScaler
        Xs = StandardScaler().fit_transform(hng)
        np.testing.assert_allclose(X.cpu().numpy(), Xs, rtol=1e-9, atol=1e-9)
        eps = max(0.2, min(0.5, np.mean(np.std(Xs, axis=0)) * 0.5))
        assert desc.eps == pytest.approx(eps, rel=1e-12)
        # the transformed bbox of the non-ground points IS the bbox of the scaled copy
        lo3 = (np.array(desc.bbox_ng[:3]) - np.array(desc.sc_mean)) / np.array(desc.scale)
        hi3 = (np.array(desc.bbox_ng[3:]) - np.array(desc.sc_mean)) / np.array(desc.scale)
        hX = X.cpu().numpy()
        np.testing.assert_array_equal(lo3, hX.min(axis=0))
        np.testing.assert_array_equal(hi3, hX.max(axis=0))


def test_front_degenerate_clouds():
    from lidar_ai_recommendation_software_b200 import ops
    # every point identical: std = 0, nothing passes the strict 3-sigma test (the reference then fails in np.percentile)
    d = torch.zeros((100, 3), dtype=torch.float64, device="cuda")
    desc, inl, *_ = ops.preprocess_front(d, want_colors=False)
    assert int(desc.n_in) == 0 and inl.shape[0] == 0
    # one point
    d = torch.tensor([[1.0, 2.0, 3.0]], dtype=torch.float64, device="cuda")
    desc, inl, *_ = ops.preprocess_front(d, want_colors=False)
    assert int(desc.n_in) == 0


def test_scaler_treats_near_constant_columns_like_sklearn():
    """StandardScaler.fit marks a column constant when var <= n*eps*var + (n*mean*eps)^2
    (`_is_constant_feature`) — a scan line at UTM-sized x with sub-nanometre jitter gets scale 1, not 1/1e-10."""
    from sklearn.preprocessing import StandardScaler
    from lidar_ai_recommendation_software_b200 import ops
    rng = np.random.default_rng(11)
    n = 6000
    pts = np.column_stack([4.0e5 + rng.normal(0, 2e-11, n) * 0 + rng.integers(0, 2, n) * 5.8e-11,   # two adjacent doubles
                           rng.uniform(-20, 20, n), rng.uniform(0.0, 2.0, n)])
    d = torch.from_numpy(np.ascontiguousarray(pts)).cuda()
    desc, inl, col, ng, idx, X = ops.preprocess_front(d, want_colors=False, scaler=True)
    hng = ng.cpu().numpy()
    sk = StandardScaler().fit(hng)
    assert sk.scale_[0] == 1.0                                   # sklearn calls the column constant
    np.testing.assert_allclose(np.array(desc.scale), sk.scale_, rtol=1e-9)
    np.testing.assert_allclose(X.cpu().numpy(), sk.transform(hng), rtol=1e-9, atol=1e-9)


def test_people_positions_accepts_sparse_and_huge_cluster_ids():
    """a caller-built processed_data may carry any int64 cluster ids (np.unique in the reference,
    utils/data_processing.py:262-275): they are ranked first, not used to size the accumulators."""
    from lidar_ai_recommendation_software_b200.utils import data_processing as dp
    rng = np.random.default_rng(2)
    pts = rng.uniform(-5, 5, (4000, 3))
    ids = np.array([3, 17, 2**31 + 5, 2**40, 10**15], dtype=np.int64)
    lab = np.where(rng.uniform(size=4000) < 0.2, -1, ids[rng.integers(0, len(ids), 4000)]).astype(np.int64)
    got = dp.extract_people_positions({"points": pts, "clusters": lab})
    want = np.array([pts[lab == c].mean(axis=0)[:2] for c in np.unique(lab[lab >= 0])])
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)
