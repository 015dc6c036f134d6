"""GPU parity tests of the set-abstraction ops (NEW rows a16-a19) against the frozen oracle
definitions (oracle/new_ops.py): FPS / ball-query / grouping indices bit-exact, MLP rtol 1e-3."""
import numpy as np
import pytest
import torch

from oracle import new_ops

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pn2():
    from lidar_ai_recommendation_software_b200 import pointnet2
    return pointnet2


@pytest.fixture(scope="module")
def synth():
    from lidar_ai_recommendation_software_b200 import synth as s
    return s


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("b,n,m", [(1, 1, 1), (2, 33, 33), (3, 257, 64), (2, 2048, 128), (2, 2049, 100),
                                   (2, 5000, 256), (2, 16384, 512)])
def test_fps_bit_exact(pn2, b, n, m):
    rng = np.random.default_rng(n)
    xyz = rng.uniform(-1, 1, (b, n, 3)).astype(np.float32)
    if n > 40:
        xyz[:, 10:20] = xyz[:, :10]          # exact duplicates: ties must go to the lowest index
    want = new_ops.furthest_point_sample(xyz, m)
    got = pn2.furthest_point_sample(dev(xyz), m).cpu().numpy()
    assert np.array_equal(got, want)


def test_fps_many_clouds_waves_and_repeatability(pn2):
    """More clusters than fit at once (40 clouds x 8 CTAs: several waves), every cluster size (1, 2, 4, 8 CTAs),
    m == n (every point gets picked, the tail runs on all-zero distances), identical clouds in one batch, and the
    same launch repeated: the mbarrier / st.async exchange must neither hang nor depend on scheduling."""
    rng = np.random.default_rng(3)
    big = rng.uniform(-1, 1, (40, 16384, 3)).astype(np.float32)
    big[7] = big[3]
    got = pn2.furthest_point_sample(dev(big), 96).cpu().numpy()
    assert np.array_equal(got[7], got[3])
    for b in (0, 3, 39):
        assert np.array_equal(got[b], new_ops.furthest_point_sample(big[b:b + 1], 96)[0])
    for _ in range(3):
        assert np.array_equal(pn2.furthest_point_sample(dev(big), 96).cpu().numpy(), got)
    for n in (7, 2048, 2100, 4096, 4100, 8192, 9000):          # 1, 1, 2, 2, 4, 4, 8 CTAs per cloud
        xyz = rng.uniform(-1, 1, (3, n, 3)).astype(np.float32)
        m = n if n <= 2100 else 200
        assert np.array_equal(pn2.furthest_point_sample(dev(xyz), m).cpu().numpy(), new_ops.furthest_point_sample(xyz, m))


def test_fps_large_cloud_fallback(pn2):
    xyz = np.random.default_rng(0).normal(size=(1, 20000, 3)).astype(np.float32)
    want = new_ops.furthest_point_sample(xyz, 64)
    assert np.array_equal(pn2.furthest_point_sample(dev(xyz), 64).cpu().numpy(), want)


@pytest.mark.parametrize("b,n,m,k,r", [(2, 256, 32, 8, 0.4), (2, 1000, 100, 16, 0.3), (1, 1001, 33, 32, 0.25),
                                       (2, 4096, 256, 32, 0.2), (1, 4100, 64, 64, 0.15)])
def test_ball_query_and_grouping_bit_exact(pn2, b, n, m, k, r):
    rng = np.random.default_rng(m)
    xyz = rng.uniform(-1, 1, (b, n, 3)).astype(np.float32)
    fidx = new_ops.furthest_point_sample(xyz, m)
    new_xyz = np.stack([xyz[i][fidx[i]] for i in range(b)])
    new_xyz[:, -1] = 10.0                     # a centre with no neighbour at all -> zeros
    d_xyz, d_new = dev(xyz), dev(new_xyz)
    assert np.array_equal(pn2.gather_points(d_xyz, dev(fidx)).cpu().numpy()[:, :-1], new_xyz[:, :-1])
    want = new_ops.ball_query(xyz, new_xyz, r, k)
    got = pn2.ball_query(d_xyz, d_new, r, k)
    assert np.array_equal(got.cpu().numpy(), want)
    feats = rng.normal(size=(b, 5, n)).astype(np.float32)
    g_want = new_ops.group_points(xyz, feats, want, new_xyz)
    g_got = pn2.group_points(d_xyz, dev(feats), got, d_new).cpu().numpy()
    assert np.array_equal(g_got, g_want)
    assert np.array_equal(pn2.group_points(d_xyz, None, got, d_new).cpu().numpy(), g_want[:, :3])


def _mlp_case(synth, b, n, m, k, r, seed=1, c_feat=0, widths=(64, 64, 128)):
    xyz = synth.sa_batch(b, n, seed=seed)
    fidx = new_ops.furthest_point_sample(xyz, m)
    new_xyz = np.stack([xyz[i][fidx[i]] for i in range(b)])
    idx = new_ops.ball_query(xyz, new_xyz, r, k)
    ws, bs = synth.sa_weights(seed=seed, c_in=3 + c_feat, widths=widths)
    return xyz, new_xyz, idx, ws, bs


@pytest.mark.parametrize("widths,c_feat,k", [((64, 64, 128), 0, 32), ((32, 48, 64), 4, 16), ((16, 16, 16), 0, 5)])
def test_shared_mlp_simt_matches_oracle(pn2, synth, widths, c_feat, k):
    b, n, m = 2, 2048, 96
    xyz, new_xyz, idx, ws, bs = _mlp_case(synth, b, n, m, k, 0.2, c_feat=c_feat, widths=widths)
    feats = np.random.default_rng(3).normal(size=(b, c_feat, n)).astype(np.float32) if c_feat else None
    grouped = new_ops.group_points(xyz, feats, idx, new_xyz)
    want = new_ops.shared_mlp_maxpool(grouped, ws, bs)
    dw, db = [dev(w) for w in ws], [dev(x) for x in bs]
    fused = pn2.shared_mlp_maxpool(dw, db, xyz=dev(xyz), idx=dev(idx), new_xyz=dev(new_xyz),
                                   features=dev(feats) if c_feat else None, impl=pn2.MLP_SIMT).cpu().numpy()
    assert np.allclose(fused, want, rtol=1e-3, atol=1e-5)
    mat = pn2.shared_mlp_maxpool(dw, db, grouped=dev(grouped), impl=pn2.MLP_SIMT).cpu().numpy()
    assert np.allclose(mat, want, rtol=1e-3, atol=1e-5)


def test_shared_mlp_tcgen05_matches_oracle(pn2, synth):
    """tensor-core path (bf16 hi/lo split, fp32 accumulate in TMEM) within rtol 1e-3 of the fp64 oracle."""
    b, n, m, k = 2, 4096, 250, 32           # 500 centres: not a multiple of 4 per CTA tile boundary on purpose
    xyz, new_xyz, idx, ws, bs = _mlp_case(synth, b, n, m, k, 0.2)
    grouped = new_ops.group_points(xyz, None, idx, new_xyz)
    want = new_ops.shared_mlp_maxpool(grouped, ws, bs)
    dw, db = [dev(w) for w in ws], [dev(x) for x in bs]
    got = pn2.shared_mlp_maxpool(dw, db, xyz=dev(xyz), idx=dev(idx), new_xyz=dev(new_xyz), impl=pn2.MLP_TCGEN05)
    got = got.cpu().numpy()
    assert np.isfinite(got).all()
    assert np.allclose(got, want, rtol=1e-3, atol=1e-5), float(np.abs(got - want).max())
    simt = pn2.shared_mlp_maxpool(dw, db, xyz=dev(xyz), idx=dev(idx), new_xyz=dev(new_xyz), impl=pn2.MLP_SIMT)
    assert np.allclose(got, simt.cpu().numpy(), rtol=2e-4, atol=2e-5)


def test_set_abstraction_level_end_to_end(pn2, synth):
    b, n, m, k, r = 2, 4096, 128, 32, 0.2
    xyz = synth.sa_batch(b, n, seed=5)
    ws, bs = synth.sa_weights(seed=2)
    sa = pn2.SetAbstraction(m, r, k, [dev(w) for w in ws], [dev(x) for x in bs])
    new_xyz, feats, fidx, idx = sa(dev(xyz))
    w_fidx = new_ops.furthest_point_sample(xyz, m)
    assert np.array_equal(fidx.cpu().numpy(), w_fidx)
    w_new = np.stack([xyz[i][w_fidx[i]] for i in range(b)])
    w_idx = new_ops.ball_query(xyz, w_new, r, k)
    assert np.array_equal(idx.cpu().numpy(), w_idx)
    want = new_ops.shared_mlp_maxpool(new_ops.group_points(xyz, None, w_idx, w_new), ws, bs)
    assert np.allclose(feats.cpu().numpy(), want, rtol=1e-3, atol=1e-5)
