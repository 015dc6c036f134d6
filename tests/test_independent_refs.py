"""Independent restatements of the NEW ops, written against the published formulations with plain
PyTorch / numpy primitives and NO code shared with oracle/new_ops.py — so that the frozen oracle the CUDA path is tested
against is itself checked by something other than its author's first reading of the contract ("the NEW-op
oracle is self-authored").  CPU only."""
import numpy as np
import torch

from oracle import new_ops


def _cloud(b, n, seed):
    from lidar_ai_recommendation_software_b200 import synth
    return synth.sa_batch(b, n, seed=seed)


def fps_torch(xyz: torch.Tensor, m: int) -> torch.Tensor:
    """PointNet++ furthest point sampling as its CUDA op defines it: start at index 0, running minimum of the squared
    distance to the chosen set, next = first index of the maximum.  fp32, (dx*dx + dy*dy) + dz*dz."""
    b, n, _ = xyz.shape
    out = torch.zeros((b, m), dtype=torch.int64)
    for i in range(b):
        p = xyz[i]
        mind = torch.full((n,), 1e10, dtype=torch.float32)
        last = 0
        for j in range(1, m):
            d = p - p[last]
            d2 = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]
            mind = torch.minimum(mind, d2)
            last = int(torch.argmax(mind))            # first maximal index
            out[i, j] = last
    return out


def ball_query_torch(xyz: torch.Tensor, new_xyz: torch.Tensor, r: float, k: int) -> torch.Tensor:
    """First k indices (ascending) with d2 < r*r; the first hit pre-fills every slot; no hit leaves zeros."""
    b, n, _ = xyz.shape
    m = new_xyz.shape[1]
    out = torch.zeros((b, m, k), dtype=torch.int64)
    r2 = torch.tensor(r, dtype=torch.float32) * torch.tensor(r, dtype=torch.float32)
    ar = torch.arange(n)
    for i in range(b):
        d = xyz[i][None, :, :] - new_xyz[i][:, None, :]
        d2 = d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] + d[..., 2] * d[..., 2]
        hit = d2 < r2
        order = torch.where(hit, ar[None, :], torch.full((1, n), n)).sort(dim=1).values[:, :k]     # hits first, ascending
        first = order[:, :1]
        filled = torch.where(order < n, order, first.expand(-1, order.shape[1]))
        filled = torch.where(first < n, filled, torch.zeros_like(filled))
        out[i, :, :filled.shape[1]] = filled
        if filled.shape[1] < k:
            out[i, :, filled.shape[1]:] = filled[:, :1]
    return out


def test_fps_restated_with_torch_matches_the_oracle():
    clouds = _cloud(2, 1500, seed=5)
    want = new_ops.furthest_point_sample(clouds, 96)
    got = fps_torch(torch.from_numpy(clouds), 96).numpy()
    assert np.array_equal(got, want)
    # ties: duplicated points have equal distances, the lower index must win
    dup = np.concatenate([clouds[:, :300], clouds[:, :300]], axis=1)
    assert np.array_equal(fps_torch(torch.from_numpy(dup), 64).numpy(), new_ops.furthest_point_sample(dup, 64))


def test_ball_query_and_grouping_restated_with_torch_match_the_oracle():
    clouds = _cloud(2, 2048, seed=6)
    t = torch.from_numpy(clouds)
    idx = new_ops.furthest_point_sample(clouds, 64)
    centres = np.take_along_axis(clouds, idx[:, :, None].astype(np.int64), axis=1)
    for r, k in ((0.2, 32), (0.05, 8), (0.6, 16), (1e-4, 4)):
        want = new_ops.ball_query(clouds, centres, r, k)
        got = ball_query_torch(t, torch.from_numpy(centres), r, k).numpy()
        assert np.array_equal(got, want), (r, k)
    want_idx = new_ops.ball_query(clouds, centres, 0.2, 32)
    gi = torch.from_numpy(want_idx.astype(np.int64))
    grouped = torch.stack([t[b][gi[b]] for b in range(2)]) - torch.from_numpy(centres)[:, :, None, :]   # (B, M, k, 3)
    assert np.array_equal(grouped.permute(0, 3, 1, 2).numpy(), new_ops.group_points(clouds, None, want_idx, centres))


def test_shared_mlp_restated_with_conv1d_matches_the_oracle():
    from lidar_ai_recommendation_software_b200 import synth
    rng = np.random.default_rng(3)
    g = rng.normal(0, 0.3, (2, 3, 40, 32)).astype(np.float32)
    ws, bs = synth.sa_weights(seed=1)
    want = new_ops.shared_mlp_maxpool(g, ws, bs)
    x = torch.from_numpy(g).double().reshape(2, 3, -1)
    for w, b in zip(ws, bs):                                   # a 1x1 convolution IS the shared MLP layer
        x = torch.relu(torch.nn.functional.conv1d(x, torch.from_numpy(w).double()[:, :, None], torch.from_numpy(b).double()))
    got = x.reshape(2, -1, 40, 32).max(dim=-1).values.numpy()
    assert got.shape == want.shape and np.allclose(got, want, rtol=1e-10, atol=1e-12)


def test_voxel_downsample_restated_without_unique_matches_the_oracle():
    """Sort-based restatement (stable argsort of the keys, run boundaries, per-run sums): no np.unique, no np.add.at."""
    from lidar_ai_recommendation_software_b200 import synth
    for n, ext, voxel in ((5000, 6.0, 0.05), (20000, 3.0, 0.1), (1, 1.0, 0.05)):
        pts = synth.crowd_frame(n, seed=n % 5, extent=ext)
        want = new_ops.voxel_downsample(pts, voxel)
        p64 = pts.astype(np.float64)
        org = p64[:, :3].min(axis=0)
        ijk = np.floor((p64[:, :3] - org) / voxel).astype(np.int64)
        dims = ijk.max(axis=0) + 1
        key = (ijk[:, 0] * dims[1] + ijk[:, 1]) * dims[2] + ijk[:, 2]
        order = np.argsort(key, kind="stable")
        sk = key[order]
        head = np.concatenate([[True], sk[1:] != sk[:-1]])
        starts = np.flatnonzero(head)
        counts = np.diff(np.concatenate([starts, [n]]))
        rank_sorted = np.cumsum(head) - 1
        inverse = np.empty(n, dtype=np.int64)
        inverse[order] = rank_sorted
        cent = np.stack([[p64[order[s:s + c], k].sum() / c for k in range(4)] for s, c in zip(starts, counts)]).astype(np.float32)
        assert np.array_equal(key, want["voxel_key"]) and np.array_equal(inverse, want["inverse"])
        assert np.array_equal(sk[head], want["unique_keys"]) and np.array_equal(counts, want["counts"])
        assert np.allclose(cent, want["centroids"], rtol=1e-6, atol=1e-7)
