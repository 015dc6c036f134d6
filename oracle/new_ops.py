"""Frozen CPU definitions of the ops the north star names but the reference lacks (NEW rows).

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED: the reference has no voxel
downsample, ROI crop, frame-to-frame flow, FPS, ball query, grouping or shared MLP,
so there is nothing upstream to pin these against.  The contracts are SURVEY.md Appendix B; FPS /
ball query / grouping follow the published PointNet++ CUDA-op semantics (Qi et al. 2017,
`pointnet2_ops`: furthest_point_sample, ball_query, group_points).
"""
from __future__ import annotations

import numpy as np


# B.1 ------------------------------------------------------------------------------------------
def voxel_downsample(points: np.ndarray, voxel: float, origin=None):
    """points (n,4) float32 -> dict(voxel_key, inverse, centroids, counts, unique_keys, dims, origin).

    i = floor((f64(p) - origin) / voxel); key = (ix*Dy + iy)*Dz + iz; voxels in ascending key
    order; centroid = fp32 of the fp64 mean of the members summed in ascending original index.
    """
    pts = np.asarray(points, dtype=np.float32)
    p64 = pts.astype(np.float64)
    n = p64.shape[0]
    if n == 0:
        z = np.zeros(0, dtype=np.int64)
        return {"voxel_key": z, "inverse": z, "centroids": np.zeros((0, 4), np.float32),
                "counts": z, "unique_keys": z, "dims": (1, 1, 1), "origin": (0.0, 0.0, 0.0)}
    org = p64[:, :3].min(0) if origin is None else np.asarray(origin, dtype=np.float64)
    ijk = np.floor((p64[:, :3] - org) / np.float64(voxel)).astype(np.int64)
    dims = ijk.max(0) + 1
    key = (ijk[:, 0] * dims[1] + ijk[:, 1]) * dims[2] + ijk[:, 2]
    uniq, inverse, counts = np.unique(key, return_inverse=True, return_counts=True)
    sums = np.zeros((len(uniq), 4), dtype=np.float64)
    np.add.at(sums, inverse, p64)  # unbuffered: sequential in ascending original index
    cent = (sums / counts[:, None]).astype(np.float32)
    return {"voxel_key": key, "inverse": inverse.astype(np.int64), "centroids": cent,
            "counts": counts.astype(np.int64), "unique_keys": uniq, "dims": tuple(int(d) for d in dims),
            "origin": tuple(float(o) for o in org)}


# B.2 ------------------------------------------------------------------------------------------
def roi_crop(points: np.ndarray, lo, hi):
    """keep <=> lo <= p <= hi on x,y,z.  float32 clouds compare in fp32 (bounds cast to fp32),
    float64 clouds in fp64.  Returns (cropped, mask)."""
    pts = np.asarray(points)
    dt = pts.dtype
    lo = np.asarray(lo, dtype=np.float64).astype(dt)
    hi = np.asarray(hi, dtype=np.float64).astype(dt)
    m = np.all((pts[:, :3] >= lo) & (pts[:, :3] <= hi), axis=1)
    return pts[m], m


# B.3 ------------------------------------------------------------------------------------------
def frame_flow_match(prev: np.ndarray, cur: np.ndarray, dt: float, gate: float = 1.5):
    """Nearest previous centroid per current centroid (fp32 squared distance, lowest index on ties),
    matched iff d <= gate.  Returns (match int32 (-1 = none), velocity (C2,2) float32)."""
    prev = np.asarray(prev, dtype=np.float32)
    cur = np.asarray(cur, dtype=np.float32)
    c2 = cur.shape[0]
    match = np.full(c2, -1, dtype=np.int32)
    vel = np.zeros((c2, 2), dtype=np.float32)
    if prev.shape[0] == 0:
        return match, vel
    g2 = np.float32(gate) * np.float32(gate)
    for j in range(c2):
        dx = cur[j, 0] - prev[:, 0]
        dy = cur[j, 1] - prev[:, 1]
        d2 = dx * dx + dy * dy          # fp32, (dx*dx) + (dy*dy), no FMA
        i = int(np.argmin(d2))          # first minimum = lowest index
        if d2[i] <= g2:
            match[j] = i
            vel[j, 0] = (cur[j, 0] - prev[i, 0]) / np.float32(dt)
            vel[j, 1] = (cur[j, 1] - prev[i, 1]) / np.float32(dt)
    return match, vel


def frame_flow_field(lattice: np.ndarray, cur: np.ndarray, match: np.ndarray, vel: np.ndarray,
                     radius: float = 3.0):
    """Each lattice node takes the mean velocity of the matched people within `radius` (inclusive,
    fp64 distances on the widened fp32 centroids), else 0.  Returns (vectors (G,2) f64, magnitudes)."""
    lattice = np.asarray(lattice, dtype=np.float64)
    cur64 = np.asarray(cur, dtype=np.float32).astype(np.float64)
    v64 = np.asarray(vel, dtype=np.float32).astype(np.float64)
    ok = np.asarray(match) >= 0
    G = lattice.shape[0]
    vec = np.zeros((G, 2))
    r2 = float(radius) * float(radius)
    if ok.any():
        pc, pv = cur64[ok], v64[ok]
        for s in range(0, G, 2048):
            L = lattice[s:s + 2048]
            dx = L[:, 0, None] - pc[None, :, 0]
            dy = L[:, 1, None] - pc[None, :, 1]
            inside = (dx * dx + dy * dy) <= r2
            cnt = inside.sum(1)
            sx = (inside * pv[None, :, 0]).sum(1)
            sy = (inside * pv[None, :, 1]).sum(1)
            nz = cnt > 0
            vec[s:s + 2048][nz, 0] = sx[nz] / cnt[nz]
            vec[s:s + 2048][nz, 1] = sy[nz] / cnt[nz]
    mag = np.sqrt(vec[:, 0] ** 2 + vec[:, 1] ** 2)
    return vec, mag


# B.4 ------------------------------------------------------------------------------------------
def furthest_point_sample(xyz: np.ndarray, m: int) -> np.ndarray:
    """(B,N,3) float32 -> (B,m) int32.  idx[0] = 0; running min of fp32 squared distances
    ((dx*dx + dy*dy) + dz*dz, no FMA) initialised to 1e10; next = lowest index of the maximum."""
    xyz = np.asarray(xyz, dtype=np.float32)
    B, N, _ = xyz.shape
    out = np.zeros((B, m), dtype=np.int32)
    for b in range(B):
        p = xyz[b]
        mind = np.full(N, np.float32(1e10), dtype=np.float32)
        cur = 0
        for k in range(1, m):
            d = p - p[cur]
            d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
            mind = np.minimum(mind, d2)
            cur = int(np.argmax(mind))
            out[b, k] = cur
    return out


# B.5 ------------------------------------------------------------------------------------------
def ball_query(xyz: np.ndarray, new_xyz: np.ndarray, radius: float, k: int) -> np.ndarray:
    """(B,N,3),(B,M,3) -> (B,M,k) int32: first k indices (ascending) with d² < r² (strict, fp32, same
    expression as FPS); the first hit pre-fills all k slots; no hit at all leaves zeros."""
    xyz = np.asarray(xyz, dtype=np.float32)
    new_xyz = np.asarray(new_xyz, dtype=np.float32)
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    r2 = np.float32(radius) * np.float32(radius)
    out = np.zeros((B, M, k), dtype=np.int32)
    for b in range(B):
        p = xyz[b]
        for s in range(0, M, 256):
            c = new_xyz[b, s:s + 256]
            dx = c[:, None, 0] - p[None, :, 0]
            dy = c[:, None, 1] - p[None, :, 1]
            dz = c[:, None, 2] - p[None, :, 2]
            d2 = (dx * dx + dy * dy) + dz * dz
            hit = d2 < r2
            for j in range(c.shape[0]):
                idx = np.flatnonzero(hit[j])[:k]
                if idx.size:
                    out[b, s + j, :] = idx[0]
                    out[b, s + j, :idx.size] = idx
    return out


# B.6 ------------------------------------------------------------------------------------------
def group_points(xyz: np.ndarray, feats, idx: np.ndarray, new_xyz: np.ndarray) -> np.ndarray:
    """out[b,:3,m,j] = xyz[b, idx[b,m,j]] - new_xyz[b,m]; feature channels gathered unchanged after."""
    xyz = np.asarray(xyz, dtype=np.float32)
    B = xyz.shape[0]
    g = np.stack([xyz[b][idx[b]] for b in range(B)], 0)              # (B,M,k,3)
    g = g - np.asarray(new_xyz, dtype=np.float32)[:, :, None, :]
    out = np.transpose(g, (0, 3, 1, 2))
    if feats is not None:
        feats = np.asarray(feats, dtype=np.float32)                   # (B,C,N)
        gf = np.stack([feats[b][:, idx[b]] for b in range(B)], 0)    # (B,C,M,k)
        out = np.concatenate([out, gf], axis=1)
    return np.ascontiguousarray(out, dtype=np.float32)


# B.7 ------------------------------------------------------------------------------------------
def shared_mlp_maxpool(grouped: np.ndarray, weights, biases) -> np.ndarray:
    """(B,C,M,k) float32 -> (B,C_out,M): 1x1 conv + bias + ReLU per layer (BatchNorm folded), max over
    k.  fp64 matmuls; the CUDA path must agree within rtol 1e-3 (+ atol 1e-5)."""
    h = np.asarray(grouped, dtype=np.float64)
    for W, b in zip(weights, biases):
        h = np.einsum("oc,bcmk->bomk", np.asarray(W, np.float64), h) + np.asarray(b, np.float64)[None, :, None, None]
        h = np.maximum(h, 0.0)
    return h.max(axis=3)
