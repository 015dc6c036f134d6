"""Seeded synthetic LiDAR inputs.  All clouds are fp32 `float4`
(x, y, z, intensity); the CPU oracle / reference sees `xyz.astype(np.float64)` — the exact widening.

The reference ships no data files (its .gitignore excludes them); its only reproducible input is the
inline 10 000-point sample of app_simplified.py:994-1024, restated in `reference_sample()`.
"""
from __future__ import annotations

import numpy as np


def crowd_frame(n: int, seed: int = 0, extent: float = 50.0, people_frac: float = 0.6,
                extent_y: float | None = None) -> np.ndarray:
    """C.1 crowd frame: undulating ground + vertical person clusters, shuffled. (n,4) float32."""
    rng = np.random.default_rng(seed)
    ex, ey = float(extent), float(extent if extent_y is None else extent_y)
    n_people_pts = int(n * people_frac)
    n_ground = n - n_people_pts
    n_people = max(1, int(n * people_frac / 400))
    gx = rng.uniform(-ex, ex, n_ground)
    gy = rng.uniform(-ey, ey, n_ground)
    gz = 0.1 * np.sin(0.5 * gx) * np.cos(0.5 * gy) + rng.normal(0.0, 0.01, n_ground)
    centres = np.stack([rng.uniform(-0.9 * ex, 0.9 * ex, n_people), rng.uniform(-0.9 * ey, 0.9 * ey, n_people)], 1)
    pid = rng.integers(0, n_people, n_people_pts)
    px = centres[pid, 0] + rng.normal(0.0, 0.12, n_people_pts)
    py = centres[pid, 1] + rng.normal(0.0, 0.12, n_people_pts)
    pz = rng.uniform(0.1, 1.8, n_people_pts)
    xyz = np.concatenate([np.stack([gx, gy, gz], 1), np.stack([px, py, pz], 1)], 0)
    inten = rng.uniform(0.0, 1.0, n)
    perm = rng.permutation(n)
    out = np.empty((n, 4), dtype=np.float32)
    out[:, :3] = xyz[perm]
    out[:, 3] = inten
    return out


def add_outliers(points: np.ndarray, every: int = 400, dz: float = 25.0) -> np.ndarray:
    """Lift every `every`-th point by `dz` metres so the 3-sigma filter has something to remove."""
    out = np.array(points, copy=True)
    out[::every, 2] += dz
    return out


def venue_scan_shard(n: int, seed: int, shard: int, n_shards: int) -> np.ndarray:
    """C.4 venue scan (400 m x 300 m, people_frac 0.4), generated shard by shard so that no host
    ever materialises the whole 50 M-point scan.  Every shard draws from its own seeded stream."""
    per = n // n_shards
    m = per + (n - per * n_shards if shard == n_shards - 1 else 0)
    return crowd_frame(m, seed=seed * 1000 + shard, extent=200.0, extent_y=150.0, people_frac=0.4)


def sa_batch(batch: int = 16, n: int = 16384, seed: int = 0) -> np.ndarray:
    """C.2 set-abstraction batch: `batch` clouds of n points normalised to the unit sphere. (B,n,3) f32."""
    out = np.empty((batch, n, 3), dtype=np.float32)
    for b in range(batch):
        pts = crowd_frame(n, seed=seed * 100 + b, extent=4.0)[:, :3].astype(np.float64)
        pts -= pts.mean(0)
        pts /= np.sqrt((pts ** 2).sum(1)).max()
        out[b] = pts.astype(np.float32)
    return out


def sa_weights(seed: int = 1, c_in: int = 3, widths=(64, 64, 128)):
    """Shared-MLP weights W ~ N(0, 1/sqrt(fan_in)), small biases; BatchNorm(eval) pre-folded."""
    rng = np.random.default_rng(seed)
    ws, bs = [], []
    fan_in = c_in
    for w in widths:
        ws.append((rng.normal(0.0, 1.0, (w, fan_in)) / np.sqrt(fan_in)).astype(np.float32))
        bs.append(rng.normal(0.0, 0.1, w).astype(np.float32))
        fan_in = w
    return ws, bs


def reference_sample() -> np.ndarray:
    """The reference's own demo cloud (app_simplified.py:994-1024), legacy np.random.seed(42).
    (10000,3) float64.  NOTE: seeds the global legacy RNG exactly like the app does."""
    n_points = 10000
    np.random.seed(42)
    x = np.random.uniform(-15, 15, n_points)
    y = np.random.uniform(-15, 15, n_points)
    z = np.zeros(n_points)
    z += 0.1 * np.sin(x * 0.5) * np.cos(y * 0.5)
    people = np.random.uniform(-10, 10, (50, 2))
    for i in range(n_points):
        d = np.sqrt((x[i] - people[:, 0]) ** 2 + (y[i] - people[:, 1]) ** 2)
        if np.min(d) < 0.3:
            z[i] = np.random.uniform(0.1, 1.8)
    return np.column_stack((x, y, z))


def tiny_cloud() -> np.ndarray:
    """14 points: both degenerate guards of the reference's preprocess fire (<= 10 ground, <= 10 non-ground)."""
    return np.random.default_rng(4).uniform(-2, 2, (14, 3))


def sparse_cloud() -> np.ndarray:
    """300 scattered returns on a 60 m x 60 m x 3 m volume: no DBSCAN cluster at eps = 0.3 m."""
    r = np.random.default_rng(9)
    return np.column_stack([r.uniform(-30, 30, 300), r.uniform(-30, 30, 300), r.uniform(0, 3, 300)])


def ring_sequence_frame(frame: int, rings: int = 128, azimuth_steps: int = 20480, seed: int = 0,
                        extent: float = 50.0, n_people: int = 600, dt: float = 0.1) -> np.ndarray:
    """C.3 128-beam frame (~2.6 M returns), scan ordered (ring-major, azimuth-minor).

    Sensor at (0,0,3 m); ring elevations linspace(-25 deg, +15 deg).  Rays hit the C.1 ground (flat
    approximation z = 0.1 sin(.5x)cos(.5y) evaluated at the flat-ground hit point) or the nearest
    vertical person cylinder (r = 0.25 m, h = 1.7 m) found on a coarse azimuth sweep; returns past
    120 m are dropped.  People drift toward the exit at the centre of the +x edge by 1 m/s * dt per
    frame, so consecutive frames carry a real displacement signal for frame_flow.
    """
    rng = np.random.default_rng(seed)
    centres = np.stack([rng.uniform(-0.9 * extent, 0.9 * extent, n_people),
                        rng.uniform(-0.9 * extent, 0.9 * extent, n_people)], 1)
    to_exit = np.array([extent, 0.0]) - centres
    dist = np.linalg.norm(to_exit, axis=1, keepdims=True)
    centres = centres + to_exit / np.maximum(dist, 1e-9) * (1.0 * dt * frame)
    elev = np.deg2rad(np.linspace(-25.0, 15.0, rings))
    az = np.linspace(-np.pi, np.pi, azimuth_steps, endpoint=False)
    h = 3.0
    # ground hit range per ring (only downward rings hit the ground)
    pts = []
    # person lookup by azimuth bucket: nearest cylinder per azimuth step
    p_az = np.arctan2(centres[:, 1], centres[:, 0])
    p_rng = np.linalg.norm(centres, axis=1)
    half_w = np.arctan2(0.25, np.maximum(p_rng, 0.3))
    nearest = np.full(azimuth_steps, np.inf)
    step = 2 * np.pi / azimuth_steps
    for a, r, w in zip(p_az, p_rng, half_w):
        lo = int(np.floor((a - w + np.pi) / step))
        hi = int(np.ceil((a + w + np.pi) / step))
        idx = np.arange(lo, hi + 1) % azimuth_steps
        nearest[idx] = np.minimum(nearest[idx], r)
    frng = np.random.default_rng(seed * 100003 + frame)
    with np.errstate(invalid="ignore"):   # rays that hit nothing carry inf ranges until they are masked out
        for e in elev:
            ce, se = np.cos(e), np.sin(e)
            ground_r = h / np.tan(-e) if e < -1e-6 else np.inf          # horizontal range of ground hit
            hr = np.full(azimuth_steps, ground_r)
            # person hit if the ray is between z=0 and z=1.7 at the person's range
            z_at_person = h + nearest * np.tan(e)
            hit = (nearest < hr) & (z_at_person >= 0.0) & (z_at_person <= 1.7)
            hr = np.where(hit, nearest, hr)
            ok = np.isfinite(hr) & (hr / max(ce, 1e-9) <= 120.0)
            x = hr * np.cos(az)
            y = hr * np.sin(az)
            z = np.where(hit, z_at_person, 0.1 * np.sin(0.5 * x) * np.cos(0.5 * y))
            x, y, z = x[ok], y[ok], z[ok]
            noise = frng.normal(0.0, 0.01, (x.size, 3))
            inten = frng.uniform(0.0, 1.0, x.size)
            pts.append(np.column_stack([x + noise[:, 0], y + noise[:, 1], z + noise[:, 2], inten]))
    return np.concatenate(pts, 0).astype(np.float32)

