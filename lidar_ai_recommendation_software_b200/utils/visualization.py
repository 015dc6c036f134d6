"""Compute cores of the reference's utils/visualization.py, CUDA-backed.

The reference functions build plotly figures; what they *compute* per point is re-exported here under the
same roles, so the figure code can keep consuming numpy arrays:

    local_point_density(points, r)        KDTree(points).query_radius(points, r, count_only=True)
                                          utils/visualization.py:43-45 (3-D), :164-168 (2-D projection);
                                          app_simplified.py:158-159
    distance_from_center(points)          utils/visualization.py:50-54
    projection_histogram(processed, ...)  np.histogram2d of a projection, utils/visualization.py:116-137
    crowd_metrics_table(density, flow)    the join behind plot_crowd_metrics, utils/visualization.py:295-323: nearest
                                          density cell per flow node (cKDTree.query), congestion risk, its 0-10 scale

No CPU fallback: every per-point computation runs in the sm_100a core.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops
from .. import preprocess as _pre

_DIM = {"x": 0, "y": 1, "z": 2}


def _device_points(points) -> torch.Tensor:
    dev = ops.require_cuda()
    if isinstance(points, torch.Tensor):
        return points.to(device=dev, dtype=torch.float64).contiguous()
    pts = np.asarray(points, dtype=np.float64)
    if pts.ndim != 2 or pts.shape[1] not in (2, 3):
        raise ValueError(f"expected an (n,2) or (n,3) point array, got shape {pts.shape}")
    return torch.from_numpy(np.ascontiguousarray(pts)).to(dev)


def local_point_density(points, r=0.5):
    """Number of points within `r` of every point (itself included), int64 (n,) like
    `KDTree(points).query_radius(points, r=r, count_only=True)`.  (n,3) or (n,2) input."""
    d = _device_points(points)
    n = d.shape[0]
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    if d.shape[1] == 2:
        d = torch.cat([d, torch.zeros((n, 1), dtype=torch.float64, device=d.device)], 1).contiguous()
    return ops.ball_count(d, float(r)).cpu().numpy()


def distance_from_center(points):
    """utils/visualization.py:50-54: Euclidean distance of every point from the centroid (np.mean), float64 (n,).
    Device path: the moments kernel gives the centroid, one per-point kernel the distances (`lidar_centroid_distances`)."""
    d = _device_points(points)
    if d.shape[1] != 3:
        raise ValueError("distance_from_center expects (n,3) points")
    return ops.centroid_distances(d).cpu().numpy()


def projection_histogram(processed_data, projection_dims=("x", "y"), resolution=100):
    """utils/visualization.py:116-137: (heatmap transposed for display, x_centers, y_centers) of the points
    projected on two axes, np.histogram2d(bins=resolution, range=dimension ranges) semantics, integer counts."""
    pts, _ = _pre.device_view(processed_data)
    a, b = _DIM[projection_dims[0]], _DIM[projection_dims[1]]
    r0 = processed_data["dimensions"][f"{projection_dims[0]}_range"]
    r1 = processed_data["dimensions"][f"{projection_dims[1]}_range"]
    x_edges = np.linspace(r0[0], r0[1], resolution + 1)
    y_edges = np.linspace(r1[0], r1[1], resolution + 1)
    counts = ops.hist2d_counts(pts[:, a], pts[:, b], x_edges, y_edges)
    heat = counts.cpu().numpy().astype(np.float64).T
    return heat, (x_edges[:-1] + x_edges[1:]) / 2, (y_edges[:-1] + y_edges[1:]) / 2


def crowd_metrics_table(density_results, flow_results):
    """The table `plot_crowd_metrics` draws (utils/visualization.py:295-323), computed on the device: for every flow
    lattice node its x, y, speed, the density of the NEAREST density cell (`cKDTree(density_points).query(flow_points,
    k=1)`), congestion_risk = density / (speed + 0.1) and congestion_risk_normalized = risk / max(risk) * 10.

    Returns a dict of float64 arrays (plus `nearest_index` int64 and `nearest_distance`), column for column what the
    reference's `combined_df` holds.  The density cells form the rectilinear grid of CrowdDensityModel.analyze
    (`grid_coordinates` = repeat(grid_x, ny), tile(grid_y, nx)), so the neighbour search is per axis.  Tie note: with the
    reference's own grids (1 m cells whose centres sit at x_min - 1.5 + k, flow nodes at x_min + i) every node is
    EXACTLY half-way between two centres per axis, and the winner is decided by the last bit of the subtraction and, on
    exact ties, by cKDTree's traversal order; here an exact tie goes to the lowest flat index.  The nearest DISTANCE is
    identical in all cases; the joined density can differ from scipy's only where two cells are equally near.
    The two `griddata(..., method='linear')` resamplings of the figure (:224, :354) stay with scipy: on a lattice the
    Delaunay triangulation is degenerate (four co-circular points per cell) and the interpolant depends on which
    diagonal Qhull happens to pick."""
    fx, fy = density_results["grid_coordinates"]
    fx, fy = np.asarray(fx, dtype=np.float64), np.asarray(fy, dtype=np.float64)
    dens = np.asarray(density_results["density_values"], dtype=np.float64).reshape(-1)
    pos = np.asarray(flow_results["flow_vectors"]["positions"], dtype=np.float64).reshape(-1, 2)
    speed = np.asarray(flow_results["flow_vectors"]["magnitudes"], dtype=np.float64).reshape(-1)
    if dens.size == 0 or pos.shape[0] == 0:
        raise ValueError("crowd_metrics_table needs a non-empty density grid and flow lattice")
    # grid lines back from the flattened coordinates: flat_x = repeat(grid_x, ny), flat_y = tile(grid_y, nx)
    ny = int(np.argmax(fx != fx[0])) if np.any(fx != fx[0]) else fx.size
    nx = fx.size // ny
    gx, gy = np.ascontiguousarray(fx[::ny]), np.ascontiguousarray(fy[:ny])
    if nx * ny != fx.size or not np.array_equal(np.repeat(gx, ny), fx) or not np.array_equal(np.tile(gy, nx), fy):
        raise ValueError("grid_coordinates are not the repeat / tile grid of CrowdDensityModel.analyze")
    index, dist, at, risk, rmax = ops.nearest_grid_cell(pos, gx, gy, dens, speed)
    h_index, h_dist, h_at, h_risk, h_max = ops.fetch("crowd_metrics", index, dist, at, risk, rmax)
    risk = h_risk.copy()
    return {"x": pos[:, 0].copy(), "y": pos[:, 1].copy(), "speed": speed.copy(), "density": h_at.copy(),
            "congestion_risk": risk, "congestion_risk_normalized": risk / float(h_max[0]) * 10,
            "nearest_index": h_index.copy(), "nearest_distance": h_dist.copy()}
