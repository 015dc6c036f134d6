"""Compute cores of the reference's utils/visualization.py, CUDA-backed.

The reference functions build plotly figures; what they *compute* per point is re-exported here under the
same roles, so the figure code can keep consuming numpy arrays:

    local_point_density(points, r)        KDTree(points).query_radius(points, r, count_only=True)
                                          utils/visualization.py:43-45 (3-D), :164-168 (2-D projection);
                                          app_simplified.py:158-159
    distance_from_center(points)          utils/visualization.py:50-54
    projection_histogram(processed, ...)  np.histogram2d of a projection, utils/visualization.py:116-137

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
