"""Drop-in for the reference's utils/data_processing.py — same names, argument order, defaults,
result-dict keys, dtypes and error behaviour; every per-point computation runs in the sm_100a CUDA
core (no CPU fallback).

    from lidar_ai_recommendation_software_b200.utils.data_processing import (
        load_lidar_data, preprocess_lidar_data, downsample_point_cloud,
        extract_people_positions, calculate_grid_density)
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops
from .. import preprocess as _pre
from ..io import load_lidar_data  # noqa: F401  (re-exported: app.py:7 imports it from this module)

__all__ = ["load_lidar_data", "preprocess_lidar_data", "downsample_point_cloud", "extract_people_positions",
           "calculate_grid_density"]


def preprocess_lidar_data(points):
    """utils/data_processing.py:127-229.  Returns the processed_data dict:
    points (n',3) f64, colors (n',3) f64, normals (n',3) f64, clusters (n',) int64, ground_plane (4,),
    dimensions {x_range, y_range, z_range, width, length, height}."""
    return _pre.run(points, variant="A")


def downsample_point_cloud(points, factor=0.1):
    """utils/data_processing.py:231-249: keep a random `factor` of the points (unseeded global numpy RNG,
    exactly like the reference — the draw is host-side parameter generation), gather on the device."""
    if factor >= 1.0:
        return points
    num_points = len(points)
    num_keep = max(1, int(num_points * factor))
    indices = np.random.choice(num_points, num_keep, replace=False)
    return ops.gather_rows(points, indices)


def extract_people_positions(processed_data):
    """utils/data_processing.py:251-280: (C,2) float64 centroid xy per cluster id >= 0, ascending id."""
    return _pre.people_positions(processed_data)


def calculate_grid_density(people_positions, x_range, y_range, grid_size=1.0):
    """utils/data_processing.py:282-328: (grid_x, grid_y, density[x][y]) with margin 2g, np.arange
    edges and np.histogram2d binning; counts are bit-exact, density = counts / g².
    Returns (None, None, None) for an empty input."""
    if len(people_positions) == 0:
        return None, None, None
    x_min, x_max = x_range
    y_min, y_max = y_range
    margin = grid_size * 2
    x_min -= margin
    x_max += margin
    y_min -= margin
    y_max += margin
    x_edges = np.arange(x_min, x_max + grid_size, grid_size)
    y_edges = np.arange(y_min, y_max + grid_size, grid_size)
    dev = ops.require_cuda()
    if isinstance(people_positions, torch.Tensor):
        pos = people_positions.to(device=dev, dtype=torch.float64)
    else:
        pos = torch.from_numpy(np.asarray(people_positions, dtype=np.float64)).to(dev)
    counts = ops.hist2d_counts(pos[:, 0], pos[:, 1], x_edges, y_edges)
    hist = counts.cpu().numpy().astype(np.float64)
    density_grid = hist / (grid_size * grid_size)
    grid_x = (x_edges[:-1] + x_edges[1:]) / 2
    grid_y = (y_edges[:-1] + y_edges[1:]) / 2
    return grid_x, grid_y, density_grid
