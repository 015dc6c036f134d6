"""Drop-in for the reference's models/crowd_density_model.py (CrowdDensityModel), CUDA-backed.

Same constructor, `analyze(processed_data)` result keys / dtypes / shapes and `calculate_risk_level`
thresholds as models/crowd_density_model.py:6-117; centroids and the density histogram are computed
on the device (K8 + K6), the top-5 hotspot selection is host-side list handling like upstream.
"""
from __future__ import annotations

import numpy as np

from ..utils.data_processing import calculate_grid_density, extract_people_positions


class CrowdDensityModel:
    """Crowd density from a clustered point cloud (models/crowd_density_model.py:6-21)."""

    def __init__(self, grid_size=1.0):
        self.grid_size = grid_size

    def analyze(self, processed_data):
        """models/crowd_density_model.py:23-98."""
        people_positions = extract_people_positions(processed_data)
        if len(people_positions) == 0:
            return {
                "total_people": 0, "avg_density": 0.0, "max_density": 0.0, "density_map": np.zeros((1, 1)),
                "grid_coordinates": (np.array([0]), np.array([0])), "density_values": np.array([0]),
                "hotspots": [],
            }
        x_range = processed_data["dimensions"]["x_range"]
        y_range = processed_data["dimensions"]["y_range"]
        grid_x, grid_y, density_grid = calculate_grid_density(people_positions, x_range, y_range, self.grid_size)

        flat_density = density_grid.flatten()
        flat_x = np.repeat(grid_x, len(grid_y))
        flat_y = np.tile(grid_y, len(grid_x))
        total_people = len(people_positions)
        max_density = np.max(flat_density)
        occupied = flat_density > 0
        avg_density = np.mean(flat_density[occupied]) if np.any(occupied) else 0

        threshold = max(0.5, avg_density * 1.5)
        # upstream builds a dict per cell over the threshold and keeps the first five of a stable descending sort;
        # a stable argsort of the negated densities picks the same five without the per-cell Python objects
        cand = np.where(flat_density >= threshold)[0]
        top = cand[np.argsort(-flat_density[cand], kind="stable")[:5]]
        hotspots = [{"x": flat_x[i], "y": flat_y[i], "density": flat_density[i]} for i in top]
        return {
            "total_people": total_people, "avg_density": avg_density, "max_density": max_density,
            "density_map": density_grid, "grid_coordinates": (flat_x, flat_y), "density_values": flat_density,
            "hotspots": hotspots,
        }

    def calculate_risk_level(self, density):
        """models/crowd_density_model.py:100-117."""
        if density < 1.0:
            return "Low"
        elif density < 2.5:
            return "Moderate"
        elif density < 4.0:
            return "High"
        return "Critical"
