"""`ProjectManager.run_analysis` without the simulation.

Upstream (windows_implementation/core/project_manager.py:291-348) fills the result dict with made-up numbers
("This is where the actual analysis would occur").  `run_analysis(points, parameters)` returns a dict with
exactly the keys that block produces and `main.py:328-335` / `database_manager.insert_analysis` consume —

    total_people, avg_density, max_density, density_map, hotspots,
    avg_speed, dominant_direction, bottlenecks, timestamp

— computed by the CUDA-backed pipeline: preprocess -> CrowdDensityModel.analyze -> CrowdFlowModel.analyze.
`Dataset.points` ((n,3) float64, core/data_loader.py:15-27) is the input.  The dict is JSON-serialisable through
`convert_numpy` (the same conversion as core/database_manager.py:501-508).

Swap-in inside ProjectManager.run_analysis (replaces lines 291-348):

    from lidar_ai_recommendation_software_b200.windows_core import run_analysis as _b200_run
    results = _b200_run(self.current_dataset.points, parameters)
"""
from __future__ import annotations

from datetime import datetime

import numpy as np

from .. import preprocess as _pre
from ..models.crowd_density_model import CrowdDensityModel
from ..models.crowd_flow_model import CrowdFlowModel


def convert_numpy(obj):
    """numpy -> plain python, recursively (core/database_manager.py:501-508 semantics)."""
    if isinstance(obj, np.ndarray):
        return obj.tolist()
    if isinstance(obj, (np.integer,)):
        return int(obj)
    if isinstance(obj, (np.floating,)):
        return float(obj)
    if isinstance(obj, dict):
        return {k: convert_numpy(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [convert_numpy(v) for v in obj]
    return obj


def run_analysis(points, parameters=None):
    """Analysis results for one dataset.  `parameters` (all optional): `grid_size` (density cell, default 1.0 m),
    `variant` ("B" = the apps' preprocess with DBSCAN eps 0.3 m on raw metres, default; "A" =
    utils.data_processing.preprocess_lidar_data)."""
    parameters = parameters or {}
    variant = parameters.get("variant", "B")
    if variant not in ("A", "B"):
        raise ValueError("parameters['variant'] must be 'A' or 'B'")
    processed = _pre.run(points, variant=variant, host_arrays=False)
    density = CrowdDensityModel(grid_size=float(parameters.get("grid_size", 1.0))).analyze(processed)
    flow = CrowdFlowModel().analyze(processed)
    return {
        "total_people": int(density["total_people"]),
        "avg_density": float(density["avg_density"]),
        "max_density": float(density["max_density"]),
        "density_map": density["density_map"],
        "hotspots": [{"x": float(h["x"]), "y": float(h["y"]), "density": float(h["density"])} for h in density["hotspots"]],
        "avg_speed": float(flow["avg_speed"]),
        "dominant_direction": flow["dominant_direction"],
        "bottlenecks": [{"x": float(b["x"]), "y": float(b["y"]), "severity": int(b["severity"])} for b in flow["bottlenecks"]],
        "timestamp": datetime.now().isoformat(),
    }
