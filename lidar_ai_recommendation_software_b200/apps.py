"""Drop-in for the compute functions defined inline in the reference's Streamlit apps
(app_simplified.py:76-464, byte-identical in app_with_db.py:80-468) — variant B semantics and keys.

Swap-in: add one line after the inline definitions of the app file (before its UI section,
app_simplified.py:942):

    from lidar_ai_recommendation_software_b200.apps import (preprocess_point_cloud, create_density_heatmap,
                                                            analyze_crowd_density, analyze_crowd_flow)

All four names the apps call (app_simplified.py:1028, 1074, 1159/1184, 1220) are shadowed by that one line; nothing
inside the app files is edited.
"""
from __future__ import annotations

import numpy as np
import torch

from . import flow as _flow
from . import ops
from . import preprocess as _pre


def preprocess_point_cloud(points):
    """app_simplified.py:76-137: keys points, colors, clusters, dimensions (DBSCAN eps=0.3 on raw metres)."""
    return _pre.run(points, variant="B")


def density_heatmap_counts(processed_data, bins=100, projection=(0, 1)):
    """The histogram core of create_density_heatmap (app_simplified.py:198-209,
    utils/visualization.py:116-134): np.histogram2d(u, v, bins, range) of the raw points.
    Returns (hist (bins,bins) float64 like numpy, x_edges, y_edges) — the caller transposes for display."""
    pts, _ = _pre.device_view(processed_data)
    names = ("x_range", "y_range", "z_range")
    r0 = processed_data["dimensions"][names[projection[0]]]
    r1 = processed_data["dimensions"][names[projection[1]]]
    x_edges = np.linspace(r0[0], r0[1], bins + 1)
    y_edges = np.linspace(r1[0], r1[1], bins + 1)
    counts = ops.hist2d_counts(pts[:, projection[0]], pts[:, projection[1]], x_edges, y_edges)
    return counts.cpu().numpy().astype(np.float64), x_edges, y_edges


def density_heatmap_spec(processed_data, bins=100):
    """Everything create_density_heatmap (app_simplified.py:198-232) puts into its figure, as plain arrays:
    z = hist.T (rows = y), x / y = bin centres, and the layout strings.  The histogram runs on the device."""
    hist, x_edges, y_edges = density_heatmap_counts(processed_data, bins=bins)
    return {
        "z": hist.T, "x": (x_edges[:-1] + x_edges[1:]) / 2, "y": (y_edges[:-1] + y_edges[1:]) / 2,
        "colorscale": "Viridis", "colorbar": dict(title="Point Density"),
        "layout": dict(xaxis_title="X (m)", yaxis_title="Y (m)", title="Point Density Heatmap", height=500),
    }


def create_density_heatmap(processed_data):
    """app_simplified.py:198-232 under its own name and return contract: a plotly `go.Figure` holding one
    `go.Heatmap(z=hist.T, x=x_centers, y=y_centers, colorscale='Viridis', colorbar=dict(title='Point Density'))`
    with the reference's layout.  np.histogram2d(bins=100, range=[x_range, y_range]) of the raw points is the device
    histogram (`density_heatmap_counts`); plotly is imported here, as the apps themselves import it
    (app_simplified.py:6) — a host without plotly gets the same ImportError the app would raise."""
    import plotly.graph_objects as go
    spec = density_heatmap_spec(processed_data)
    fig = go.Figure(data=go.Heatmap(z=spec["z"], x=spec["x"], y=spec["y"], colorscale=spec["colorscale"],
                                    colorbar=spec["colorbar"]))
    fig.update_layout(**spec["layout"])
    return fig


def _people(processed_data):
    """Centroids (C,3) on the device + cluster count (app_simplified.py:239-255)."""
    pts, lab = _pre.device_view(processed_data)
    if lab.numel() == 0:
        return None, 0
    mx = int(lab.max().item())
    if mx < 0:
        return None, 0
    cent, counts = ops.cluster_centroids(pts, lab, mx + 1)
    keep = counts > 0
    cent = cent[keep]
    return cent, int(cent.shape[0])


def analyze_crowd_density(processed_data):
    """app_simplified.py:234-316: keys total_people, avg_density, max_density, density_grid [y][x], hotspots."""
    cent, num_people = _people(processed_data)
    dims = processed_data["dimensions"]
    area = dims["width"] * dims["length"]
    avg_density = num_people / max(1, area)
    if num_people > 0:
        x_range, y_range = dims["x_range"], dims["y_range"]
        grid_size = 1.0
        x_grid = np.arange(x_range[0], x_range[1] + grid_size, grid_size)
        y_grid = np.arange(y_range[0], y_range[1] + grid_size, grid_size)
        cx = (x_grid[:-1] + x_grid[1:]) / 2
        cy = (y_grid[:-1] + y_grid[1:]) / 2
        if len(cx) and len(cy):
            counts = ops.radius_count(cent[:, :2].contiguous(), cx, cy, 2.0).cpu().numpy()
            density_grid = counts / 4.0
        else:
            density_grid = np.zeros((len(y_grid) - 1, len(x_grid) - 1))
        max_density = np.max(density_grid)
        threshold = max(0.5, avg_density * 1.5)
        jj, ii = np.nonzero(density_grid >= threshold)          # row-major: j outer, i inner — upstream's loop order
        # upstream's stable descending sort of one dict per hot cell, cut to five = a stable argsort of the negated values
        top = np.argsort(-density_grid[jj, ii], kind="stable")[:5]
        hot = [{"x": cx[ii[t]], "y": cy[jj[t]], "density": density_grid[jj[t], ii[t]]} for t in top]
    else:
        density_grid, max_density, hot = np.zeros((1, 1)), 0, []
    return {"total_people": num_people, "avg_density": avg_density, "max_density": max_density,
            "density_grid": density_grid, "hotspots": hot}


def analyze_crowd_flow(processed_data):
    """app_simplified.py:318-464: keys avg_speed, dominant_direction, bottlenecks, flow_vectors."""
    _, num_people = _people(processed_data)
    if num_people == 0:
        return {"avg_speed": 0, "dominant_direction": "N/A", "bottlenecks": [],
                "flow_vectors": {"positions": np.zeros((0, 2)), "vectors": np.zeros((0, 2)), "magnitudes": np.zeros(0)}}
    dims = processed_data["dimensions"]
    flow, handles, avg_speed, direction = _flow.simulated_flow(dims["x_range"], dims["y_range"], variant="B")
    return {"avg_speed": avg_speed, "dominant_direction": direction,
            "bottlenecks": _flow.bottlenecks_b(flow, handles), "flow_vectors": flow}
