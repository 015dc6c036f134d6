"""Device-side flow analysis shared by surface A (models.crowd_flow_model) and surface B (apps).

Host-side parameter derivation only: the lattice axes (np.arange, like upstream), the exit point, the
three bottleneck discs drawn from the legacy MT19937 stream (np.random.seed(42) + uniform — a side
effect on the global numpy RNG that the reference has and callers may observe), and the final top-5
list handling.  Everything per lattice node runs in csrc/flow.cu.
"""
from __future__ import annotations

import numpy as np

from . import ops

_DIRECTIONS = ["E", "NE", "N", "NW", "W", "SW", "S", "SE", "E"]


def lattice_axes(x_range, y_range, grid_size=1.0):
    """models/crowd_flow_model.py:107-109."""
    x_grid = np.arange(x_range[0], x_range[1] + grid_size, grid_size)
    y_grid = np.arange(y_range[0], y_range[1] + grid_size, grid_size)
    return x_grid, y_grid


def draw_discs(x_range, y_range, count=3, seed=42):
    """models/crowd_flow_model.py:100,151-154 / app_simplified.py:366,390-393."""
    np.random.seed(seed)
    discs = []
    for _ in range(count):
        bx = np.random.uniform(x_range[0] + 1, x_range[1] - 1)
        by = np.random.uniform(y_range[0] + 1, y_range[1] - 1)
        discs.append((bx, by))
    return discs


def simulated_flow(x_range, y_range, variant="A", complexity=2, count=3, speed_range=(0.2, 1.5), seed=42):
    """_generate_simulated_flow on the device.  Returns (flow_vectors dict of numpy arrays,
    device handles (pos, vec, mag, nxy), avg_speed, dominant_direction)."""
    x_grid, y_grid = lattice_axes(x_range, y_range)
    discs = draw_discs(x_range, y_range, count, seed)
    exit_xy = (x_range[1], (y_range[0] + y_range[1]) / 2)
    if variant == "A":
        lo, hi = speed_range
        pos, vec, mag, sums, nxy = ops.flow_field(x_grid, y_grid, exit_xy, complexity, 0.5, discs, hi - lo,
                                                  clip=(lo, hi))
    else:
        pos, vec, mag, sums, nxy = ops.flow_field(x_grid, y_grid, exit_xy, 0.3, 0.5, discs, 1.3, clip=None)
    g = nxy[0] * nxy[1]
    s = sums.cpu().numpy()
    avg_speed = s[0] / g
    avg_vector = s[1:3] / g
    angle = np.arctan2(avg_vector[1], avg_vector[0]) * 180 / np.pi
    direction = _DIRECTIONS[int((angle + 22.5) % 360 / 45)]
    flow = {"positions": pos.cpu().numpy(), "vectors": vec.cpu().numpy(), "magnitudes": mag.cpu().numpy()}
    return flow, (pos, vec, mag, nxy), np.float64(avg_speed), direction


def bottlenecks_a(flow, handles):
    """_identify_bottlenecks (models/crowd_flow_model.py:186-279): severity per node on the device, the
    `> 1.0` filter, `min(10, round(sev))` (Python banker's rounding) and the stable top-5 on the host."""
    pos, vec, mag, nxy = handles
    sev = ops.flow_bottleneck_severity(nxy, pos, vec, mag).cpu().numpy()
    P = flow["positions"]
    out = [{"x": P[i, 0], "y": P[i, 1], "severity": min(10, round(float(sev[i])))} for i in np.flatnonzero(sev > 1.0)]
    return sorted(out, key=lambda b: b["severity"], reverse=True)[:5]


def bottlenecks_b(flow, handles):
    """app_simplified.py:425-450: slow nodes (< 0.3) whose open +-3 m box holds a node faster than 0.5."""
    pos, _, mag, nxy = handles
    box = ops.flow_box_max(nxy, pos, mag, 0.3).cpu().numpy()
    P, M = flow["positions"], flow["magnitudes"]
    out = []
    for i in np.flatnonzero(box > 0.5):
        severity = min(10, int(10 * (box[i] - M[i]) / box[i]))
        if severity >= 3:
            out.append({"x": P[i, 0], "y": P[i, 1], "severity": severity})
    return sorted(out, key=lambda b: b["severity"], reverse=True)[:5]


def frame_flow(prev_positions, positions, dt, x_range, y_range, gate=1.5, radius=3.0, _centroids=None):
    """NEW op: real frame-to-frame displacement binned onto the same lattice
    as the simulated field.  Returns the reference's flow_vectors dict plus the match indices.
    (`positions` may be a (C,2) CUDA tensor; `_centroids` = device tensors to bring back in the same read-back.)"""
    x_grid, y_grid = lattice_axes(x_range, y_range)
    # lattice (np.vstack([X.ravel(), Y.ravel()]).T of np.meshgrid: built on the device from the two axes), match and field
    # behind one call, one copy in, one copy out
    extra = tuple(_centroids) if _centroids is not None else ()
    got = ops.frame_flow_step(prev_positions, positions, x_grid, y_grid, dt, gate, radius, extra=extra)
    lattice, h_vec, h_mag, h_match, h_vel = got[:5]
    out = ({"positions": lattice, "vectors": h_vec, "magnitudes": h_mag}, h_match, h_vel)
    return out + (tuple(got[5:]),) if extra else out


def frame_flow_from_clusters(prev_positions, points, clusters, n_clusters, dt, x_range, y_range, gate=1.5, radius=3.0):
    """`frame_flow` for a frame whose clusters are still on the device (sequence mode): the centroids go from the
    accumulation kernel straight into the match, and they come back together with the flow field -- ONE wait per frame
    for the ordered stage instead of two.  Returns (flow, match, velocity, people_positions (C,2) float64), or None when a
    cluster id has no member (caller-built labels): the general path handles that."""
    cent, counts = ops.cluster_centroids(points, clusters, n_clusters)
    cur = cent[:, :2].contiguous()
    flow, match, vel, (h_cent, h_counts) = frame_flow(prev_positions, cur, dt, x_range, y_range, gate, radius,
                                                      _centroids=(cent, counts))
    if not (h_counts > 0).all():
        return None
    return flow, match, vel, np.ascontiguousarray(h_cent[:, :2])
