"""Drop-in for the reference's models/crowd_flow_model.py (CrowdFlowModel), CUDA-backed.

`analyze(processed_data)` returns the same keys as models/crowd_flow_model.py:28-86.  The reference's
flow field is a closed-form simulation (its `prev_positions` slot is dead code, :16-17); that behaviour
is reproduced, and the dead slot is filled in by `analyze_sequence_frame`, which uses the real
frame-to-frame displacement (NEW op) once a previous frame exists.
"""
from __future__ import annotations

import numpy as np

from .. import flow as _flow
from .. import preprocess as _pre
from ..utils.data_processing import extract_people_positions


class CrowdFlowModel:
    def __init__(self):
        self.prev_positions = None
        self.flow_vectors = None
        self.simulation_params = {
            "flow_field_complexity": 2,
            "bottleneck_count": 3,
            "flow_speed_range": (0.2, 1.5),
            "random_seed": 42,
        }

    @staticmethod
    def _empty():
        return {
            "flow_vectors": {"positions": np.zeros((0, 2)), "vectors": np.zeros((0, 2)), "magnitudes": np.zeros(0)},
            "avg_speed": 0.0, "dominant_direction": "N/A", "bottlenecks": [],
        }

    def analyze(self, processed_data):
        """models/crowd_flow_model.py:28-86."""
        people_positions = extract_people_positions(processed_data)
        if len(people_positions) == 0:
            return self._empty()
        dims = processed_data["dimensions"]
        p = self.simulation_params
        flow, handles, avg_speed, direction = _flow.simulated_flow(
            dims["x_range"], dims["y_range"], variant="A", complexity=p["flow_field_complexity"],
            count=p["bottleneck_count"], speed_range=p["flow_speed_range"], seed=p["random_seed"])
        return {"flow_vectors": flow, "avg_speed": avg_speed, "dominant_direction": direction,
                "bottlenecks": _flow.bottlenecks_a(flow, handles)}

    def analyze_sequence_frame(self, processed_data, dt=0.1, gate=1.5):
        """Frame of a sequence: the first call behaves like `analyze`; later calls replace the simulated
        field by the measured displacement field of the people matched against the previous frame."""
        result, people_positions = self.sequence_step(processed_data, self.prev_positions, dt=dt, gate=gate)
        self.prev_positions = people_positions if len(people_positions) else None
        if "matches" in result:
            self.flow_vectors = result["flow_vectors"]
        return result

    def sequence_step(self, processed_data, prev_positions, dt=0.1, gate=1.5, people_positions=None):
        """`analyze_sequence_frame` without the state: the result of one frame given the people positions of the frame
        before it (None or empty: there is none, the simulated field of `analyze` is returned).  A frame depends on its
        predecessor's POSITIONS only, not on its flow result, so `sequence.SequenceRunner` evaluates the frames of a
        sequence concurrently.  `people_positions`: the frame's own positions when the caller has them already.
        Returns (result, people_positions)."""
        dims = processed_data["dimensions"]
        have_prev = prev_positions is not None and len(prev_positions) > 0
        fast = None
        cache = processed_data.get(_pre.DEVICE_KEY)
        if (people_positions is None and have_prev and isinstance(cache, _pre.DeviceCache) and cache.matches(processed_data)
                and cache.n_clusters):
            # clusters still on the device: centroids -> match -> field without a trip to the host in between
            fast = _flow.frame_flow_from_clusters(prev_positions, cache.points, cache.clusters, cache.n_clusters,
                                                  dt, dims["x_range"], dims["y_range"], gate=gate)
        if fast is not None:
            flow, match, _, people_positions = fast
        else:
            if people_positions is None:
                people_positions = extract_people_positions(processed_data)
            if len(people_positions) == 0:
                return self._empty(), people_positions
            if not have_prev:
                return self.analyze(processed_data), people_positions
            flow, match, _ = _flow.frame_flow(prev_positions, people_positions, dt, dims["x_range"],
                                              dims["y_range"], gate=gate)
        vectors, magnitudes = flow["vectors"], flow["magnitudes"]
        avg_vector = np.mean(vectors, axis=0)
        angle = np.arctan2(avg_vector[1], avg_vector[0]) * 180 / np.pi
        direction = _flow._DIRECTIONS[int((angle + 22.5) % 360 / 45)]
        return ({"flow_vectors": flow, "avg_speed": np.mean(magnitudes), "dominant_direction": direction,
                 "bottlenecks": [], "matches": match}, people_positions)
