"""Generate the golden vectors that pin the CPU oracle to the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Variant A functions are imported from /root/reference (utils.data_processing, models.*); variant B
functions are defined inside the Streamlit app files, which cannot be imported (streamlit / plotly
are absent and the UI runs at import), so their FunctionDef nodes are extracted with `ast` and
executed with only {np, DBSCAN, KDTree} in scope — the code that runs is the reference's, verbatim,
in memory; nothing is copied into this repository.

Outputs: tests/golden/*.npz (small) — inputs are regenerated from seeds by
lidar_ai_recommendation_software_b200.synth, only reference OUTPUTS are stored.
"""
from __future__ import annotations

import ast
import hashlib
import os
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
REF = Path(os.environ.get("LIDAR_REFERENCE", "/root/reference"))
sys.dont_write_bytecode = True
sys.path.insert(0, str(ROOT))

from lidar_ai_recommendation_software_b200 import synth  # noqa: E402  (pure numpy, no CUDA)


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def load_variant_a():
    sys.path.insert(0, str(REF))
    from models.crowd_density_model import CrowdDensityModel
    from models.crowd_flow_model import CrowdFlowModel
    from utils import data_processing as dp
    return dp, CrowdDensityModel, CrowdFlowModel


def load_variant_b():
    from sklearn.cluster import DBSCAN
    from sklearn.neighbors import KDTree
    src = (REF / "app_simplified.py").read_text()
    tree = ast.parse(src)
    wanted = {"preprocess_point_cloud", "analyze_crowd_density", "analyze_crowd_flow"}
    ns = {"np": np, "DBSCAN": DBSCAN, "KDTree": KDTree}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in wanted:
            code = compile(ast.Module(body=[node], type_ignores=[]), str(REF / "app_simplified.py"), "exec")
            exec(code, ns)
    return ns


def pack_hotspots(hs, key="density"):
    return np.array([[h["x"], h["y"], h[key]] for h in hs], dtype=np.float64).reshape(-1, 3)


def run_case(name, pts64, dp, CDM, CFM, B):
    out = {"input_sha": np.array(sha(pts64))}
    # ---- variant A
    pa = dp.preprocess_lidar_data(pts64)
    out["a_points_sha"] = np.array(sha(pa["points"]))
    out["a_n_inliers"] = np.array(len(pa["points"]))
    out["a_clusters"] = pa["clusters"].astype(np.int64)
    out["a_colors_sha"] = np.array(sha(pa["colors"]))
    out["a_colors_head"] = pa["colors"][:64]
    out["a_plane"] = pa["ground_plane"].astype(np.float64)
    d = pa["dimensions"]
    out["a_dims"] = np.array([*d["x_range"], *d["y_range"], *d["z_range"], d["width"], d["length"], d["height"]])
    pos = dp.extract_people_positions(pa)
    out["a_people"] = np.asarray(pos, dtype=np.float64).reshape(-1, 2)
    for g in (1.0, 0.5):
        gx, gy, dens = dp.calculate_grid_density(pa["points"][:, :2], d["x_range"], d["y_range"], g)
        out[f"a_grid_counts_g{g}"] = np.rint(dens * g * g).astype(np.int32)
        out[f"a_grid_x_g{g}"] = gx
        out[f"a_grid_y_g{g}"] = gy
    hist, ex, ey = np.histogram2d(pa["points"][:, 0], pa["points"][:, 1], bins=100, range=[d["x_range"], d["y_range"]])
    out["heat_counts"] = hist.astype(np.int32)
    out["heat_ex"] = ex
    out["heat_ey"] = ey
    ra = CDM().analyze(pa)
    out["a_density_scalars"] = np.array([ra["total_people"], ra["avg_density"], ra["max_density"]], dtype=np.float64)
    out["a_density_map"] = ra["density_map"]
    out["a_hotspots"] = pack_hotspots(ra["hotspots"])
    fa = CFM().analyze(pa)
    out["a_flow_positions_sha"] = np.array(sha(fa["flow_vectors"]["positions"]))
    out["a_flow_vectors"] = fa["flow_vectors"]["vectors"]
    out["a_flow_magnitudes"] = fa["flow_vectors"]["magnitudes"]
    out["a_flow_scalars"] = np.array([fa["avg_speed"]])
    out["a_flow_direction"] = np.array(fa["dominant_direction"])
    out["a_bottlenecks"] = pack_hotspots(fa["bottlenecks"], "severity")
    # ---- variant B
    pb = B["preprocess_point_cloud"](pts64)
    out["b_n_inliers"] = np.array(len(pb["points"]))
    out["b_clusters"] = pb["clusters"].astype(np.int64)
    out["b_colors_sha"] = np.array(sha(pb["colors"]))
    rb = B["analyze_crowd_density"](pb)
    out["b_density_scalars"] = np.array([rb["total_people"], rb["avg_density"], rb["max_density"]], dtype=np.float64)
    out["b_density_grid"] = rb["density_grid"]
    out["b_hotspots"] = pack_hotspots(rb["hotspots"])
    fb = B["analyze_crowd_flow"](pb)
    out["b_flow_vectors"] = fb["flow_vectors"]["vectors"]
    out["b_flow_magnitudes"] = fb["flow_vectors"]["magnitudes"]
    out["b_flow_scalars"] = np.array([fb["avg_speed"]])
    out["b_flow_direction"] = np.array(fb["dominant_direction"])
    out["b_bottlenecks"] = pack_hotspots(fb["bottlenecks"], "severity")
    np.savez_compressed(HERE / f"{name}.npz", **out)
    print(f"{name}: inliers A={len(pa['points'])} clusters A={len(np.unique(pa['clusters'][pa['clusters']>=0]))} "
          f"B={rb['total_people']}  people_sha={sha(out['a_people'])}")


def main():
    dp, CDM, CFM = load_variant_a()
    B = load_variant_b()
    # case 1: the reference's own demo cloud (app_simplified.py:994-1024)
    run_case("ref_sample_10k", synth.reference_sample(), dp, CDM, CFM, B)
    # case 2: synthetic crowd frame C.1, 20 k points on a 30 m x 30 m patch (with injected outliers)
    run_case("crowd_20k", synth.add_outliers(synth.crowd_frame(20000, seed=3, extent=15.0))[:, :3].astype(np.float64),
             dp, CDM, CFM, B)
    # case 3: BASELINE config 1 — 100 k-point crowd frame, 100 m x 100 m
    run_case("crowd_100k", synth.crowd_frame(100000, seed=0, extent=50.0)[:, :3].astype(np.float64), dp, CDM, CFM, B)
    # case 4: degenerate guards — 14 points: <= 10 ground points (fallback plane, data_processing.py:181-183) and
    #         <= 10 non-ground points (all-zero labels, :199-200)
    run_case("tiny_14", synth.tiny_cloud(), dp, CDM, CFM, B)
    # case 5: 300 scattered returns — several clusters in variant A (eps in sigma units), NO cluster in variant B
    #         (the empty-people dicts of app_simplified.py:234-316 / 318-464)
    run_case("sparse_300", synth.sparse_cloud(), dp, CDM, CFM, B)


if __name__ == "__main__":
    main()
