"""Device pipeline behind both preprocess surfaces:

    variant "A"  utils.data_processing.preprocess_lidar_data   (utils/data_processing.py:127-229)
    variant "B"  preprocess_point_cloud of the Streamlit apps   (app_simplified.py:76-137)

Every per-point stage runs in the CUDA core (bbox, moments, 3-sigma filter + compaction + colours,
radix select, ground split + plane moments, scaler, DBSCAN, label scatter).  The stages before DBSCAN are
ONE enqueue (`lidar_preprocess_front`): the scalars one stage hands to the next (mean/std from sums, the
percentile rank and lerp, the scaler statistics, eps) are derived on the device with the same float64
expressions numpy uses, and come back in one read-back; a whole call waits on the device two or three
times (front, DBSCAN's grid sizing needs the bbox on the host, results).  The host solves the 3x3 plane
system.  There is no CPU path for the per-point work.
"""
from __future__ import annotations

import math
import warnings
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _capi, ops

DEVICE_KEY = "_lidar_b200"   # extra dict entry carrying the device-resident copy of the outputs


@dataclass
class DeviceCache:
    """Device-resident twin of a processed_data dict, so later stages skip the H2D copy.
    `ids` ties it to the host arrays it mirrors (stale caches are ignored)."""
    points: torch.Tensor
    clusters: torch.Tensor
    ids: tuple
    n_clusters: int
    guards: dict = field(default_factory=dict)
    positions: np.ndarray | None = None      # people positions, when the producer has computed them already (sequence mode)

    def matches(self, processed: dict) -> bool:
        if self.ids is None:      # device-only dict (run(..., host_arrays=False)): nothing to go stale
            return "points" not in processed and "clusters" not in processed
        return self.ids == (id(processed.get("points")), id(processed.get("clusters")))


def _as_f64x3(points) -> np.ndarray:
    pts = np.asarray(points)
    if pts.ndim != 2 or pts.shape[1] < 3:
        raise ValueError(f"expected an (n,3) point array, got shape {pts.shape}")
    return np.ascontiguousarray(pts[:, :3], dtype=np.float64)


def _to_host(*tensors):
    """Device tensors -> numpy arrays through page-locked memory: all copies are enqueued, then ONE wait.
    (`t.cpu()` goes through pageable memory: ~10 GB/s and a synchronisation per array; 56 MB of outputs per
    1 M-point frame made that the dominant cost of the drop-in call.)  The arrays own their (pinned) storage
    through torch's caching host allocator; it is returned to the cache when the array is garbage-collected."""
    outs = []
    for t in tensors:
        if t is None:
            outs.append(None)
            continue
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t, non_blocking=True)
        outs.append(h)
    torch.cuda.current_stream().synchronize()
    return [None if h is None else h.numpy() for h in outs]


def percentile_from_order_stats(a: float, b: float, n: int, q: float) -> float:
    """np.percentile(x, q) (method 'linear') from x_(lo), x_(lo+1): numpy/lib/_function_base_impl.py
    `_quantile` + `_lerp` — virtual index (n-1)*q/100, lerp a + (b-a)*t, or b - (b-a)*(1-t) for t>=.5."""
    quant = np.float64(q) / 100.0
    virt = (n - 1) * quant
    lo = math.floor(virt)
    t = virt - lo
    a, b = np.float64(a), np.float64(b)
    d = b - a
    return float(b - d * (1 - t)) if t >= 0.5 else float(a + d * t)


def _plane_from_sums(s: np.ndarray, center) -> np.ndarray:
    """Least-squares plane z = ax + by + c from the centred normal-equation sums of lidar_ground_split;
    returned as [a, b, -1, c] like utils/data_processing.py:175-177."""
    n, sx, sy, sz, sxx, sxy, syy, sxz, syz = (float(v) for v in s[:9])
    M = np.array([[sxx, sxy, sx], [sxy, syy, sy], [sx, sy, n]], dtype=np.float64)
    rhs = np.array([sxz, syz, sz], dtype=np.float64)
    a, b, c = np.linalg.lstsq(M, rhs, rcond=None)[0]      # 3x3 system; min-norm if rank deficient
    c0 = (c + center[2]) - a * center[0] - b * center[1]
    return np.array([a, b, -1, c0])


def run(points, variant: str = "A", host_arrays: bool = True) -> dict:
    """Full preprocess on the GPU; returns the reference's processed_data dict (numpy arrays) plus a
    DeviceCache under DEVICE_KEY.

    `points` is the reference's (n,3) array, or an (n,3) float64 CUDA tensor that is already resident.  Any other
    dtype (float32, integers) is widened to float64 first: the reference computes in float64 on float64 loader
    output (utils/data_processing.py:34-41), and the widening of a float32 cloud is exact — note that the REFERENCE
    itself, fed float32, would compute its means and thresholds in float32 under NumPy 2;
    results here are those of the widened cloud.

    Knife edges: a point within 1e-9 sigma of the 3-sigma threshold, or (variant A) a pair within 1e-12 of eps^2, could
    be decided differently by numpy's strictly sequential mean / the scaler's rounding.  The counts of such cases are
    returned under `out["_lidar_b200"].guards`; when one is non-zero a `RuntimeWarning` says so (zero proves the mask
    and labels equal the reference's for this input).
    `host_arrays=False` is the sequence mode (BASELINE configs[3]): the per-point outputs (points, colors,
    normals, clusters) stay on the device under DEVICE_KEY and are NOT copied back; the dict carries only
    `dimensions` (+ `ground_plane`).  extract_people_positions / the density and flow models accept it."""
    dev = ops.require_cuda()
    if isinstance(points, torch.Tensor) and points.is_cuda:
        if points.dim() != 2 or points.shape[1] != 3 or points.dtype != torch.float64:
            raise ValueError("a CUDA input must be an (n,3) float64 tensor")
        d_pts = points.contiguous()
        n = d_pts.shape[0]
    else:
        host = _as_f64x3(points)
        n = host.shape[0]
        d_pts = None
    if n == 0:
        raise ValueError("zero-size array to reduction operation minimum which has no identity")
    if d_pts is None:
        d_pts = torch.from_numpy(host).to(dev, non_blocking=False)

    # --- everything before DBSCAN is one enqueue and one read-back (lidar_preprocess_front): z min/max of the RAW
    #     cloud for the colours (data_processing.py:143), mean / population std (:151-152), the 3-sigma filter
    #     (:153-157), np.percentile(z, 30) and the ground split (:164-166), the plane sums (:169-177), the bbox of the
    #     inliers (:207-217) and, for variant A, StandardScaler + the adaptive eps (:190-196) ----------------------
    desc, inl, col, ng, ng_index, X = ops.preprocess_front(d_pts, want_colors=host_arrays, scaler=(variant == "A"))
    n_in = inl.shape[0]
    if n_in == 0:
        raise IndexError("index -1 is out of bounds for axis 0 with size 0")   # np.percentile of an empty array
    mean = np.array(desc.mean)
    plane_sums = np.array(desc.plane)
    n_ground = int(round(plane_sums[0]))
    m = ng.shape[0]
    lo3, hi3 = np.array(desc.bbox_in[:3]), np.array(desc.bbox_in[3:])

    out: dict = {}
    guards = {"sigma": int(desc.guard_sigma), "dbscan": 0}
    if variant == "A":
        if n_ground > 10:
            plane = _plane_from_sums(plane_sums, mean)
        else:
            plane = np.array([0, 0, 1, -lo3[2]])
    # --- clustering of the non-ground points (:186-200 / app_simplified.py:102-110) -------------
    info = None
    if m > 10:
        b_lo, b_hi = np.array(desc.bbox_ng[:3]), np.array(desc.bbox_ng[3:])
        if variant == "A":
            sc_mean, scale = np.array(desc.sc_mean), np.array(desc.scale)
            b_lo, b_hi = (b_lo - sc_mean) / scale, (b_hi - sc_mean) / scale   # the scaler is monotone per axis
            Xd, eps, tol_db = X, float(desc.eps), 1e-12
        else:
            Xd, eps, tol_db = ng, 0.3, 0.0
        labels, info = ops.dbscan(Xd, eps, 5, tol=tol_db, bounds=(b_lo, b_hi), defer=True)
        n_clusters = None
    else:
        labels = torch.zeros(m, dtype=torch.int32, device=dev)
        n_clusters = 1 if m > 0 else 0
    full = ops.scatter_labels(labels, ng_index, n_in)

    dims = {
        "x_range": (lo3[0], hi3[0]), "y_range": (lo3[1], hi3[1]), "z_range": (lo3[2], hi3[2]),
        "width": hi3[0] - lo3[0], "length": hi3[1] - lo3[1], "height": hi3[2] - lo3[2],
    }

    def _info(h):
        nc, guard = (int(v) for v in h.tolist())
        guards["dbscan"] = guard
        return nc & 0xffffffff

    def _report_guards():
        if guards["sigma"] or guards["dbscan"]:
            warnings.warn(f"lidar_b200 preprocess: {guards['sigma']} point(s) within 1e-9 sigma of the 3-sigma "
                          f"threshold and {guards['dbscan']} neighbour pair(s) within 1e-12 of eps^2 — the inlier mask / "
                          "cluster labels may differ from numpy / scikit-learn for exactly those cases", RuntimeWarning,
                          stacklevel=3)

    if not host_arrays:
        if info is not None:
            n_clusters = _info(ops.fetch("dbscan_info", info)[0])
        if variant == "A":
            out["ground_plane"] = plane
        out["dimensions"] = dims
        out[DEVICE_KEY] = DeviceCache(inl, full, None, n_clusters, guards)
        _report_guards()
        return out
    h_points, h_clusters, h_colors, h_info = _to_host(inl, full, col, info)
    if info is not None:
        n_clusters = _info(h_info)
    out["points"] = h_points
    out["colors"] = h_colors
    if variant == "A":
        normals = np.zeros_like(h_points)
        normals[:, 2] = 1.0
        out["normals"] = normals
    out["clusters"] = h_clusters
    if variant == "A":
        out["ground_plane"] = plane
    out["dimensions"] = dims
    out[DEVICE_KEY] = DeviceCache(inl, full, (id(h_points), id(h_clusters)), n_clusters, guards)
    _report_guards()
    return out


def run_sequence_frame(points: torch.Tensor, centroid_cap: int = 8192) -> dict:
    """`run(points, variant="B", host_arrays=False)` followed by `people_positions`, behind ONE call of the C ABI
    (`lidar_sequence_frame_b`): the same entries in the same order, the host steps between them (wait for the front's
    descriptor, size the cell grid, wait for the cluster count, size the accumulators) in C.  `points`: (n,3) float64 CUDA
    tensor.  Returns the same dict; the people positions ride in the DeviceCache, so `extract_people_positions` and the
    flow step find them without another kernel.  The whole call runs outside the interpreter lock."""
    dev = ops.require_cuda()
    if not (isinstance(points, torch.Tensor) and points.is_cuda and points.dim() == 2 and points.shape[1] == 3
            and points.dtype == torch.float64):
        raise ValueError("run_sequence_frame takes an (n,3) float64 CUDA tensor")
    d_pts = points.contiguous()
    n = d_pts.shape[0]
    if n == 0:
        raise ValueError("zero-size array to reduction operation minimum which has no identity")
    lib, C = _capi.lib, _capi.C
    inl = torch.empty_like(d_pts)
    ng = torch.empty_like(d_pts)
    idx = torch.empty(n, dtype=torch.int32, device=dev)
    labels = torch.empty(n, dtype=torch.int32, device=dev)
    full = torch.empty(n, dtype=torch.int64, device=dev)
    S = ops._scratch
    cap = int(centroid_cap)
    small = S.get("seq_small", C.sizeof(_capi.FrontDesc) + 128 + cap * 32, dev)        # descriptor | info | centroids | counts
    o_info = (C.sizeof(_capi.FrontDesc) + 15) & ~15
    o_cent, o_cnt = o_info + 64, o_info + 64 + cap * 24
    ws_front = S.get("front", lib.lidar_preprocess_front_workspace_bytes(n), dev)
    ws_cent = S.get("centroid", lib.lidar_centroid_workspace_bytes(cap), dev)
    h_pin = ops._pinned.get("seq_frame", C.sizeof(_capi.FrontDesc) + 64 + cap * 32)
    out = _capi.SequenceFrameOut()
    ws_db = S.bufs.get(("dbscan", dev.index))
    base = small.data_ptr()
    for attempt in range(2):
        rc = lib.lidar_sequence_frame_b(d_pts.data_ptr(), n, 0.3, 5, inl.data_ptr(), ng.data_ptr(), idx.data_ptr(),
                                        labels.data_ptr(), full.data_ptr(), base + o_cent, base + o_cnt, cap, base,
                                        base + o_info, h_pin.data_ptr(), h_pin.numel(), ws_front.data_ptr(), ws_front.numel(),
                                        ws_db.data_ptr() if ws_db is not None else None, ws_db.numel() if ws_db is not None else 0,
                                        ws_cent.data_ptr(), ws_cent.numel(), C.byref(out), ops._stream_ptr())
        if rc == -3 and attempt == 0 and out.need_dbscan_ws:      # LIDAR_ERR_WORKSPACE: the cell grid of this bbox needs more
            ws_db = S.get("dbscan", int(out.need_dbscan_ws) + (int(out.need_dbscan_ws) >> 2), dev)   # 25 % head room: the bbox moves
            continue
        _capi.check(rc)
        break
    desc = out.front
    n_in, m = int(desc.n_in), int(desc.n_nonground)
    if n_in == 0:
        raise IndexError("index -1 is out of bounds for axis 0 with size 0")   # np.percentile of an empty array
    lo3, hi3 = np.array(desc.bbox_in[:3]), np.array(desc.bbox_in[3:])
    dims = {
        "x_range": (lo3[0], hi3[0]), "y_range": (lo3[1], hi3[1]), "z_range": (lo3[2], hi3[2]),
        "width": hi3[0] - lo3[0], "length": hi3[1] - lo3[1], "height": hi3[2] - lo3[2],
    }
    guards = {"sigma": int(desc.guard_sigma), "dbscan": int(out.guard_dbscan)}
    nc = int(out.n_clusters)
    positions = None
    if nc == 0:
        positions = np.array([])
    elif out.centroids_done:
        hv = h_pin.numpy()
        o = C.sizeof(_capi.FrontDesc) + 64
        cent = np.frombuffer(hv, dtype=np.float64, count=3 * nc, offset=o).reshape(nc, 3)
        counts = np.frombuffer(hv, dtype=np.int64, count=nc, offset=o + cap * 24)
        positions = np.ascontiguousarray(cent[counts > 0][:, :2])
    result = {"dimensions": dims}
    result[DEVICE_KEY] = DeviceCache(inl[:n_in], full[:n_in], None, nc, guards, positions)
    if guards["sigma"] or guards["dbscan"]:
        warnings.warn(f"lidar_b200 preprocess: {guards['sigma']} point(s) within 1e-9 sigma of the 3-sigma threshold and "
                      f"{guards['dbscan']} neighbour pair(s) within 1e-12 of eps^2", RuntimeWarning, stacklevel=2)
    return result


def device_view(processed: dict):
    """(points (n,3) f64 CUDA, clusters (n,) int64 CUDA) of a processed_data dict — from the cache
    when it still mirrors the host arrays, else uploaded."""
    cache = processed.get(DEVICE_KEY)
    if isinstance(cache, DeviceCache) and cache.matches(processed):
        return cache.points, cache.clusters
    dev = ops.require_cuda()
    pts = torch.from_numpy(_as_f64x3(processed["points"])).to(dev)
    lab = torch.from_numpy(np.ascontiguousarray(processed["clusters"], dtype=np.int64)).to(dev)
    return pts, lab


def people_positions(processed: dict) -> np.ndarray:
    """extract_people_positions (utils/data_processing.py:251-280): centroid xy of every cluster id >= 0,
    ascending id.  Labels of our own DBSCAN are 0..C-1 and go straight to the device accumulators; a
    caller-built dict may carry ANY int64 ids (the reference's np.unique accepts them): sparse or huge ids are ranked
    with np.unique on the host first, so the accumulators are sized by the number of clusters, never by the largest id."""
    cache = processed.get(DEVICE_KEY)
    if isinstance(cache, DeviceCache) and cache.positions is not None and cache.matches(processed):
        return cache.positions.copy()
    pts, lab = device_view(processed)
    if lab.numel() == 0:
        return np.array([])
    if isinstance(cache, DeviceCache) and cache.matches(processed) and cache.n_clusters is not None:
        n_ids = cache.n_clusters             # labels of our own DBSCAN: 0..n_clusters-1, known without a read-back
    else:
        n_ids = int(lab.max().item()) + 1
        if n_ids > lab.numel() + 1024:
            # sparse ids: dense ranks (ascending id order is preserved, which is the output order of the reference)
            host = lab.cpu().numpy()
            ids = np.unique(host[host >= 0])
            rank = np.where(host >= 0, np.searchsorted(ids, host), -1).astype(np.int64)
            lab = torch.from_numpy(rank).to(lab.device)
            n_ids = len(ids)
    if n_ids <= 0:
        return np.array([])
    cent, counts = ops.cluster_centroids(pts, lab, n_ids)
    h_cent, h_counts = ops.fetch("centroids", cent, counts)
    return np.ascontiguousarray(h_cent[h_counts > 0][:, :2])      # fancy index: a fresh array, not the staging view
