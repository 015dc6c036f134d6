"""The numpy / scikit-learn arithmetic the reference's hot path delegates to, restated.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference vendors none of this: it calls
numpy (`arange`, `linspace`, `histogram2d`, `percentile`, `mean`, `std`) and scikit-learn
(`DBSCAN`, `KDTree.query_radius`, `StandardScaler`) from wheels that it never pins.  The versions
restated here are the ones in this image: numpy 2.3.5, scikit-learn 1.9.0.  Each function names the
upstream routine whose published algorithm it follows and is checked against the real library in
tests/test_oracle.py.
"""
from __future__ import annotations

import math

import numpy as np


# --------------------------------------------------------------------------------------------
# numpy.arange for float64 (numpy/_core/src/multiarray/ctors.c: PyArray_ArangeObj + DOUBLE_fill)
#   length = ceil((stop - start) / step); a[0] = start; a[1] = start + step;
#   delta = a[1] - a[0]; a[i] = start + i * delta   (separate multiply and add, no FMA)
# call sites: utils/data_processing.py:312-313, models/crowd_flow_model.py:108-109,
#             app_simplified.py:262-263, 354-355
# --------------------------------------------------------------------------------------------
def arange_f64(start: float, stop: float, step: float) -> np.ndarray:
    start, stop, step = np.float64(start), np.float64(stop), np.float64(step)
    length = int(math.ceil((stop - start) / step))
    if length <= 0:
        return np.empty(0, dtype=np.float64)
    out = np.empty(length, dtype=np.float64)
    out[0] = start
    if length > 1:
        nxt = start + step
        out[1] = nxt
        delta = nxt - start
        idx = np.arange(2, length, dtype=np.float64)
        out[2:] = start + idx * delta
    return out


# numpy.linspace(lo, hi, n+1) (numpy/_core/function_base.py): step = (hi-lo)/n,
# y = arange(0, n+1) * step + lo, last element forced to hi.
# call sites: np.histogram2d(bins=int, range=...) at utils/visualization.py:130-134,
#             app_simplified.py:205-209 (via numpy/lib/_histograms_impl.py:_get_bin_edges)
def linspace_edges(lo: float, hi: float, bins: int) -> np.ndarray:
    lo, hi = np.float64(lo), np.float64(hi)
    step = (hi - lo) / bins
    y = np.arange(0, bins + 1, dtype=np.float64) * step + lo
    y[-1] = hi
    return y


# --------------------------------------------------------------------------------------------
# numpy.histogramdd (numpy/lib/_histograms_impl.py:histogramdd), 2-D, explicit edges:
#   per axis b = searchsorted(edges, x, side='right'); samples equal to the last edge move into
#   the last bin; flat index over (len(edges)+1)-sized axes; bincount; outlier rows/cols sliced off.
# --------------------------------------------------------------------------------------------
def histogram2d_counts(u: np.ndarray, v: np.ndarray, ex: np.ndarray, ey: np.ndarray) -> np.ndarray:
    """Integer counts, shape (len(ex)-1, len(ey)-1), index order [u-bin][v-bin]."""
    u = np.asarray(u, dtype=np.float64)
    v = np.asarray(v, dtype=np.float64)
    bu = np.searchsorted(ex, u, side="right")
    bv = np.searchsorted(ey, v, side="right")
    bu[u == ex[-1]] -= 1
    bv[v == ey[-1]] -= 1
    nu, nv = len(ex) + 1, len(ey) + 1
    flat = bu.astype(np.int64) * nv + bv
    full = np.bincount(flat, minlength=nu * nv).reshape(nu, nv)
    return full[1:-1, 1:-1].astype(np.int64)


# --------------------------------------------------------------------------------------------
# numpy.percentile(x, q), method='linear' (numpy/lib/_function_base_impl.py: _quantile, _lerp)
# call sites: utils/data_processing.py:164, app_simplified.py:98
# --------------------------------------------------------------------------------------------
def percentile_linear(x: np.ndarray, q: float) -> float:
    x = np.sort(np.asarray(x, dtype=np.float64))
    n = x.size
    quant = np.float64(q) / 100.0
    virt = (n - 1) * quant
    lo = int(math.floor(virt))
    hi = min(lo + 1, n - 1)
    t = virt - lo
    a, b = x[lo], x[hi]
    d = b - a
    return float(b - d * (1 - t)) if t >= 0.5 else float(a + d * t)


# --------------------------------------------------------------------------------------------
# np.mean / np.std over axis 0 of a C-contiguous (n,3) float64 array: the add-reduce walks the rows
# strictly sequentially (no pairwise blocking along the non-contiguous reduction axis) — verified
# against numpy in tests/test_oracle.py.  call sites: utils/data_processing.py:151-152.
# --------------------------------------------------------------------------------------------
def mean_std_axis0(p: np.ndarray):
    p = np.asarray(p, dtype=np.float64)
    n = p.shape[0]
    s = np.zeros(p.shape[1])
    for row in p:  # small inputs only; the tests compare with np.mean on larger ones
        s = s + row
    mean = s / n
    s2 = np.zeros(p.shape[1])
    for row in p:
        d = row - mean
        s2 = s2 + d * d
    return mean, np.sqrt(s2 / n)


# --------------------------------------------------------------------------------------------
# sklearn.cluster.DBSCAN(eps, min_samples).fit(X).labels_ in closed form
# (sklearn/cluster/_dbscan.py + _dbscan_inner.pyx, sklearn 1.9.0):
#   neighbourhood: squared euclidean distance accumulated over the axes in fp64, INCLUSIVE
#   `<= eps*eps`, a point is its own neighbour; core <=> #neighbours >= min_samples;
#   the DFS of dbscan_inner labels, in ascending seed order, every point reachable from a core
#   point through core points.  Equivalent closed form: connected components of the core-core
#   graph, numbered by the rank of their smallest core index; a border point takes the smallest
#   cluster id among its core neighbours; everything else is -1.
# call sites: utils/data_processing.py:197 (on StandardScaler output), app_simplified.py:107 (raw).
# --------------------------------------------------------------------------------------------
def dbscan_labels(X: np.ndarray, eps: float, min_samples: int = 5) -> np.ndarray:
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    from scipy.spatial import cKDTree

    X = np.ascontiguousarray(X, dtype=np.float64)
    n = X.shape[0]
    if n == 0:
        return np.empty(0, dtype=np.int64)
    tree = cKDTree(X)
    # candidate pairs from a slightly inflated radius, then the exact inclusive fp64 test
    pairs = tree.query_pairs(r=float(eps) * (1.0 + 1e-9) + 1e-300, output_type="ndarray")
    if pairs.size:
        d = X[pairs[:, 0]] - X[pairs[:, 1]]
        r2 = np.zeros(len(pairs))
        for j in range(X.shape[1]):  # same accumulation order as sklearn's rdist
            r2 = r2 + d[:, j] * d[:, j]
        pairs = pairs[r2 <= float(eps) * float(eps)]
    deg = np.ones(n, dtype=np.int64)  # self
    if pairs.size:
        deg += np.bincount(pairs[:, 0], minlength=n) + np.bincount(pairs[:, 1], minlength=n)
    core = deg >= min_samples
    labels = np.full(n, -1, dtype=np.int64)
    if not core.any():
        return labels
    cc_pairs = pairs[core[pairs[:, 0]] & core[pairs[:, 1]]] if pairs.size else np.empty((0, 2), dtype=np.int64)
    g = coo_matrix((np.ones(len(cc_pairs), dtype=np.int8), (cc_pairs[:, 0], cc_pairs[:, 1])), shape=(n, n))
    _, comp = connected_components(g, directed=False)
    core_idx = np.flatnonzero(core)
    # number the components by their smallest core index
    first = {}
    for i in core_idx:
        c = comp[i]
        if c not in first:
            first[c] = len(first)
    labels[core_idx] = np.array([first[comp[i]] for i in core_idx], dtype=np.int64)
    # border points: smallest cluster id among core neighbours
    if pairs.size:
        a, b = pairs[:, 0], pairs[:, 1]
        for src, dst in ((a, b), (b, a)):
            m = core[src] & ~core[dst]
            if m.any():
                cand = np.full(n, np.iinfo(np.int64).max, dtype=np.int64)
                np.minimum.at(cand, dst[m], labels[src[m]])
                upd = cand != np.iinfo(np.int64).max
                cur = labels[upd]
                labels[upd] = np.where(cur < 0, cand[upd], np.minimum(cur, cand[upd]))
    return labels


# sklearn.neighbors.KDTree.query_radius(count_only=True): inclusive `<= r` on the fp64 reduced
# distance (sklearn/neighbors/_binary_tree.pxi.tp).  call sites: app_simplified.py:269-281,
# models/crowd_flow_model.py:205-228.
def radius_count(centres: np.ndarray, queries: np.ndarray, r: float) -> np.ndarray:
    centres = np.asarray(centres, dtype=np.float64)
    queries = np.asarray(queries, dtype=np.float64)
    out = np.zeros(len(queries), dtype=np.int64)
    r2 = float(r) * float(r)
    for s in range(0, len(queries), 4096):
        q = queries[s:s + 4096]
        d2 = np.zeros((len(q), len(centres)))
        for j in range(centres.shape[1]):
            dj = q[:, j, None] - centres[None, :, j]
            d2 = d2 + dj * dj
        out[s:s + 4096] = (d2 <= r2).sum(1)
    return out


# KDTree(points).query_radius(points, r, count_only=True) of the visualisation paths
# (utils/visualization.py:43-45, 164-168; app_simplified.py:158-159): per point, the number of points
# (itself included) with fp64 reduced distance <= r*r.  Brute force in chunks.
def local_density_counts(points: np.ndarray, r: float) -> np.ndarray:
    pts = np.asarray(points, dtype=np.float64)
    out = np.zeros(len(pts), dtype=np.int64)
    r2 = float(r) * float(r)
    for s in range(0, len(pts), 512):
        q = pts[s:s + 512]
        d2 = np.zeros((len(q), len(pts)))
        for j in range(pts.shape[1]):
            dj = q[:, j, None] - pts[None, :, j]
            d2 = d2 + dj * dj
        out[s:s + 512] = (d2 <= r2).sum(1)
    return out
