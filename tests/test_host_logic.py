"""CPU tests of host-side logic that needs no GPU: loader contract, percentile lerp, plane solve."""
import numpy as np
import pytest


def test_percentile_from_order_stats_matches_numpy():
    from lidar_ai_recommendation_software_b200.preprocess import percentile_from_order_stats
    rng = np.random.default_rng(0)
    for n in (1, 2, 3, 10, 11, 1000, 7919):
        x = np.sort(rng.normal(size=n))
        lo = int(np.floor((n - 1) * 0.3))
        got = percentile_from_order_stats(x[lo], x[min(lo + 1, n - 1)], n, 30)
        assert got == float(np.percentile(x, 30))


def test_plane_from_sums_matches_lstsq():
    from lidar_ai_recommendation_software_b200.preprocess import _plane_from_sums
    rng = np.random.default_rng(1)
    g = rng.uniform(-50, 50, (5000, 3))
    g[:, 2] = 0.01 * g[:, 0] - 0.02 * g[:, 1] + 0.3 + rng.normal(0, 0.01, 5000)
    c = g.mean(0) + 0.5
    d = g - c
    s = np.array([len(g), d[:, 0].sum(), d[:, 1].sum(), d[:, 2].sum(), (d[:, 0] ** 2).sum(), (d[:, 0] * d[:, 1]).sum(),
                  (d[:, 1] ** 2).sum(), (d[:, 0] * d[:, 2]).sum(), (d[:, 1] * d[:, 2]).sum(), 0.0])
    A = np.column_stack((g[:, 0], g[:, 1], np.ones(len(g))))
    want = np.linalg.lstsq(A, g[:, 2], rcond=None)[0]
    got = _plane_from_sums(s, c)
    assert np.allclose([got[0], got[1], got[3]], want, rtol=1e-9) and got[2] == -1


def test_loader_formats(tmp_path):
    from lidar_ai_recommendation_software_b200.io import load_lidar_data
    pts = np.random.default_rng(0).normal(size=(50, 3))
    np.save(tmp_path / "a.npy", np.column_stack([pts, np.ones(50)]))
    assert np.array_equal(load_lidar_data(str(tmp_path / "a.npy")), pts)
    np.savetxt(tmp_path / "a.xyz", pts)
    assert np.allclose(load_lidar_data(str(tmp_path / "a.xyz")), pts)
    np.savetxt(tmp_path / "a.txt", pts)
    assert np.allclose(load_lidar_data(str(tmp_path / "a.txt")), pts)
    with open(tmp_path / "a.csv", "w") as f:
        f.write("intensity,Y,x,z\n")
        for p in pts:
            f.write(f"1.0,{float(p[1])!r},{float(p[0])!r},{float(p[2])!r}\n")
    got = load_lidar_data(str(tmp_path / "a.csv"))       # columns in file order: Y, x, z (reference behaviour)
    assert np.allclose(got, pts[:, [1, 0, 2]])
    with open(tmp_path / "a.pcd", "w") as f:
        f.write("# .PCD v0.7\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 50\nHEIGHT 1\n"
                "POINTS 50\nDATA ascii\n")
        for p in pts:
            f.write(f"{float(p[0])!r} {float(p[1])!r} {float(p[2])!r}\n")
    assert np.allclose(load_lidar_data(str(tmp_path / "a.pcd")), pts)
    with open(tmp_path / "b.pcd", "wb") as f:
        f.write(b"VERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\nWIDTH 50\nHEIGHT 1\n"
                b"POINTS 50\nDATA binary\n")
        f.write(np.column_stack([pts, np.ones(50)]).astype(np.float32).tobytes())
    assert np.allclose(load_lidar_data(str(tmp_path / "b.pcd")), pts.astype(np.float32))
    with open(tmp_path / "a.ply", "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex 50\nproperty float x\nproperty float y\nproperty float z\n"
                "element face 0\nend_header\n")
        for p in pts:
            f.write(f"{float(p[0])!r} {float(p[1])!r} {float(p[2])!r}\n")
    assert np.allclose(load_lidar_data(str(tmp_path / "a.ply")), pts)
    with pytest.raises(Exception, match="Failed to load point cloud file"):
        load_lidar_data(str(tmp_path / "nope.xyz"))
    with pytest.raises(Exception, match="Unsupported file format"):
        load_lidar_data(str(tmp_path / "a.bin"))


def test_loader_matches_reference_fixtures():
    """Every format the reference's load_lidar_data reads (utils/data_processing.py:8-125): the arrays the UNMODIFIED
    reference returned for the files under tests/golden/loader (made by tests/golden/make_golden_loader.py), bit for
    bit; inputs the reference rejects must be rejected with the same wrapped message."""
    from pathlib import Path
    import warnings
    from lidar_ai_recommendation_software_b200.io import load_lidar_data
    d = Path(__file__).resolve().parent / "golden" / "loader"
    exp = np.load(d / "expected.npz", allow_pickle=False)
    assert len(exp.files) == 10
    for key in exp.files:
        if key.endswith(":error"):
            name = key[: -len(":error")]
            with pytest.raises(Exception) as ei, warnings.catch_warnings():
                warnings.simplefilter("ignore")
                load_lidar_data(str(d / name))
            assert str(ei.value).startswith("Failed to load point cloud file:")
            if name.endswith(".las"):
                assert str(ei.value) == str(exp[key])
        else:
            got = np.asarray(load_lidar_data(str(d / key)), dtype=np.float64)
            assert got.shape == exp[key].shape and np.array_equal(got, exp[key]), key


def test_windows_data_loader_matches_reference_fixtures():
    """The desktop shell's DataLoader (windows_implementation/core/data_loader.py:30-447): points, metadata and
    exceptions of the UNMODIFIED reference for every fixture file (tests/golden/make_golden_loader.py), reproduced by
    `DataLoader(reference_compat=True)`; the default mode reads the same files through the streaming loader."""
    import json
    import logging
    import warnings
    from pathlib import Path
    from lidar_ai_recommendation_software_b200.windows_core.data_loader import DataLoader, Dataset
    logging.disable(logging.CRITICAL)
    try:
        d = Path(__file__).resolve().parent / "golden" / "loader"
        exp = np.load(d / "expected_dataloader.npz", allow_pickle=False)
        names = sorted({k.split(":")[0] for k in exp.files})
        assert len(names) == 20
        for name in names:
            if name + ":error" in exp.files:
                with pytest.raises(Exception) as ei, warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    DataLoader(reference_compat=True).load_file(str(d / name))
                got = type(ei.value).__name__ + ": " + str(ei.value).replace(str(d), "<dir>")
                assert got == str(exp[name + ":error"]), name
            else:
                ds = DataLoader(reference_compat=True).load_file(str(d / name))
                assert isinstance(ds, Dataset)
                pts = np.asarray(ds.points, dtype=np.float64)
                assert pts.shape == exp[name].shape and np.array_equal(pts, exp[name], equal_nan=True), name
                assert ds.metadata["file_path"] == str(d / name)
                meta = {k: v for k, v in ds.metadata.items() if k != "file_path"}
                assert json.loads(json.dumps(meta, sort_keys=True)) == json.loads(str(exp[name + ":meta"])), name
        # default mode: same rows for everything the reference reads correctly, and the files it cannot read
        for name in ("cloud.pcd", "cloud.xyz", "cloud.txt", "semi.txt", "named.csv", "plain.csv", "mesh.ply", "dirty.pcd", "dirty.ply"):
            ds = DataLoader().load_file(str(d / name))
            assert np.array_equal(np.asarray(ds.points, dtype=np.float64), exp[name], equal_nan=True), name
        for name in ("bin.pcd", "bin.ply"):
            ds = DataLoader().load_file(str(d / name))               # refused by the reference, read here
            assert ds.points.shape[1] == 3 and ds.metadata["point_count"] == len(ds.points) > 0
        full = DataLoader().load_file(str(d / "scan.las"))            # header scale / offset instead of the fixed 0.01
        # (the fixture's last record is cut short on purpose: 30 whole records on disk under a header that says 31)
        assert full.metadata["total_points"] == 30 and len(full.points) == 30 and "scale" in full.metadata
    finally:
        logging.disable(logging.NOTSET)


def test_hotspot_selection_is_upstreams_stable_sort():
    """The top-5 hotspot rule of both density surfaces (models/crowd_density_model.py:93-105,
    app_simplified.py:289-305): one dict per cell over the threshold, stable sort by density descending, first
    five.  The package picks them with a stable argsort of the negated densities: same cells, same order, ties
    included."""
    rng = np.random.default_rng(7)
    for _ in range(500):
        v = rng.integers(0, 6, size=int(rng.integers(0, 60))).astype(np.float64) / 4
        thr = max(0.5, float(rng.uniform(0, 1.5)))
        cand = np.where(v >= thr)[0]
        ref = sorted([{"i": int(i), "d": v[i]} for i in cand], key=lambda h: h["d"], reverse=True)[:5]
        top = cand[np.argsort(-v[cand], kind="stable")[:5]]
        assert [h["i"] for h in ref] == [int(i) for i in top]


def test_parallel_host_memcpy_both_store_kinds():
    """lidar_host_memcpy (the staging copy of the host-buffer entries) with plain and with non-temporal stores: every
    byte arrives, nothing outside the destination range is touched, whatever the alignment of either side.  Host code
    only -- no device is needed."""
    from lidar_ai_recommendation_software_b200 import _capi
    lib = _capi.lib
    rng = np.random.default_rng(5)
    src = rng.integers(0, 256, (6 << 20) + 333, dtype=np.uint8)
    dst = np.zeros(src.size + 256, dtype=np.uint8)
    try:
        for nt in (0, 1):
            assert lib.lidar_host_copy_nontemporal(nt) == 0
            for threads in (1, 4):
                assert lib.lidar_host_copy_threads(threads) == 0
                for d_off, s_off, n in ((0, 0, src.size), (1, 3, src.size - 64), (63, 17, (1 << 20) + 5), (7, 0, 1000), (5, 5, 0)):
                    dst[:] = 0xEE
                    assert lib.lidar_host_memcpy(dst.ctypes.data + d_off, src.ctypes.data + s_off, n) == 0
                    assert np.array_equal(dst[d_off:d_off + n], src[s_off:s_off + n])
                    assert (dst[:d_off] == 0xEE).all() and (dst[d_off + n:] == 0xEE).all()
    finally:
        lib.lidar_host_copy_nontemporal(1)
        lib.lidar_host_copy_threads(2)


def test_result_pool_reserve_and_reuse():
    """ResultPool (ops.py): reserved buffers are handed out without a new allocation, a buffer is reused only when no array
    references it any more, and nothing a caller still holds is ever handed out again.  Pageable buffers: no device needed."""
    from lidar_ai_recommendation_software_b200 import ops
    pool = ops.ResultPool(max_buffers=3, pinned=False)
    pool.reserve(2, 1 << 16)
    assert len(pool.bufs) == 2
    ids = {id(b) for b in pool.bufs}
    a = pool.take(1000)
    view = np.frombuffer(a, dtype=np.int32, count=10)         # what collect() hands to the caller: a view of the buffer
    view[:] = 7
    del a
    b = pool.take(1000)
    assert id(b) in ids and not np.shares_memory(b, view)     # the second reserved buffer, not the one still in use
    c = pool.take(1000)                                       # both in use: a third one is allocated (and kept)
    assert id(c) not in ids and len(pool.bufs) == 3
    d = pool.take(1000)                                       # pool full and all in use: a buffer outside the pool
    assert not any(np.shares_memory(d, x) for x in (view, b, c)) and len(pool.bufs) == 3
    assert (view == 7).all()
    del view, d
    e = pool.take(1000)                                       # the first buffer is free again
    assert id(e) in ids
    big = pool.take(1 << 20)                                  # larger than anything pooled: a new buffer
    assert big.nbytes >= (1 << 20)
