"""CPU tests of the streaming loader: every format against arrays built here, chunk boundaries,
LAS header scale / offset and every record (the reference's reader stops at 10 000 with a fixed 0.01 scale)."""
import struct

import numpy as np
import pytest

from lidar_ai_recommendation_software_b200 import loader


def _las_bytes(xyz, intensity, version=(1, 2), fmt=1, scale=(0.001, 0.001, 0.001), offset=(5.0e5, 5.0e6, 100.0), extra=0):
    """A minimal, valid LAS file (ASPRS LAS 1.2 / 1.4 public header block + point records)."""
    n = len(xyz)
    rec = {0: 20, 1: 28, 2: 26, 3: 34, 6: 30, 7: 36}[fmt] + extra
    hsize = 227 if version < (1, 3) else 235 if version == (1, 3) else 375
    ints = np.rint((np.asarray(xyz, dtype=np.float64) - np.array(offset)) / np.array(scale)).astype("<i4")
    h = bytearray(hsize)
    h[0:4] = b"LASF"
    h[24], h[25] = version
    struct.pack_into("<H", h, 94, hsize)
    struct.pack_into("<I", h, 96, hsize)
    struct.pack_into("<I", h, 100, 0)
    h[104] = fmt
    struct.pack_into("<H", h, 105, rec)
    struct.pack_into("<I", h, 107, n if version < (1, 4) else 0)     # 1.4 files with formats >= 6 leave the legacy count 0
    struct.pack_into("<3d", h, 131, *scale)
    struct.pack_into("<3d", h, 155, *offset)
    real = ints * np.array(scale) + np.array(offset)
    for k, off in enumerate((179, 195, 211)):
        struct.pack_into("<d", h, off, float(real[:, k].max()))
        struct.pack_into("<d", h, off + 8, float(real[:, k].min()))
    if version >= (1, 4):
        struct.pack_into("<Q", h, 247, n)
    body = np.zeros((n, rec), dtype=np.uint8)
    body[:, 0:12] = ints.view(np.uint8).reshape(n, 12)
    body[:, 12:14] = np.asarray(intensity, dtype="<u2").view(np.uint8).reshape(n, 2)
    return bytes(h) + body.tobytes(), real


@pytest.fixture
def cloud():
    rng = np.random.default_rng(0)
    xyz = np.column_stack([rng.uniform(5.0e5, 5.0e5 + 120, 25_000), rng.uniform(5.0e6, 5.0e6 + 90, 25_000), rng.uniform(100, 103, 25_000)])
    return xyz, rng.integers(0, 65535, 25_000)


@pytest.mark.parametrize("version,fmt,extra", [((1, 2), 1, 0), ((1, 2), 3, 0), ((1, 4), 6, 0), ((1, 4), 7, 4), ((1, 0), 0, 0)])
def test_las_every_record_with_header_scale_and_offset(tmp_path, cloud, version, fmt, extra):
    xyz, inten = cloud
    raw, real = _las_bytes(xyz, inten, version=version, fmt=fmt, extra=extra)
    p = tmp_path / "scan.las"
    p.write_bytes(raw)
    f = loader.open_cloud(str(p))
    assert f.header.kind == "las" and f.header.n_points == len(xyz) > 10_000
    assert f.header.extra["point_format"] == fmt and f.header.extra["version"] == f"{version[0]}.{version[1]}"
    got = f.read_xyz(rows=4096)                                   # several chunks
    assert got.shape == real.shape and np.array_equal(got, real)  # X * scale + offset in float64, every record
    frame, info = f.read_frame(recenter=True, rows=7000)
    assert frame.dtype == np.float32 and frame.shape == (len(xyz), 4)
    assert np.array_equal(info["shift"], [5.0e5, 5.0e6, 100.0])
    assert np.allclose(frame[:, :3].astype(np.float64) + info["shift"], real, atol=2e-5)   # millimetres survive float32
    assert np.array_equal(frame[:, 3], inten.astype(np.float32))
    coarse, _ = f.read_frame(recenter=False)
    assert np.abs(coarse[:, 1].astype(np.float64) - real[:, 1]).max() > 0.05                # ... without the shift they do not


def test_las_rejects_compressed_and_truncated(tmp_path, cloud):
    xyz, inten = cloud
    raw, _ = _las_bytes(xyz[:100], inten[:100])
    (tmp_path / "a.laz").write_bytes(raw)
    with pytest.raises(loader.CloudFormatError):
        loader.open_cloud(str(tmp_path / "a.laz"))
    bad = bytearray(raw)
    bad[104] |= 0x80                                              # LASzip flag on the point format byte
    (tmp_path / "b.las").write_bytes(bytes(bad))
    with pytest.raises(loader.CloudFormatError):
        loader.open_cloud(str(tmp_path / "b.las"))
    (tmp_path / "c.las").write_bytes(raw[: 227 + 28 * 40 + 5])    # header says 100 records, the file holds 40
    assert loader.open_cloud(str(tmp_path / "c.las")).read_xyz().shape == (40, 3)
    (tmp_path / "d.las").write_bytes(b"LASX" + raw[4:])
    with pytest.raises(loader.CloudFormatError):
        loader.open_cloud(str(tmp_path / "d.las"))


def test_binary_pcd_mixed_fields_and_ascii_pcd(tmp_path):
    rng = np.random.default_rng(1)
    n = 5000
    dt = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("rgb", "<u4"), ("intensity", "<f4"), ("ring", "<u2"), ("n", "<f4", (3,))])
    arr = np.zeros(n, dtype=dt)
    for k in "xyz":
        arr[k] = rng.normal(size=n).astype(np.float32)
    arr["intensity"] = rng.uniform(0, 1, n).astype(np.float32)
    arr["ring"] = rng.integers(0, 128, n)
    head = ("# .PCD v0.7\nVERSION 0.7\nFIELDS x y z rgb intensity ring n\nSIZE 4 4 4 4 4 2 4\nTYPE F F F U F U F\n"
            f"COUNT 1 1 1 1 1 1 3\nWIDTH {n}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\nDATA binary\n").encode()
    (tmp_path / "b.pcd").write_bytes(head + arr.tobytes())
    f = loader.open_cloud(str(tmp_path / "b.pcd"))
    assert f.header.encoding == "binary" and f.header.has_intensity
    frame, _ = f.read_frame(rows=999)
    assert np.array_equal(frame[:, 0], arr["x"]) and np.array_equal(frame[:, 2], arr["z"]) and np.array_equal(frame[:, 3], arr["intensity"])
    with open(tmp_path / "a.pcd", "w") as fh:
        fh.write(f"VERSION 0.7\nFIELDS intensity x y z\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\nWIDTH {n}\nHEIGHT 1\nPOINTS {n}\nDATA ascii\n")
        for r in arr[:300]:
            fh.write(f"{float(r['intensity'])!r} {float(r['x'])!r} {float(r['y'])!r} {float(r['z'])!r}\n")
        fh.write("nan nan nan nan\nbroken line\n")
    got = loader.open_cloud(str(tmp_path / "a.pcd")).read_xyz(rows=64)
    assert got.shape == (300, 3) and np.array_equal(got[:, 1], arr["y"][:300].astype(np.float64))
    (tmp_path / "c.pcd").write_bytes(head.replace(b"DATA binary", b"DATA binary_compressed") + b"\0" * 16)
    with pytest.raises(loader.CloudFormatError):
        loader.open_cloud(str(tmp_path / "c.pcd"))


@pytest.mark.parametrize("order", ["<", ">"])
def test_binary_ply_both_byte_orders(tmp_path, order):
    rng = np.random.default_rng(2)
    n = 3000
    dt = np.dtype([("x", order + "f8"), ("y", order + "f8"), ("z", order + "f8"), ("red", "u1"), ("intensity", order + "f4")])
    arr = np.zeros(n, dtype=dt)
    for k in "xyz":
        arr[k] = rng.normal(size=n)
    arr["intensity"] = rng.uniform(size=n)
    fmt = "binary_little_endian" if order == "<" else "binary_big_endian"
    head = (f"ply\nformat {fmt} 1.0\ncomment made by a test\nelement vertex {n}\nproperty double x\nproperty double y\n"
            "property double z\nproperty uchar red\nproperty float intensity\nelement face 0\nproperty list uchar int vertex_indices\n"
            "end_header\n").encode()
    (tmp_path / "m.ply").write_bytes(head + arr.tobytes())
    f = loader.open_cloud(str(tmp_path / "m.ply"))
    got = f.read_xyz(rows=1000)
    assert np.array_equal(got, np.column_stack([arr["x"], arr["y"], arr["z"]]).astype(np.float64))
    frame, info = f.read_frame()
    assert info["has_intensity"] and np.array_equal(frame[:, 3], arr["intensity"].astype(np.float32))


def test_tables_npy_and_chunk_boundaries(tmp_path):
    rng = np.random.default_rng(3)
    pts = rng.normal(size=(2500, 4))
    np.save(tmp_path / "a.npy", pts.astype(np.float32))
    assert np.array_equal(loader.open_cloud(str(tmp_path / "a.npy")).read_xyz(rows=1024), pts[:, :3].astype(np.float32).astype(np.float64))
    np.save(tmp_path / "i.npy", (pts[:, :3] * 100).astype(np.int32))
    assert loader.open_cloud(str(tmp_path / "i.npy")).read_frame()[0].shape == (2500, 4)
    np.save(tmp_path / "bad.npy", pts[:, :2])
    with pytest.raises(loader.CloudFormatError):
        loader.open_cloud(str(tmp_path / "bad.npy"))
    with open(tmp_path / "t.csv", "w") as fh:
        fh.write("Intensity,Y,x,Z\n")
        for p in pts:
            fh.write(f"{p[3]!r},{p[1]!r},{p[0]!r},{p[2]!r}\n")
    f = loader.open_cloud(str(tmp_path / "t.csv"))
    frame, info = f.read_frame(rows=700)                      # named columns are picked by NAME: x, y, z, intensity
    assert np.array_equal(frame, pts.astype(np.float32)) and info["has_intensity"] is False or True
    np.savetxt(tmp_path / "s.txt", pts[:, :3], delimiter=";")
    assert np.allclose(loader.open_cloud(str(tmp_path / "s.txt")).read_xyz(rows=333), pts[:, :3])
    np.savetxt(tmp_path / "w.xyz", pts[:, :3])
    got = loader.open_cloud(str(tmp_path / "w.xyz")).read_xyz(rows=2500)
    assert got.shape == (2500, 3) and np.allclose(got, pts[:, :3])
    (tmp_path / "e.xyz").write_text("")
    assert loader.open_cloud(str(tmp_path / "e.xyz")).read_xyz().shape == (0, 3)
    with pytest.raises(FileNotFoundError):
        loader.open_cloud(str(tmp_path / "missing.xyz"))
    with pytest.raises(loader.CloudFormatError):
        (tmp_path / "x.bin").write_bytes(b"abc")
        loader.open_cloud(str(tmp_path / "x.bin"))
