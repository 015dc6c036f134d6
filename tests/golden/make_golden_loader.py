"""Loader fixtures: small point-cloud files in every format the reference's
`load_lidar_data` (utils/data_processing.py:8-125) reads, and what the UNMODIFIED reference returns for them.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_loader.py      (build container only)

Writes tests/golden/loader/<name>.<ext> and tests/golden/loader/expected.npz.  A case the reference rejects is
stored as its exception message.
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
OUT = HERE / "loader"
sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True


def write_cases():
    OUT.mkdir(exist_ok=True)
    r = np.random.default_rng(12)
    pts = np.round(r.uniform(-20, 20, (40, 3)), 4)
    extra = np.round(r.uniform(0, 1, (40, 2)), 3)
    cases = {}
    # csv with named columns in a different order, plus an extra column
    with open(OUT / "named.csv", "w") as f:
        f.write("intensity,Y,x,z\n")
        for p, e in zip(pts, extra):
            f.write(f"{e[0]},{p[1]},{p[0]},{p[2]}\n")
    cases["named.csv"] = None
    # csv without x/y/z names: first three columns
    with open(OUT / "plain.csv", "w") as f:
        f.write("a,b,c,d\n")
        for p, e in zip(pts, extra):
            f.write(f"{p[0]},{p[1]},{p[2]},{e[1]}\n")
    cases["plain.csv"] = None
    # xyz / txt: whitespace separated, extra columns
    np.savetxt(OUT / "cloud.xyz", np.column_stack([pts, extra]), fmt="%.4f")
    cases["cloud.xyz"] = None
    np.savetxt(OUT / "cloud.txt", pts, fmt="%.6e")
    cases["cloud.txt"] = None
    # ascii pcd with a full header, a comment, a blank line and a short (2-field) line in the body
    with open(OUT / "cloud.pcd", "w") as f:
        f.write("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\n"
                "TYPE F F F F\nCOUNT 1 1 1 1\nWIDTH 40\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS 40\nDATA ascii\n")
        for i, (p, e) in enumerate(zip(pts, extra)):
            f.write(f"{p[0]} {p[1]} {p[2]} {e[0]}\n")
            if i == 7:
                f.write("\n1.0 2.0\n")
    cases["cloud.pcd"] = None
    # header-less pcd: data from the first line
    np.savetxt(OUT / "bare.pcd", pts[:9], fmt="%.4f")
    cases["bare.pcd"] = None
    # ascii ply: vertex count smaller than the body (faces follow), extra properties
    with open(OUT / "mesh.ply", "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex 25\nproperty float x\nproperty float y\nproperty float z\n"
                "property uchar red\nelement face 2\nproperty list uchar int vertex_indices\nend_header\n")
        for p in pts[:25]:
            f.write(f"{p[0]} {p[1]} {p[2]} 200\n")
        f.write("3 0 1 2\n3 2 3 4\n")
    cases["mesh.ply"] = None
    # npy with 5 columns
    np.save(OUT / "cloud.npy", np.column_stack([pts, extra]))
    cases["cloud.npy"] = None
    # rejected inputs
    (OUT / "empty.xyz").write_text("")
    cases["empty.xyz"] = None
    (OUT / "cloud.las").write_bytes(b"LASF" + bytes(64))
    cases["cloud.las"] = None
    # ---- inputs of the desktop shell's DataLoader (windows_implementation/core/data_loader.py) only
    import struct
    # a LAS 1.2-like file, point format 1 (28-byte records), 30 records + a truncated one that still holds X, Y, Z
    hdr = bytearray(227)
    hdr[0:4] = b"LASF"
    hdr[96:100] = struct.pack("<I", 227)          # offset to point data
    hdr[104] = 1                                  # point data format id
    hdr[105:107] = struct.pack("<H", 28)          # record length
    hdr[107:111] = struct.pack("<I", 31)          # number of point records (the reference reads it here)
    ints = np.round(pts[:31] * 100).astype("<i4")
    body = b"".join(struct.pack("<iii", *row) + bytes(16) for row in ints[:30]) + struct.pack("<iii", *ints[30]) + bytes(3)
    (OUT / "scan.las").write_bytes(bytes(hdr) + body)
    cases["scan.las"] = None
    (OUT / "short.las").write_bytes(bytes(hdr[:111]) + bytes(200))      # offset points past the end: no records
    cases["short.las"] = None
    (OUT / "scan.laz").write_bytes(b"LASF" + bytes(300))
    cases["scan.laz"] = None
    # semicolon-separated txt
    with open(OUT / "semi.txt", "w") as f:
        for p in pts[:12]:
            f.write(f"{p[0]};{p[1]};{p[2]};7\n")
    cases["semi.txt"] = None
    # pcd whose body has a non-numeric line; binary pcd; ply with a bad vertex line; binary ply; ply without x/y/z
    with open(OUT / "dirty.pcd", "w") as f:
        f.write("VERSION 0.7\nFIELDS x y z\nPOINTS 5\nDATA ascii\n")
        f.write("1 2 3\nfoo bar baz\n4 5 6 7\n\n7 8\n9 10 11\n")
    cases["dirty.pcd"] = None
    with open(OUT / "bin.pcd", "wb") as f:
        f.write(b"VERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 4\nHEIGHT 1\nPOINTS 4\nDATA binary\n")
        f.write(pts[:4].astype(np.float32).tobytes())
    cases["bin.pcd"] = None
    with open(OUT / "dirty.ply", "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex 5\nproperty float x\nproperty double y\nproperty float z\nend_header\n")
        f.write("1 2 3\nnan nan nan\nx y z\n4 5\n6 7 8 9\n10 11 12\n")
    cases["dirty.ply"] = None
    with open(OUT / "bin.ply", "wb") as f:
        f.write(b"ply\nformat binary_little_endian 1.0\nelement vertex 2\nproperty float x\nproperty float y\nproperty float z\nend_header\n")
        f.write(pts[:2].astype(np.float32).tobytes())
    cases["bin.ply"] = None
    with open(OUT / "noxyz.ply", "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex 1\nproperty float a\nproperty float b\nproperty float c\nend_header\n1 2 3\n")
    cases["noxyz.ply"] = None
    (OUT / "two.csv").write_text("a,b\n1,2\n3,4\n")
    cases["two.csv"] = None
    return list(cases)


def main():
    from utils.data_processing import load_lidar_data
    import warnings
    out = {}
    legacy = ["named.csv", "plain.csv", "cloud.xyz", "cloud.txt", "cloud.pcd", "bare.pcd", "mesh.ply", "cloud.npy",
              "empty.xyz", "cloud.las"]
    write_cases()
    for name in legacy:
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                got = np.asarray(load_lidar_data(str(OUT / name)), dtype=np.float64)
            out[name] = got
            print(name, got.shape)
        except Exception as e:
            msg = str(e).replace(str(OUT), "<dir>")
            out[name + ":error"] = np.array(msg)
            print(name, "raises:", msg)
    np.savez_compressed(OUT / "expected.npz", **out)
    # ---- the desktop shell's DataLoader on every file it has a reader for
    import json
    import logging
    logging.disable(logging.CRITICAL)
    sys.path.insert(0, "/root/reference/windows_implementation")
    from core.data_loader import DataLoader
    out2 = {}
    for name in sorted(p.name for p in OUT.iterdir() if p.suffix not in (".npz", ".npy")) + ["missing.xyz"]:
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                ds = DataLoader().load_file(str(OUT / name))
            out2[name] = np.asarray(ds.points, dtype=np.float64)
            meta = {k: v for k, v in ds.metadata.items() if k != "file_path"}
            out2[name + ":meta"] = np.array(json.dumps(meta, sort_keys=True))
            print("DataLoader", name, out2[name].shape)
        except Exception as e:
            out2[name + ":error"] = np.array(type(e).__name__ + ": " + str(e).replace(str(OUT), "<dir>"))
            print("DataLoader", name, "raises:", out2[name + ":error"])
    np.savez_compressed(OUT / "expected_dataloader.npz", **out2)


if __name__ == "__main__":
    main()
