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
    return list(cases)


def main():
    from utils.data_processing import load_lidar_data
    import warnings
    out = {}
    for name in write_cases():
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


if __name__ == "__main__":
    main()
