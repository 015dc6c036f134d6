"""load_lidar_data — the step before the hot path.

Same contract as the reference loader (utils/data_processing.py:8-125): format chosen by extension
(csv, xyz, pcd, ply, txt, npy), returns an (n,3) array of x,y,z, raises
`Exception("Failed to load point cloud file: ...")` on any problem.  The ASCII bodies are parsed by
a vectorised tokenizer instead of the reference's per-line Python loop, and binary PCD (DATA binary,
float32/float64 fields) is accepted in addition.
"""
from __future__ import annotations

import io as _io

import numpy as np
import pandas as pd


def _ascii_rows_exact(text: str, max_rows: int | None = None) -> np.ndarray:
    """The reference's own loop (utils/data_processing.py:66-74, 96-105): the first three fields of every
    non-empty line that has at least three; `max_rows` limits the LINES looked at, like its `range`."""
    lines = text.splitlines()
    if max_rows is not None:
        lines = lines[:max_rows]
    rows = []
    for line in lines:
        line = line.strip()
        if line:
            values = line.split()
            if len(values) >= 3:
                rows.append([float(v) for v in values[:3]])
    return np.array(rows)


def _ascii_rows(text: str, max_rows: int | None = None) -> np.ndarray:
    """First three whitespace-separated numeric columns of every non-empty line with >= 3 fields: a vectorised
    tokenizer for the regular case, the reference's loop whenever a line is short, blank inside a counted range, or
    holds something pandas would read as missing (so the result is the reference's in every case)."""
    if not text.strip():
        return np.empty((0, 3))
    body = text if max_rows is None else "\n".join(text.splitlines()[:max_rows])
    regular = max_rows is None or "\n\n" not in body.strip("\n")
    if regular:
        try:
            df = pd.read_csv(_io.StringIO(body), sep=r"\s+", header=None, usecols=[0, 1, 2], dtype=np.float64,
                             engine="c", skip_blank_lines=True)
            arr = df.to_numpy(dtype=np.float64)
            if not np.isnan(arr).any():
                return arr
        except Exception:
            pass
    return _ascii_rows_exact(text, max_rows)


def _load_pcd(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        raw = f.read()
    pos, header, data_kind = 0, {}, "ascii"
    while pos < len(raw):
        end = raw.find(b"\n", pos)
        end = len(raw) if end < 0 else end
        line = raw[pos:end].decode("ascii", "replace").strip()
        if line and not line.startswith("#"):
            parts = line.split()
            if parts[0].isupper() and len(parts) >= 2 and not _is_number(parts[0]):
                header[parts[0]] = parts[1:]
                if parts[0] == "DATA":
                    data_kind = parts[1].lower()
                    pos = end + 1
                    break
            else:
                break          # first data line of a header-less / DATA-less file (reference behaviour)
        pos = end + 1
    body = raw[pos:]
    if data_kind == "binary":
        sizes = [int(s) for s in header["SIZE"]]
        types = header["TYPE"]
        counts = [int(c) for c in header.get("COUNT", ["1"] * len(sizes))]
        fields = header["FIELDS"]
        dt = []
        for name, s, t, c in zip(fields, sizes, types, counts):
            code = {"F": "f", "I": "i", "U": "u"}[t] + str(s)
            dt.append((name, code) if c == 1 else (name, code, (c,)))
        n = int(header["POINTS"][0])
        arr = np.frombuffer(body, dtype=np.dtype(dt), count=n)
        return np.stack([arr[fields[0]], arr[fields[1]], arr[fields[2]]], 1).astype(np.float64)
    return _ascii_rows(body.decode("ascii", "replace"))


def _is_number(tok: str) -> bool:
    try:
        float(tok)
        return True
    except ValueError:
        return False


def _load_ply(path: str) -> np.ndarray:
    with open(path, "r") as f:
        text = f.read()
    n_points, start = None, 0
    offset = 0
    for line in text.splitlines(keepends=True):
        offset += len(line)
        if line.strip() == "end_header":
            start = offset
            break
        if "element vertex" in line:
            n_points = int(line.split()[-1])
    return _ascii_rows(text[start:], max_rows=n_points)


def load_lidar_data(file_path):
    """Load a point cloud file into an (n,3) array of x,y,z (utils/data_processing.py:8-125)."""
    try:
        ext = file_path.lower().split(".")[-1]
        if ext == "csv":
            data = pd.read_csv(file_path)
            xyz = [c for c in data.columns if str(c).lower() in ("x", "y", "z")]
            points = data[xyz[:3]].values if len(xyz) >= 3 else data.iloc[:, :3].values
        elif ext in ("xyz", "txt"):
            points = np.loadtxt(file_path, delimiter=None)[:, :3]
        elif ext == "pcd":
            points = _load_pcd(file_path)
        elif ext == "ply":
            points = _load_ply(file_path)
        elif ext == "npy":
            points = np.load(file_path)[:, :3]
        else:
            raise ValueError(f"Unsupported file format: {ext}")
        if len(points) == 0:
            raise ValueError("The loaded point cloud contains no points")
        return points
    except Exception as e:  # the apps show str(e) (app.py:103-104)
        raise Exception(f"Failed to load point cloud file: {str(e)}")
