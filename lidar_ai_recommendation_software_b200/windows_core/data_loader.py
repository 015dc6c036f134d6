"""Mirror of the desktop shell's loader, `windows_implementation/core/data_loader.py:15-447`.

Same classes (`Dataset`, `DataLoader`), same dispatch by extension, same accepted inputs, same metadata keys and the
same exceptions — pinned to the unmodified reference by the fixtures under tests/golden/loader.  The bodies are
parsed with vectorised readers (pandas / numpy) where that provably gives the reference's rows, and with the
reference's own per-line rule (first three fields of a line, lines that do not convert are skipped) otherwise.
`Dataset.points` is what `windows_core.run_analysis` takes.
"""
from __future__ import annotations

import io as _io
import logging
import os
import struct

import numpy as np
import pandas as pd

logger = logging.getLogger(__name__)


class Dataset:
    """Represents a dataset with point cloud data (core/data_loader.py:15-27)."""

    def __init__(self, points, metadata=None):
        self.points = points
        self.metadata = metadata or {}


def _rows_skip_invalid(lines) -> list:
    """core/data_loader.py:201-209 / 332-339: first three fields of every line with >= 3 fields; lines whose
    fields do not convert to float are skipped."""
    out = []
    for line in lines:
        values = line.strip().split()
        if len(values) >= 3:
            try:
                out.append([float(values[0]), float(values[1]), float(values[2])])
            except ValueError:
                continue
    return out


def _ascii_block(lines) -> np.ndarray:
    """Rows of a block of text lines: one vectorised parse when every line is a regular numeric row, the
    reference's per-line rule otherwise."""
    text = "".join(lines)
    if text.strip():
        try:
            df = pd.read_csv(_io.StringIO(text), sep=r"\s+", header=None, usecols=[0, 1, 2], dtype=np.float64,
                             engine="c", skip_blank_lines=True)
            arr = df.to_numpy(dtype=np.float64)
            n_candidates = sum(1 for ln in lines if len(ln.split()) >= 3)
            if not np.isnan(arr).any() and len(arr) == n_candidates:
                return arr
        except Exception:
            pass
    rows = _rows_skip_invalid(lines)
    return np.array(rows, dtype=float) if rows else np.empty((0, 3))


class DataLoader:
    """Loads and parses various LiDAR data formats (core/data_loader.py:30-447)."""

    def __init__(self):
        pass

    def load_file(self, file_path):
        if not os.path.exists(file_path):
            raise FileNotFoundError(f"File not found: {file_path}")
        ext = os.path.splitext(file_path)[1].lower()
        if ext == ".csv":
            return self._load_csv(file_path)
        elif ext == ".xyz" or ext == ".txt":
            return self._load_xyz(file_path)
        elif ext == ".pcd":
            return self._load_pcd(file_path)
        elif ext == ".ply":
            return self._load_ply(file_path)
        elif ext == ".las" or ext == ".laz":
            return self._load_las(file_path)
        else:
            raise ValueError(f"Unsupported file format: {ext}")

    def _load_csv(self, file_path):
        try:
            headers = pd.read_csv(file_path, nrows=0).columns.tolist()
            x_col, y_col, z_col = None, None, None
            for header in headers:
                low = header.lower()
                if low == "x":
                    x_col = header
                elif low == "y":
                    y_col = header
                elif low == "z":
                    z_col = header
            if x_col and y_col and z_col:
                df = pd.read_csv(file_path, usecols=[x_col, y_col, z_col])
                points = df[[x_col, y_col, z_col]].values
            else:
                df = pd.read_csv(file_path)
                if len(df.columns) >= 3:
                    points = df.iloc[:, :3].values
                else:
                    raise ValueError("CSV file doesn't have at least 3 columns for X, Y, Z coordinates")
            metadata = {"file_format": "csv", "file_path": file_path, "point_count": len(points), "columns": headers}
            return Dataset(points, metadata)
        except Exception as e:
            logger.error(f"Error loading CSV file: {str(e)}")
            raise

    def _load_xyz(self, file_path):
        try:
            with open(file_path, "r") as f:
                first_line = f.readline().strip()
            if "," in first_line:
                delimiter = ","
            elif ";" in first_line:
                delimiter = ";"
            else:
                delimiter = None
            points = np.loadtxt(file_path, delimiter=delimiter)
            if points.shape[1] > 3:
                points = points[:, :3]
            metadata = {"file_format": "xyz", "file_path": file_path, "point_count": len(points), "delimiter": delimiter}
            return Dataset(points, metadata)
        except Exception as e:
            logger.error(f"Error loading XYZ/TXT file: {str(e)}")
            raise

    def _load_pcd(self, file_path):
        try:
            with open(file_path, "rb") as f:
                lines = [ln.decode("utf-8", errors="ignore") for ln in f]
            header = {}
            data_start = None
            for i, line in enumerate(lines):
                if line.startswith("#"):
                    continue
                if line.strip() == "DATA ascii":
                    data_start = i + 1
                    break
                if line.strip() == "DATA binary":
                    raise ValueError("Binary PCD format not supported by this implementation")
                parts = line.strip().split()
                if len(parts) >= 2:
                    header[parts[0].lower()] = " ".join(parts[1:])
            points_array = _ascii_block(lines[data_start:]) if data_start is not None else np.empty((0, 3))
            if len(points_array) == 0:
                raise ValueError("No valid points found in PCD file")
            metadata = {"file_format": "pcd", "file_path": file_path, "point_count": len(points_array), "header": header}
            return Dataset(points_array, metadata)
        except Exception as e:
            logger.error(f"Error loading PCD file: {str(e)}")
            raise

    def _load_ply(self, file_path):
        try:
            with open(file_path, "rb") as f:
                lines = [ln.decode("utf-8", errors="ignore") for ln in f]
            vertex_count = 0
            x_prop = y_prop = z_prop = None
            data_format = "ascii"
            body_start = len(lines)
            for i, raw in enumerate(lines):
                line = raw.strip()
                if line == "end_header":
                    body_start = i + 1
                    break
                if line.startswith("format"):
                    parts = line.split()
                    if len(parts) >= 2:
                        data_format = parts[1]
                if line.startswith("element vertex"):
                    parts = line.split()
                    if len(parts) >= 3:
                        vertex_count = int(parts[2])
                if line.startswith("property float") or line.startswith("property double"):
                    parts = line.split()
                    if len(parts) >= 3:
                        name = parts[2].lower()
                        if name == "x":
                            x_prop = True
                        elif name == "y":
                            y_prop = True
                        elif name == "z":
                            z_prop = True
            if not (x_prop and y_prop and z_prop):
                raise ValueError("PLY file doesn't have valid X, Y, Z properties")
            if data_format != "ascii":
                raise ValueError(f"PLY format '{data_format}' not supported by this implementation")
            points_array = _ascii_block(lines[body_start: body_start + vertex_count])
            if len(points_array) == 0:
                raise ValueError("No valid points found in PLY file")
            metadata = {"file_format": "ply", "file_path": file_path, "point_count": len(points_array),
                        "vertex_count": vertex_count, "data_format": data_format}
            return Dataset(points_array, metadata)
        except Exception as e:
            logger.error(f"Error loading PLY file: {str(e)}")
            raise

    def _load_las(self, file_path):
        """The reference's simplified reader (core/data_loader.py:359-447), field offsets and all: format id at byte
        104, record length at 105, record count read at byte 107, data offset at 96, at most 10 000 records, a fixed
        0.01 scale.  The records are decoded in one `np.frombuffer` instead of three `struct.unpack` per point."""
        try:
            if file_path.lower().endswith(".laz"):
                raise ValueError("LAZ files require the laspy library with laszip support")
            with open(file_path, "rb") as f:
                raw = f.read()
            if raw[:4].decode() != "LASF":
                raise ValueError("Invalid LAS file signature")
            point_data_format_id = struct.unpack("<B", raw[104:105])[0]
            record_length = struct.unpack("<H", raw[105:107])[0]
            num_point_records = struct.unpack("<I", raw[107:111])[0]
            point_data_offset = struct.unpack("<I", raw[96:100])[0]
            want = min(num_point_records, 10000)
            body = raw[point_data_offset:]
            if record_length >= 12:
                have = min(want, len(body) // record_length)
                # a trailing partial record still counts if it holds X, Y, Z
                if have < want and len(body) - have * record_length >= 12:
                    tail = body[have * record_length: have * record_length + 12]
                else:
                    tail = b""
                rec = np.frombuffer(body, dtype=np.uint8, count=have * record_length).reshape(have, record_length)
                xyz = np.ascontiguousarray(rec[:, :12]).view("<i4").reshape(have, 3)
                if tail:
                    xyz = np.concatenate([xyz, np.frombuffer(tail, dtype="<i4").reshape(1, 3)])
                points_array = xyz.astype(float) * 0.01
            else:
                points_array = np.empty((0, 3))      # f.read(record_length) < 12 bytes: the reference stops at once
            if len(points_array) == 0:
                raise ValueError("No valid points found in LAS file")
            metadata = {"file_format": "las", "file_path": file_path, "point_count": len(points_array),
                        "point_data_format_id": point_data_format_id, "total_points": num_point_records}
            return Dataset(points_array, metadata)
        except Exception as e:
            logger.error(f"Error loading LAS/LAZ file: {str(e)}")
            if "LAZ files require the laspy library" in str(e):
                raise ValueError("LAZ files require additional libraries. Please install with: pip install laspy[laszip]")
            raise
