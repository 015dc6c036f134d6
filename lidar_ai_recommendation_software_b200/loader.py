"""Streaming point-cloud loader: the step before the hot path.

The reference parses ASCII bodies line by line in Python (utils/data_processing.py:43-113,
windows_implementation/core/data_loader.py:170-356) and its LAS reader is a toy — first 10 000 records, a fixed
0.01 scale, header fields read at the wrong offsets (core/data_loader.py:359-447).  This module reads what a venue
scan actually arrives as, in bounded chunks, straight into the `float4` (x, y, z, intensity) frame layout the device
kernels take — optionally into page-locked memory allocated through the C ABI, so the next step is one DMA:

    ASCII tables   .csv / .xyz / .txt, comma / semicolon / whitespace separated, optional header row
    PCD            DATA ascii and DATA binary (any mix of F/I/U fields, COUNT > 1 fields skipped over)
    PLY            ascii, binary_little_endian, binary_big_endian (vertex element; x/y/z of any scalar type)
    NPY            (n, >=3) arrays of any real dtype, memory-mapped
    LAS 1.0 – 1.4  point formats 0 – 10: every record, x = X * scale + offset from the header, intensity kept

`open_cloud(path)` sniffs the format and parses the header; `.read_frame()` returns the (n, 4) float32 frame,
`.read_xyz()` the reference's (n, 3) float64 array, `.chunks()` streams (k, 4) float64 blocks.  The drop-in
`windows_core.DataLoader` and `io.load_lidar_data` keep the reference's interfaces on top of it.
"""
from __future__ import annotations

import io as _io
import os
import struct
from dataclasses import dataclass, field
from typing import Iterator

import numpy as np

CHUNK_ROWS = 1 << 20

_PLY_TYPES = {
    "char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2",
    "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4",
    "double": "f8", "float64": "f8",
}
# LAS point data record formats: minimum record length (ASPRS LAS 1.4 R15, tables 7-17)
_LAS_MIN_RECORD = {0: 20, 1: 28, 2: 26, 3: 34, 4: 57, 5: 63, 6: 30, 7: 36, 8: 38, 9: 59, 10: 67}


class CloudFormatError(ValueError):
    """The file is not a point cloud this loader can read (bad signature, unsupported encoding, truncated)."""


@dataclass
class CloudHeader:
    kind: str                         # "table" | "pcd" | "ply" | "npy" | "las"
    encoding: str = "ascii"           # "ascii" | "binary" | "binary_big_endian" | "mmap"
    n_points: int | None = None       # None: unknown until the body has been read (ASCII tables)
    fields: tuple = ()                # names of the columns / fields found
    has_intensity: bool = False
    data_offset: int = 0
    extra: dict = field(default_factory=dict)


def _sniff(path: str) -> str:
    ext = os.path.splitext(path)[1].lower()
    with open(path, "rb") as f:
        magic = f.read(16)
    if magic[:4] == b"LASF":
        return "las"
    if magic[:6] == b"\x93NUMPY":
        return "npy"
    if magic[:3] == b"ply":
        return "ply"
    if ext in (".las", ".laz"):
        return "las"
    if ext == ".pcd":
        return "pcd"
    if ext == ".ply":
        return "ply"
    if ext == ".npy":
        return "npy"
    if ext in (".csv", ".xyz", ".txt", ".asc", ".pts"):
        return "table"
    raise CloudFormatError(f"Unsupported file format: {ext or path}")


class CloudFile:
    """One point-cloud file: header parsed on construction, body streamed on demand."""

    def __init__(self, path: str):
        if not os.path.exists(path):
            raise FileNotFoundError(f"File not found: {path}")
        self.path = path
        self.size = os.path.getsize(path)
        kind = _sniff(path)
        self.header: CloudHeader = getattr(self, f"_head_{kind}")()

    # ---- headers -------------------------------------------------------------------------------------
    def _head_table(self) -> CloudHeader:
        with open(self.path, "r", errors="replace") as f:
            first = ""
            for line in f:
                if line.strip() and not line.lstrip().startswith(("#", "//")):
                    first = line.strip()
                    break
        sep = "," if "," in first else ";" if ";" in first else None
        toks = [t for t in (first.split(sep) if sep else first.split()) if t != ""]
        names: tuple = ()
        try:
            [float(t) for t in toks]
        except ValueError:
            names = tuple(t.strip().strip('"') for t in toks)       # a header row
        return CloudHeader("table", "ascii", None, names, False, 0, {"separator": sep, "header_row": bool(names)})

    def _head_pcd(self) -> CloudHeader:
        meta: dict = {}
        offset, encoding = 0, "ascii"
        with open(self.path, "rb") as f:
            while True:
                line = f.readline()
                if not line:
                    break
                text = line.decode("ascii", "replace").strip()
                if not text or text.startswith("#"):
                    offset = f.tell()
                    continue
                parts = text.split()
                key = parts[0].upper()
                if key in ("VERSION", "FIELDS", "SIZE", "TYPE", "COUNT", "WIDTH", "HEIGHT", "VIEWPOINT", "POINTS", "DATA"):
                    meta[key] = parts[1:]
                    offset = f.tell()
                    if key == "DATA":
                        encoding = parts[1].lower() if len(parts) > 1 else "ascii"
                        break
                else:
                    break                       # header-less body: the first data line
        fields = tuple(x.lower() for x in meta.get("FIELDS", ("x", "y", "z")))
        n = int(meta["POINTS"][0]) if "POINTS" in meta else (
            int(meta["WIDTH"][0]) * int(meta.get("HEIGHT", ["1"])[0]) if "WIDTH" in meta else None)
        if encoding not in ("ascii", "binary"):
            raise CloudFormatError(f"PCD DATA {encoding} is not supported (ascii and binary are)")
        return CloudHeader("pcd", encoding, n, fields, "intensity" in fields, offset, {"pcd": meta})

    def _head_ply(self) -> CloudHeader:
        props: list[tuple[str, str]] = []
        encoding, n, in_vertex, first_element, offset = "ascii", None, False, None, 0
        with open(self.path, "rb") as f:
            if f.readline().strip() != b"ply":
                raise CloudFormatError("not a PLY file")
            while True:
                line = f.readline()
                if not line:
                    raise CloudFormatError("PLY header without end_header")
                parts = line.decode("ascii", "replace").split()
                if not parts:
                    continue
                if parts[0] == "format" and len(parts) >= 2:
                    encoding = {"ascii": "ascii", "binary_little_endian": "binary", "binary_big_endian": "binary_big_endian"}.get(parts[1])
                    if encoding is None:
                        raise CloudFormatError(f"PLY format '{parts[1]}' is not supported")
                elif parts[0] == "element" and len(parts) >= 3:
                    in_vertex = parts[1] == "vertex"
                    if first_element is None:
                        first_element = parts[1]
                    if in_vertex:
                        n = int(parts[2])
                elif parts[0] == "property" and in_vertex:
                    if parts[1] == "list":
                        raise CloudFormatError("PLY list properties inside the vertex element are not supported")
                    props.append((parts[-1].lower(), parts[1].lower()))
                elif parts[0] == "end_header":
                    offset = f.tell()
                    break
        names = tuple(p for p, _ in props)
        if n is None or not {"x", "y", "z"} <= set(names):
            raise CloudFormatError("PLY file doesn't have valid X, Y, Z properties")
        if encoding != "ascii" and first_element != "vertex":
            raise CloudFormatError("binary PLY whose first element is not the vertex list is not supported")
        return CloudHeader("ply", encoding, n, names, "intensity" in names or "scalar_intensity" in names, offset,
                           {"properties": props})

    def _head_npy(self) -> CloudHeader:
        a = np.load(self.path, mmap_mode="r", allow_pickle=False)
        if a.ndim != 2 or a.shape[1] < 3 or a.dtype.kind not in "fiu":
            raise CloudFormatError(f"NPY array of shape {a.shape} / dtype {a.dtype} is not an (n, >=3) point array")
        return CloudHeader("npy", "mmap", int(a.shape[0]), tuple("xyzw"[:min(a.shape[1], 4)]), a.shape[1] >= 4, 0,
                           {"dtype": str(a.dtype), "columns": int(a.shape[1])})

    def _head_las(self) -> CloudHeader:
        if self.path.lower().endswith(".laz"):
            raise CloudFormatError("LAZ (compressed LAS) needs a LASzip decoder; decompress to .las first")
        with open(self.path, "rb") as f:
            h = f.read(375)
        if len(h) < 227 or h[:4] != b"LASF":
            raise CloudFormatError("Invalid LAS file signature")
        major, minor = h[24], h[25]
        header_size, data_offset = struct.unpack_from("<HI", h, 94)
        fmt_byte, rec_len, legacy_n = struct.unpack_from("<BHI", h, 104)
        if fmt_byte & 0xC0:
            raise CloudFormatError("LAS point records are LASzip-compressed; decompress to plain .las first")
        fmt = fmt_byte & 0x3F
        scale = struct.unpack_from("<3d", h, 131)
        offs = struct.unpack_from("<3d", h, 155)
        n = legacy_n
        if (major, minor) >= (1, 4) and len(h) >= 255:
            n64 = struct.unpack_from("<Q", h, 247)[0]
            if n64:
                n = n64
        if fmt not in _LAS_MIN_RECORD or rec_len < _LAS_MIN_RECORD[fmt]:
            raise CloudFormatError(f"LAS point format {fmt} with record length {rec_len} is not valid")
        avail = max(0, (self.size - data_offset) // rec_len)
        if n == 0 or n > avail:
            n = avail                       # trust the bytes on disk over a stale header count
        return CloudHeader("las", "binary", int(n), ("x", "y", "z", "intensity"), True, int(data_offset),
                           {"version": f"{major}.{minor}", "point_format": int(fmt), "record_length": int(rec_len),
                            "scale": tuple(scale), "offset": tuple(offs), "header_size": int(header_size),
                            "bbox_min": (struct.unpack_from("<d", h, 187)[0], struct.unpack_from("<d", h, 203)[0],
                                         struct.unpack_from("<d", h, 219)[0]),
                            "bbox_max": (struct.unpack_from("<d", h, 179)[0], struct.unpack_from("<d", h, 195)[0],
                                         struct.unpack_from("<d", h, 211)[0])})

    # ---- bodies: every generator yields (k, 4) float64 blocks [x, y, z, intensity] ----------------------
    def chunks(self, rows: int = CHUNK_ROWS) -> Iterator[np.ndarray]:
        return getattr(self, f"_body_{self.header.kind}")(max(1, int(rows)))

    @staticmethod
    def _block(xyz: np.ndarray, w: np.ndarray | None) -> np.ndarray:
        out = np.zeros((xyz.shape[0], 4), dtype=np.float64)
        out[:, :3] = xyz
        if w is not None:
            out[:, 3] = w
        return out

    def _ascii_blocks(self, fh, rows: int, sep, cols, limit: int | None) -> Iterator[np.ndarray]:
        import pandas as pd
        kw = dict(header=None, comment="#", skip_blank_lines=True, chunksize=rows, engine="c", usecols=None,
                  on_bad_lines="skip", nrows=limit, float_precision="round_trip")   # correctly rounded, like float()
        kw["sep"] = sep if sep else r"\s+"
        try:
            reader = pd.read_csv(fh, **kw)
        except pd.errors.EmptyDataError:
            return
        for df in reader:
            a = df.apply(pd.to_numeric, errors="coerce").to_numpy(dtype=np.float64, na_value=np.nan)
            if a.shape[1] < 3:
                continue
            take = list(cols) if cols is not None else [0, 1, 2] + ([3] if a.shape[1] > 3 else [])
            take = [c for c in take if c < a.shape[1]]
            a = a[:, take]
            a = a[~np.isnan(a[:, :3]).any(axis=1)]            # rows that are not three numbers are skipped
            if len(a):
                yield self._block(a[:, :3], a[:, 3] if a.shape[1] > 3 else None)

    def _body_table(self, rows: int) -> Iterator[np.ndarray]:
        h = self.header
        cols = None
        names = [n.lower() for n in h.fields]
        if h.extra["header_row"] and {"x", "y", "z"} <= set(names):
            cols = [names.index(c) for c in "xyz"] + ([names.index("intensity")] if "intensity" in names else [])
        with open(self.path, "r", errors="replace") as fh:
            if h.extra["header_row"]:
                for line in fh:                                  # consume up to and including the header row
                    if line.strip() and not line.lstrip().startswith(("#", "//")):
                        break
            yield from self._ascii_blocks(fh, rows, h.extra["separator"], cols, None)

    def _field_columns(self) -> list[int]:
        names = list(self.header.fields)
        idx = [names.index(c) for c in "xyz"] if {"x", "y", "z"} <= set(names) else [0, 1, 2]
        for cand in ("intensity", "scalar_intensity", "i"):
            if cand in names:
                idx.append(names.index(cand))
                break
        return idx

    def _body_pcd(self, rows: int) -> Iterator[np.ndarray]:
        h = self.header
        meta = h.extra["pcd"]
        if h.encoding == "ascii":
            counts = [int(c) for c in meta.get("COUNT", ["1"] * len(h.fields))]
            starts = np.concatenate([[0], np.cumsum(counts)[:-1]]) if counts else np.array([0, 1, 2])
            cols = [int(starts[i]) for i in self._field_columns()] if "FIELDS" in meta else None
            with open(self.path, "rb") as fb:
                fb.seek(h.data_offset)
                fh = _io.TextIOWrapper(fb, encoding="ascii", errors="replace")
                yield from self._ascii_blocks(fh, rows, None, cols, None)
            return
        sizes = [int(s) for s in meta["SIZE"]]
        types = meta["TYPE"]
        counts = [int(c) for c in meta.get("COUNT", ["1"] * len(sizes))]
        dt = np.dtype([(f"f{k}", "<" + {"F": "f", "I": "i", "U": "u"}[t.upper()] + str(s), (c,) if c != 1 else ())
                       for k, (s, t, c) in enumerate(zip(sizes, types, counts))])
        n = h.n_points if h.n_points is not None else (self.size - h.data_offset) // dt.itemsize
        n = min(n, (self.size - h.data_offset) // dt.itemsize)
        mm = np.memmap(self.path, dtype=dt, mode="r", offset=h.data_offset, shape=(n,))
        cols = self._field_columns()
        for s in range(0, n, rows):
            blk = mm[s:s + rows]
            xyz = np.stack([blk[f"f{c}"].astype(np.float64) for c in cols[:3]], 1)
            yield self._block(xyz, blk[f"f{cols[3]}"].astype(np.float64) if len(cols) > 3 else None)

    def _body_ply(self, rows: int) -> Iterator[np.ndarray]:
        h = self.header
        props = h.extra["properties"]
        cols = self._field_columns()
        if h.encoding == "ascii":
            with open(self.path, "rb") as fb:
                fb.seek(h.data_offset)
                fh = _io.TextIOWrapper(fb, encoding="ascii", errors="replace")
                left = h.n_points
                yield from self._ascii_blocks(fh, rows, None, cols, left)
            return
        order = "<" if h.encoding == "binary" else ">"
        dt = np.dtype([(f"f{k}", order + _PLY_TYPES[t]) for k, (_, t) in enumerate(props)])
        n = min(h.n_points, (self.size - h.data_offset) // dt.itemsize)
        mm = np.memmap(self.path, dtype=dt, mode="r", offset=h.data_offset, shape=(n,))
        for s in range(0, n, rows):
            blk = mm[s:s + rows]
            xyz = np.stack([blk[f"f{c}"].astype(np.float64) for c in cols[:3]], 1)
            yield self._block(xyz, blk[f"f{cols[3]}"].astype(np.float64) if len(cols) > 3 else None)

    def _body_npy(self, rows: int) -> Iterator[np.ndarray]:
        a = np.load(self.path, mmap_mode="r", allow_pickle=False)
        for s in range(0, a.shape[0], rows):
            blk = np.asarray(a[s:s + rows], dtype=np.float64)
            yield self._block(blk[:, :3], blk[:, 3] if blk.shape[1] > 3 else None)

    def _body_las(self, rows: int) -> Iterator[np.ndarray]:
        h = self.header
        rec = h.extra["record_length"]
        dt = np.dtype({"names": ["X", "Y", "Z", "I"], "formats": ["<i4", "<i4", "<i4", "<u2"], "offsets": [0, 4, 8, 12],
                       "itemsize": rec})
        mm = np.memmap(self.path, dtype=dt, mode="r", offset=h.data_offset, shape=(h.n_points,))
        sc, of = np.array(h.extra["scale"]), np.array(h.extra["offset"])
        for s in range(0, h.n_points, rows):
            blk = mm[s:s + rows]
            xyz = np.stack([blk["X"], blk["Y"], blk["Z"]], 1).astype(np.float64) * sc + of
            yield self._block(xyz, blk["I"].astype(np.float64))

    # ---- whole-file reads ------------------------------------------------------------------------------
    def read_xyz(self, rows: int = CHUNK_ROWS) -> np.ndarray:
        """(n, 3) float64 — the array the reference's loaders hand to preprocess_lidar_data."""
        parts = [b[:, :3] for b in self.chunks(rows)]
        return np.concatenate(parts) if parts else np.empty((0, 3))

    def read_frame(self, pinned: bool = False, recenter: bool = False, rows: int = CHUNK_ROWS):
        """The (n, 4) float32 frame (x, y, z, intensity) the device kernels take, filled chunk by chunk.

        pinned    allocate the frame in page-locked memory through the C ABI (`lidar_host_alloc`), so the upload is
                  one DMA; the returned array keeps the allocation alive (`frame.base` is the owning block).
        recenter  subtract a float64 shift (the LAS header offset, else the first chunk's minimum) BEFORE the cast to
                  float32: projected coordinates (UTM: 5e5, 5e6 m) have 3-6 cm of float32 resolution, centimetre
                  voxels need the shift.  Returns (frame, info) with info["shift"] = the (3,) float64 shift applied.
        """
        h = self.header
        shift = None
        if recenter and h.kind == "las":
            shift = np.array(h.extra["offset"], dtype=np.float64)
        cap = h.n_points
        buf, n, owner = None, 0, None

        def alloc(rows_):
            nonlocal owner
            if pinned:
                from . import ops
                owner = ops._PinnedBlock(max(rows_, 1) * 16)
                return owner.u8.view(np.float32).reshape(-1, 4)
            return np.empty((max(rows_, 1), 4), dtype=np.float32)

        for blk in self.chunks(rows):
            if recenter and shift is None:
                shift = blk[:, :3].min(axis=0)
            if shift is not None:
                blk[:, :3] -= shift
            k = blk.shape[0]
            if buf is None:
                buf = alloc(cap if cap is not None else max(k, rows))
            if n + k > buf.shape[0]:                              # ASCII files: the row count is only known at the end
                old, old_owner = buf, owner
                buf = alloc(max(2 * old.shape[0], n + k))
                buf[:n] = old[:n]
                if old_owner is not None:
                    old_owner.free()
            buf[n:n + k] = blk                                    # float64 -> float32, written in place
            n += k
        if buf is None:
            buf = alloc(0)
        frame = buf[:n]
        info = {"format": h.kind, "encoding": h.encoding, "point_count": n, "has_intensity": h.has_intensity,
                "shift": shift if shift is not None else np.zeros(3), "_owner": owner}
        info.update({k: v for k, v in h.extra.items() if k not in ("pcd", "properties")})
        return frame, info


def open_cloud(path: str) -> CloudFile:
    return CloudFile(path)


def load_frame(path: str, pinned: bool = False, recenter: bool = False):
    """(frame (n,4) float32, info) of any supported file; see `CloudFile.read_frame`."""
    return CloudFile(path).read_frame(pinned=pinned, recenter=recenter)
