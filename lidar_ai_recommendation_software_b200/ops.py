"""Device-level operators: torch CUDA tensors in, torch CUDA tensors out, all compute in the
hand-written sm_100a kernels behind the C ABI (include/lidar_b200.h).

These are the building blocks of the drop-in surfaces in `utils/`, `models/` and `apps.py`.
Nothing here falls back to the CPU: without a CUDA device the functions raise.
"""
from __future__ import annotations

import ctypes as C
import sys
import threading
from dataclasses import dataclass

import numpy as np
import torch

from . import _capi, _torch_ext
from ._capi import FMT_F32X4, FMT_F64X3, HIST_AUTO, FrameCaps, FrameDesc, check, lib

__all__ = [
    "require_cuda", "point_format", "bbox", "moments", "hist2d_counts", "hist2d_points_counts", "roi_crop", "ball_count", "set_dbscan_dense", "set_frame_streaming", "set_frame_scan_order", "preprocess_front",
    "FramePipeline", "HostFramePipeline", "voxel_downsample", "voxel_downsample_sorted", "arange_edges", "linspace_edges",
]


_cuda_checked = [False]
_devices: dict[int, torch.device] = {}


def require_cuda() -> torch.device:
    # called by every op: torch.cuda.is_available() / current_stream() walk through NVML, os.environ and several Python
    # layers (together 0.15 ms of a 1.4 ms sequence frame), so the availability is checked once and the raw handles are
    # taken from torch._C directly
    if not _cuda_checked[0]:
        if not torch.cuda.is_available():
            raise RuntimeError("lidar_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        torch.cuda.init()
        _cuda_checked[0] = True
    idx = torch._C._cuda_getDevice()
    dev = _devices.get(idx)
    if dev is None:
        dev = _devices[idx] = torch.device("cuda", idx)
    return dev


def _stream_ptr() -> int:
    """Raw cudaStream_t of torch's current stream on the current device."""
    if not _cuda_checked[0]:
        require_cuda()
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


def _ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else int(t.data_ptr())


class _Scratch(threading.local):
    """Per-thread, per-device scratch buffers (Streamlit runs one script thread per session)."""

    def __init__(self):
        self.bufs: dict[tuple, torch.Tensor] = {}

    def get(self, name: str, nbytes: int, device: torch.device) -> torch.Tensor:
        key = (name, device.index)
        buf = self.bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
            self.bufs[key] = buf
        return buf


_scratch = _Scratch()


class _PinnedScratch(threading.local):
    """Per-thread page-locked staging for small read-backs (descriptors, counters, centroids).  A fresh
    `torch.empty(pin_memory=True)` per call can fall through torch's caching host allocator to cudaHostAlloc,
    which costs about a millisecond and synchronises the device."""

    def __init__(self):
        self.bufs: dict[str, torch.Tensor] = {}

    def get(self, name: str, nbytes: int) -> torch.Tensor:
        buf = self.bufs.get(name)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, pin_memory=True)
            self.bufs[name] = buf
        return buf


_pinned = _PinnedScratch()


def fetch(name: str, *tensors: torch.Tensor):
    """Small device tensors -> numpy through this thread's pinned staging buffer `name`: the copies are enqueued,
    then ONE wait on the current stream.  The arrays are views of the staging buffer: valid until the next
    `fetch` with the same name on this thread (copy them to keep them)."""
    sizes = [t.numel() * t.element_size() for t in tensors]
    offs, o = [], 0
    for b in sizes:
        offs.append(o)
        o += (b + 63) & ~63
    buf = _pinned.get(name, o)
    views = []
    for t, off, b in zip(tensors, offs, sizes):
        v = buf[off:off + b].view(t.dtype).view(t.shape)
        v.copy_(t, non_blocking=True)
        views.append(v)
    check(lib.lidar_stream_synchronize(_stream_ptr()))
    return [v.numpy() for v in views]


def point_format(points: torch.Tensor) -> int:
    """F32X4 for (n,4) float32, F64X3 for (n,3) float64 — contiguous CUDA tensors only."""
    if not points.is_cuda or not points.is_contiguous():
        raise ValueError("points must be a contiguous CUDA tensor")
    if points.dim() == 2 and points.shape[1] == 4 and points.dtype == torch.float32:
        return FMT_F32X4
    if points.dim() == 2 and points.shape[1] == 3 and points.dtype == torch.float64:
        return FMT_F64X3
    raise ValueError(f"unsupported point layout {tuple(points.shape)} {points.dtype}: "
                     "expected (n,4) float32 or (n,3) float64")


# ------------------------------------------------------------------------------------------------
# K1 reductions
# ------------------------------------------------------------------------------------------------
def bbox(points: torch.Tensor) -> torch.Tensor:
    """(8,) float64 device tensor {min x,y,z,w, max x,y,z,w}; np.min/np.max of
    utils/data_processing.py:143,207-208."""
    point_format(points)
    ws = _scratch.get("reduce", lib.lidar_reduce_workspace_bytes(), points.device)
    out = torch.empty(8, dtype=torch.float64, device=points.device)
    _ext_call(_torch_ext.ops.bbox, points, out, ws)      # thin torch extension -> lidar_bbox on PyTorch's current stream
    return out


def moments(points: torch.Tensor, center=(0.0, 0.0, 0.0)) -> torch.Tensor:
    """(6,) float64 {Σ(p-c) per axis, Σ(p-c)² per axis}; np.mean/np.std of data_processing.py:151-152."""
    fmt = point_format(points)
    dev = points.device
    out = torch.empty(6, dtype=torch.float64, device=dev)
    nb = lib.lidar_reduce_workspace_bytes()
    ws = _scratch.get("reduce", nb, dev)
    c3 = (C.c_double * 3)(*[float(v) for v in center])
    check(lib.lidar_moments(_ptr(points), fmt, points.shape[0], c3, _ptr(out), _ptr(ws), ws.numel(), _stream_ptr()))
    return out


def centroid_distances(points: torch.Tensor) -> torch.Tensor:
    """sqrt(sum((p - mean)^2)) per point of an (n,3) float64 CUDA tensor (utils/visualization.py:50-54)."""
    if point_format(points) != FMT_F64X3:
        raise ValueError("expected an (n,3) float64 CUDA tensor")
    dev, n = points.device, points.shape[0]
    out = torch.empty(n, dtype=torch.float64, device=dev)
    ws = _scratch.get("reduce", lib.lidar_reduce_workspace_bytes(), dev)
    check(lib.lidar_centroid_distances(_ptr(points), n, _ptr(out), _ptr(ws), ws.numel(), _stream_ptr()))
    return out


# ------------------------------------------------------------------------------------------------
# edges (host-side parameter derivation — a few scalars, exactly the reference's numpy calls)
# ------------------------------------------------------------------------------------------------
def arange_edges(lo: float, hi: float, grid_size: float) -> np.ndarray:
    """Edges of calculate_grid_density (utils/data_processing.py:305-313): margin 2g, np.arange."""
    margin = grid_size * 2
    lo = lo - margin
    hi = hi + margin
    return np.arange(lo, hi + grid_size, grid_size)


def linspace_edges(lo: float, hi: float, bins: int) -> np.ndarray:
    """Edges np.histogram2d builds for an integer `bins` and explicit range
    (utils/visualization.py:130-134): np.linspace(lo, hi, bins + 1)."""
    return np.linspace(lo, hi, bins + 1)


# ------------------------------------------------------------------------------------------------
# K6 histogram
# ------------------------------------------------------------------------------------------------
def _edges_dev(e: np.ndarray | torch.Tensor, dev) -> torch.Tensor:
    if isinstance(e, torch.Tensor):
        return e.to(device=dev, dtype=torch.float64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(e, dtype=np.float64)).to(dev)


def hist2d_counts(u: torch.Tensor, v: torch.Tensor, x_edges, y_edges, mode: int = HIST_AUTO,
                  out: torch.Tensor | None = None) -> torch.Tensor:
    """int32 (nx, ny) counts with np.histogram2d semantics; `u`, `v` are fp64 CUDA views (any stride)."""
    if u.dtype != torch.float64 or v.dtype != torch.float64 or not u.is_cuda:
        raise ValueError("hist2d_counts expects float64 CUDA tensors")
    if u.dim() != 1 or v.dim() != 1 or u.shape[0] != v.shape[0]:
        raise ValueError("hist2d_counts expects two 1-D tensors of equal length")
    dev = u.device
    ex, ey = _edges_dev(x_edges, dev), _edges_dev(y_edges, dev)
    nx, ny = ex.numel() - 1, ey.numel() - 1
    if nx < 1 or ny < 1:
        raise ValueError("need at least two edges per axis")
    if out is None:
        out = torch.zeros((nx, ny), dtype=torch.int32, device=dev)
    n = u.shape[0]
    su = u.stride(0) if n > 1 else 1
    sv = v.stride(0) if n > 1 else 1
    check(lib.lidar_hist2d_f64(_ptr(u), su, _ptr(v), sv, n, _ptr(ex), nx, _ptr(ey), ny, _ptr(out), mode,
                               _stream_ptr()))
    return out


def hist2d_points_counts(points: torch.Tensor, x_edges, y_edges, mode: int = HIST_AUTO,
                         out: torch.Tensor | None = None) -> torch.Tensor:
    """Same as `hist2d_counts` on the x/y of a point cloud in F32X4 or F64X3 layout."""
    fmt = point_format(points)
    dev = points.device
    ex, ey = _edges_dev(x_edges, dev), _edges_dev(y_edges, dev)
    nx, ny = ex.numel() - 1, ey.numel() - 1
    if out is None:
        out = torch.zeros((nx, ny), dtype=torch.int32, device=dev)
    if nx < 1 or ny < 1:
        raise ValueError("need at least two edges per axis")
    _ext_call(_torch_ext.ops.hist2d_points, points, ex, ey, out, int(mode))
    return out


# ------------------------------------------------------------------------------------------------
# a5 ROI crop
# ------------------------------------------------------------------------------------------------
def roi_crop(points: torch.Tensor, lo, hi, return_mask: bool = True):
    """Axis-aligned box crop, order preserving.

    Returns (cropped points, mask uint8 (n,) or None).  One host sync to learn the kept count.
    """
    fmt = point_format(points)
    dev = points.device
    n = points.shape[0]
    out = torch.empty_like(points)
    mask = torch.empty(n, dtype=torch.uint8, device=dev) if return_mask else None
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    nb = lib.lidar_compact_workspace_bytes(n)
    ws = _scratch.get("compact", nb, dev)
    lo3 = (C.c_double * 3)(*[float(x) for x in lo])
    hi3 = (C.c_double * 3)(*[float(x) for x in hi])
    check(lib.lidar_roi_crop(_ptr(points), fmt, n, lo3, hi3, _ptr(mask), _ptr(out), _ptr(count), _ptr(ws),
                             ws.numel(), _stream_ptr()))
    kept = int(count.item())
    return out[:kept], mask


# ------------------------------------------------------------------------------------------------
# K5 (+K6) frame pipeline
# ------------------------------------------------------------------------------------------------
FRAME_AUTO, FRAME_MULTIKERNEL, FRAME_FUSED, FRAME_PARTITIONED = 0, 1, 2, 3
USE_TORCH_EXTENSION = True            # FramePipeline.enqueue through torch.ops.lidar_b200 (False: the ctypes binding)
_EMPTY_F64 = torch.empty(0, dtype=torch.float64)


def _ext_call(fn, *args):
    """Call an operator of the thin torch extension: it returns the C-ABI status, which becomes the package's
    `LidarError` exactly as on the ctypes path (`lidar_last_error()` is thread-local and the operator ran on this thread)."""
    check(int(fn(*args)))


def set_frame_mode(mode: int = FRAME_AUTO, threads: int = 0, ctas_per_sm: int = 0, smem_kb: int = 0) -> None:
    """How `FramePipeline.enqueue` runs a frame (process-wide): one persistent cooperative kernel with
    the frame resident in shared memory (FRAME_FUSED), five dependent kernels (FRAME_MULTIKERNEL), or
    fused with fallback (FRAME_AUTO, default).  `threads`, `ctas_per_sm`, `smem_kb` tune the fused
    kernel (0 = keep / as needed).  Outputs are identical in every mode."""
    check(lib.lidar_frame_set_fused(int(mode), int(threads), int(ctas_per_sm), int(smem_kb)))


def set_frame_scan_order(on: bool = True) -> None:
    """Scan-order variant of the fused frame kernel (per host thread, off by default): for frames as a sensor delivers
    them (adjacent points adjacent in space) and key spaces much larger than the data -- run-length aggregation of
    the L2 atomics across adjacent lanes, and a scan / clean that walks a summary bitmap of the occupied groups.
    Identical outputs; the default variant is the faster one on shuffled, dense frames."""
    check(lib.lidar_frame_set_fused_scan_order(1 if on else 0))


_streaming_owner: list = [None]


def set_frame_streaming(on: bool = True, inputs_complete: bool = False) -> None:
    """Streaming mode of the fused frame kernel (process-wide, off by default): ordinary launch + programmatic
    dependent launch, so that with frames enqueued back to back on ONE stream the launch gap disappears.
    `inputs_complete=True` additionally lets the next frame's TMA load and bounding box run under the current
    frame's tail — only valid when every frame is COMPLETE in device memory before it is enqueued (resident frames,
    or an event wait after whatever produced it): a frame written by the kernel just before `enqueue` on the same
    stream is not guaranteed visible before `griddepcontrol.wait`.
    Only for a single pipeline per device: two fused kernels of different streams launched this way could each hold
    part of the SMs and wait for each other (the cooperative launch of the default mode rules that out); the first
    `FramePipeline` that enqueues in streaming mode becomes its owner and any other pipeline raises."""
    check(lib.lidar_frame_set_fused_plain_launch(1 if on else 0))
    check(lib.lidar_frame_set_fused_pdl((2 if inputs_complete else 1) if on else 0))
    _streaming_owner[0] = True if on else None


@dataclass
class FrameResult:
    desc: FrameDesc
    voxel_key: torch.Tensor      # (n,) int32     per-point voxel key (B.1 voxel_idx)
    inverse: torch.Tensor        # (n,) int32     rank of the point's voxel
    centroids: torch.Tensor      # (V,4) float32  x,y,z,mean intensity, ascending key
    counts: torch.Tensor         # (V,) int32
    unique_keys: torch.Tensor    # (V,) int32
    grid_counts: torch.Tensor | None   # (nx,ny) int32, calculate_grid_density counts
    dims: tuple
    origin: tuple

    @property
    def n_voxels(self) -> int:
        return int(self.desc.n_voxels)

    def grid_edges(self):
        """(x_edges, y_edges) rebuilt from the device-derived arange parameters (bit-identical to
        np.arange)."""
        d = self.desc

        def edges(a, e1, dl, nb):
            i = np.arange(nb + 1, dtype=np.float64)
            e = a + i * dl
            e[0] = a
            if nb >= 1:
                e[1] = e1
            return e

        return edges(d.ex0, d.ex1, d.exd, d.nx), edges(d.ey0, d.ey1, d.eyd, d.ny)


class FramePipeline:
    """Voxel downsample (+ density grid) of float4 frames with fixed capacities, no host round trip.

    One instance per stream of frames: it owns the workspace (occupancy bitmap, accumulators) and
    the output buffers, so `enqueue()` is ONE kernel launch (the fused persistent kernel; five dependent
    launches on the multi-kernel back end) and nothing else.
    """

    def __init__(self, max_points: int, voxel_size: float, grid_size: float = 0.0,
                 max_key_space: int = 1 << 28, max_nx: int = 1024, max_ny: int = 1024,
                 device: torch.device | None = None, scan_order: bool | str = "auto"):
        """`scan_order`: which variant of the fused kernel runs the frames (identical outputs) -- False: the default
        (shuffled / dense frames), True: the scan-order variant (`set_frame_scan_order`), "auto": decided from the
        last descriptor read back by `result()` (more than 1.5 occupancy groups per point, or fewer than one voxel
        per two points, say the next frame is a sensor-ordered / sparse one too)."""
        self.scan_order = scan_order
        self._scan_hint = False
        self.device = device or require_cuda()
        self.voxel_size = float(voxel_size)
        self.grid_size = float(grid_size)
        self.caps = FrameCaps(int(max_points), int(max_key_space), int(max_nx), int(max_ny))
        nb = lib.lidar_frame_workspace_bytes(C.byref(self.caps))
        if nb == 0:
            raise ValueError("invalid frame capacities")
        dev = self.device
        self.ws = torch.empty(nb, dtype=torch.uint8, device=dev)
        n = int(max_points)
        self.voxel_key = torch.empty(n, dtype=torch.int32, device=dev)
        self.inverse = torch.empty(n, dtype=torch.int32, device=dev)
        # lidar_voxel records (32 B): float x,y,z,intensity; int32 count, key, pad[2]
        self.voxels = torch.empty((n, 8), dtype=torch.float32, device=dev)
        self.centroids = self.voxels[:, :4]
        self.counts = self.voxels.view(torch.int32)[:, 4]
        self.unique_keys = self.voxels.view(torch.int32)[:, 5]
        self.grid = (torch.empty(int(max_nx) * int(max_ny), dtype=torch.int32, device=dev)
                     if grid_size > 0 else None)
        self.desc_dev = torch.zeros(C.sizeof(FrameDesc), dtype=torch.uint8, device=dev)
        self.desc_host = torch.zeros(C.sizeof(FrameDesc), dtype=torch.uint8).pin_memory()
        self._n = 0
        self.reset()

    def reset(self) -> None:
        """(Re)establish the all-zero invariant of the persistent workspace."""
        check(lib.lidar_frame_workspace_init(_ptr(self.ws), self.ws.numel(), C.byref(self.caps), _stream_ptr()))

    def enqueue(self, points: torch.Tensor, origin=None, xy_range=None, events=None) -> None:
        """Enqueue one frame on the current stream (asynchronous).

        `events`: optional list of six `torch.cuda.Event(enable_timing=True)` recorded around the five
        kernels (bench.py attributes device time per kernel with them)."""
        if point_format(points) != FMT_F32X4:
            raise ValueError("FramePipeline takes (n,4) float32 frames")
        n = points.shape[0]
        if _streaming_owner[0] is not None:
            # streaming mode gives up the gang scheduling of the cooperative launch: exactly one pipeline may use it
            if _streaming_owner[0] is True:
                _streaming_owner[0] = id(self)
            elif _streaming_owner[0] != id(self):
                raise RuntimeError("set_frame_streaming(True) allows ONE FramePipeline per process; switch it off "
                                   "(set_frame_streaming(False)) before driving a second pipeline")
        o3 = (C.c_double * 3)(*[float(v) for v in origin]) if origin is not None else None
        r4 = (C.c_double * 4)(*[float(v) for v in xy_range]) if xy_range is not None else None
        self._n = n
        if self.scan_order is not None:
            want = self._scan_hint if self.scan_order == "auto" else bool(self.scan_order)
            check(lib.lidar_frame_set_fused_scan_order(1 if want else 0))
        args = (_ptr(points), n, self.voxel_size, self.grid_size, o3, r4, _ptr(self.voxel_key),
                _ptr(self.inverse), _ptr(self.voxels), _ptr(self.grid), _ptr(self.desc_dev), C.byref(self.caps),
                _ptr(self.ws), self.ws.numel(), _stream_ptr())
        try:
            if events is None and USE_TORCH_EXTENSION:
                # the per-frame hot call goes through the thin torch extension: one dispatcher call, the stream is taken
                # in C++, no ctypes marshalling of fifteen arguments
                _ext_call(
                    _torch_ext.ops.frame_voxel_density, points, self.voxel_size, self.grid_size,
                    _EMPTY_F64 if origin is None else torch.tensor([float(v) for v in origin], dtype=torch.float64),
                    _EMPTY_F64 if xy_range is None else torch.tensor([float(v) for v in xy_range], dtype=torch.float64),
                    self.voxel_key, self.inverse, self.voxels, self.grid, self.desc_dev, self.ws,
                    self.caps.max_points, self.caps.max_key_space, self.caps.max_nx, self.caps.max_ny)
            elif events is None:
                check(lib.lidar_frame_voxel_density(*args))
            else:
                if len(events) != 6:
                    raise ValueError("need six CUDA events")
                for e in events:
                    if not e.cuda_event:      # torch creates the cudaEvent_t lazily, on first record
                        e.record()
                ev = (C.c_void_p * 6)(*[int(e.cuda_event) for e in events])
                check(lib.lidar_frame_voxel_density_timed(*args, ev))
        except Exception:
            self.reset()
            raise

    def result(self) -> FrameResult:
        """Synchronise, read the device-derived descriptor back and slice the outputs."""
        self.desc_host.copy_(self.desc_dev, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        desc = FrameDesc.from_buffer_copy(self.desc_host.numpy().tobytes())
        if desc.status != 0:
            self.reset()
            raise _capi.LidarError(int(desc.status),
                                   f"frame exceeded its capacities: key_space={desc.key_space} "
                                   f"(max {self.caps.max_key_space}), grid {desc.nx}x{desc.ny} "
                                   f"(max {self.caps.max_nx}x{self.caps.max_ny})")
        n, v = self._n, int(desc.n_voxels)
        self._scan_hint = n > 0 and (2 * ((int(desc.key_space) + 223) // 224) > 3 * n or 2 * v < n)
        grid = None
        if self.grid is not None:
            grid = self.grid[: desc.nx * desc.ny].view(desc.nx, desc.ny)
        return FrameResult(desc, self.voxel_key[:n], self.inverse[:n], self.centroids[:v], self.counts[:v],
                           self.unique_keys[:v], grid, tuple(desc.dims[:3]), tuple(desc.origin[:3]))


_numa_bound: dict[int, tuple] = {}


def bind_to_device_numa(device: torch.device | int | None = None) -> tuple[int, int]:
    """Pin this thread (and the copy workers it will create) to the CPUs of the NUMA node the GPU hangs off and
    prefer that node's memory; returns (node, cpus) — node -1 when the platform reports none (nothing changes).
    Call it before page-locked buffers are allocated: one rank per GPU, each staging through ITS node's memory, is
    what lets the host side of the copies scale with the GPUs of an 8-GPU box."""
    idx = device.index if isinstance(device, torch.device) else (torch.cuda.current_device() if device is None else int(device))
    if idx not in _numa_bound:
        node, cpus = C.c_int(-1), C.c_int(0)
        check(lib.lidar_bind_to_device_numa(idx, C.byref(node), C.byref(cpus)))
        _numa_bound[idx] = (int(node.value), int(cpus.value))
        # one rank per GPU shares the host with its siblings: the copy workers of all ranks together must not
        # oversubscribe the CPUs (torchrun exports LOCAL_WORLD_SIZE); half the rank's share, at most 8, at least 1
        import os
        local_world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))
        share = max(1, int(cpus.value) // local_world)
        check(lib.lidar_host_copy_threads(max(1, min(8, share // 2))))
    return _numa_bound[idx]


class _PinnedBlock:
    """Page-locked host memory from the C ABI (`lidar_host_alloc`: cudaHostAlloc on the calling, NUMA-bound thread)."""

    def __init__(self, nbytes: int):
        p = C.c_void_p()
        check(lib.lidar_host_alloc(int(nbytes), C.byref(p)))
        self.ptr, self.nbytes = int(p.value), int(nbytes)
        self.u8 = np.ctypeslib.as_array((C.c_uint8 * self.nbytes).from_address(self.ptr))

    def free(self):
        if self.ptr:
            self.u8 = None
            lib.lidar_host_free(self.ptr)
            self.ptr = 0

    def __del__(self):  # pragma: no cover
        try:
            self.free()
        except Exception:
            pass


class ResultPool:
    """Host memory for results the caller owns — page-locked, so the device writes them directly.

    The drop-in surfaces return arrays the caller may keep for as long as it likes (SURVEY.md §8b "Ownership"), so a
    result can never be a view of a staging buffer the library rewrites.  A FRESH 28 MB numpy allocation per frame
    plus a copy into it costs more than the DMA itself (mmap, a page fault per page, 2 x 28 MB of memory traffic on a
    host whose memory bandwidth the DMA needs as well).  The pool keeps a few page-locked buffers (`lidar_host_alloc`)
    and hands one out only when NO array referencing it is alive — numpy makes every view, and every view of a view,
    hold a reference to the object that owns the memory, so `sys.getrefcount` of the pooled array tells exactly that.
    The read-back lands in the buffer itself: no host copy at all.  A caller that keeps its results simply makes the
    pool allocate new buffers (each frees its page-locked block when the last view dies); nothing a caller holds is
    ever overwritten.  `pinned=False` gives ordinary pageable buffers (no CUDA needed)."""

    def __init__(self, max_buffers: int = 6, pinned: bool = True):
        self.bufs: list[np.ndarray] = []
        self.max_buffers = int(max_buffers)
        self.pinned = bool(pinned)

    def _new(self, nbytes: int) -> np.ndarray:
        nbytes = (nbytes + 4095) & ~4095
        if not self.pinned:
            return np.empty(nbytes, dtype=np.uint8)
        block = _PinnedBlock(nbytes)
        cbuf = (C.c_uint8 * nbytes).from_address(block.ptr)
        cbuf._lidar_owner = block               # the ctypes buffer (numpy's base object) keeps the page-locked block alive
        block.u8 = None
        return np.ctypeslib.as_array(cbuf)

    def reserve(self, count: int, nbytes: int) -> None:
        """Allocate `count` buffers of `nbytes` now: a page-locked allocation costs ~10 ms and stalls every CUDA call of
        the process while it runs -- not something to meet on the first frames of a stream."""
        while len(self.bufs) < min(int(count), self.max_buffers):
            self.bufs.append(self._new(max(int(nbytes), 64)))

    def take(self, nbytes: int) -> np.ndarray:
        nbytes = max(int(nbytes), 64)
        for k in range(len(self.bufs)):
            # references: the list, and the temporary inside getrefcount()
            if self.bufs[k].nbytes >= nbytes and sys.getrefcount(self.bufs[k]) == 2:
                return self.bufs[k]
        buf = self._new(nbytes)
        if len(self.bufs) >= self.max_buffers:
            for k in range(len(self.bufs)):                       # drop a buffer nobody uses (too small), else keep none
                if sys.getrefcount(self.bufs[k]) == 2:
                    del self.bufs[k]
                    break
        if len(self.bufs) < self.max_buffers:
            self.bufs.append(buf)
        return buf


def _is_pinned(a: np.ndarray) -> bool:
    t = torch.from_numpy(a)
    try:
        return bool(t.is_pinned())
    except Exception:
        return False


class HostFramePipeline:
    """The call a user of the numpy surface makes: HOST float4 frame in, HOST numpy results out.

    The drop-in contract (SURVEY.md §8b "Ownership"): the frame is a caller-owned numpy array in ordinary pageable
    memory, the results are fresh arrays the caller owns.  `submit` stages the frame into the slot's page-locked
    buffer with the parallel host memcpy of the C ABI (`lidar_host_memcpy`) and enqueues copy-in, the frame kernel and
    the SoA repack on the slot's stream (`lidar_frame_voxel_density_host_begin`); `collect` reads the descriptor,
    fetches exactly what the frame produced (`lidar_frame_host_fetch`: 8 bytes per point + 20 bytes per VOXEL + the
    grid) and copies it out of the staging block into owned arrays, again on the worker pool.  Slots alternate on
    their own streams, so the copies of one frame overlap the kernels and copies of the next.

    Streaming mode for a sensor driver that owns page-locked buffers: pass a pinned frame (no staging copy) and
    `collect(copy=False)` (views of the slot's block, valid until the slot is reused); `per_point_outputs=False`
    (LIDAR_HOST_NO_PER_POINT) additionally leaves voxel_key / inverse on the device — 8 of the 28 bytes per point
    that cross PCIe on the way back.
    """

    def __init__(self, max_points: int, voxel_size: float, grid_size: float = 0.0, slots: int = 2,
                 per_point_outputs: bool = True, unique_keys: bool = False, scan_order: bool | str = "auto",
                 two_stage: bool = True, numa_bind: bool = True, **caps):
        self.device = require_cuda()
        if numa_bind:
            self.numa = bind_to_device_numa(self.device)
        self.scan_order = scan_order          # as in FramePipeline: False / True / "auto" (from the last descriptor)
        self._scan_hint = False
        self.per_point_outputs = per_point_outputs
        self.unique_keys = unique_keys
        self.two_stage = bool(two_stage)
        self.voxel_size, self.grid_size = float(voxel_size), float(grid_size)
        self.flags = (_capi.HOST_UNIQUE_KEYS if unique_keys else 0) | (0 if per_point_outputs else _capi.HOST_NO_PER_POINT)
        self.slots = []
        n = int(max_points)
        dev = self.device
        c = dict(max_key_space=1 << 28, max_nx=1024, max_ny=1024)
        c.update(caps)
        if self.grid_size <= 0:
            c["max_nx"] = c["max_ny"] = 0
        self.caps = FrameCaps(n, int(c["max_key_space"]), int(c["max_nx"]), int(c["max_ny"]))
        ws_bytes = lib.lidar_frame_workspace_bytes(C.byref(self.caps))
        block = lib.lidar_frame_host_block_bytes(n, C.byref(self.caps), self.flags)
        if ws_bytes == 0 or block == 0:
            raise ValueError("invalid frame capacities")
        for _ in range(slots):
            st = torch.cuda.Stream(device=dev)
            slot = {
                "stream": st, "n": 0, "event": torch.cuda.Event(),
                "ws": torch.empty(ws_bytes, dtype=torch.uint8, device=dev),
                "h_in": _PinnedBlock(max(n, 1) * 16),
                "d_in": torch.empty((n, 4), dtype=torch.float32, device=dev),
                "d_vox": torch.empty((n, 8), dtype=torch.float32, device=dev),
                "d_out": torch.empty(block, dtype=torch.uint8, device=dev),
                "h_out": _PinnedBlock(block),
            }
            check(lib.lidar_frame_workspace_init(_ptr(slot["ws"]), ws_bytes, C.byref(self.caps), _stream_ptr()))
            self.slots.append(slot)
        self._next = 0
        self._pending: list[dict] = []
        self._last_d2h = 0
        self._pool = ResultPool()
        if self.two_stage:
            # the results in flight plus the one the caller is still looking at while it collects the next
            self._pool.reserve(slots + 1, block)
        torch.cuda.synchronize(self.device)   # workspace zeroing ran on the constructing stream

    def close(self):
        for s in self.slots:
            s["stream"].synchronize()
            s["h_in"].free()
            s["h_out"].free()
        self.slots = []

    def _layout(self, n: int):
        off = (C.c_size_t * 7)()
        check(lib.lidar_frame_host_block_layout(n, C.byref(self.caps), self.flags, off))
        return [int(v) for v in off]

    def h2d_bytes(self, n: int) -> int:
        return n * 16

    def d2h_bytes(self, n: int, n_voxels: int | None = None, nx: int = 0, ny: int = 0) -> int:
        """Bytes copied device -> host per frame of n points.  One-copy form: the block at the size of the frame
        (the voxel count is not known on the host when the copy is enqueued); two-stage form: the descriptor, then
        exactly 8n (per-point outputs) + 20 V (+ 4 V keys) + 4 nx ny."""
        if not self.two_stage or n_voxels is None:
            total = int(lib.lidar_frame_host_block_bytes(n, C.byref(self.caps), self.flags))
            return total - (0 if self.per_point_outputs else self._layout(n)[2])
        per_voxel = 20 + (4 if self.unique_keys else 0)
        return (C.sizeof(FrameDesc) + (8 * n if self.per_point_outputs else 0) + per_voxel * int(n_voxels)
                + 4 * int(nx) * int(ny))

    def submit(self, points: np.ndarray | torch.Tensor, origin=None, xy_range=None) -> None:
        """Stage one host frame and enqueue copy-in, kernels and the first read-back on the slot's stream."""
        if _streaming_owner[0] is not None:
            raise RuntimeError("HostFramePipeline runs several frames on several streams: switch the streaming mode "
                               "of the frame kernel off first (set_frame_streaming(False))")
        slot = self.slots[self._next]
        self._next = (self._next + 1) % len(self.slots)
        if slot in self._pending:
            raise RuntimeError("all slots busy: call collect() first")
        src = points.numpy() if isinstance(points, torch.Tensor) else points
        if src.ndim != 2 or src.shape[1] != 4 or src.dtype != np.float32 or not src.flags.c_contiguous:
            raise ValueError("HostFramePipeline takes contiguous (n,4) float32 frames")
        n = src.shape[0]
        if n > self.caps.max_points:
            raise _capi.LidarError(-4, f"frame of {n} points exceeds max_points={self.caps.max_points}")
        pinned = isinstance(points, torch.Tensor) and points.is_pinned()
        if pinned:
            h_in_ptr = points.data_ptr()    # caller already owns page-locked memory: no staging copy
            slot["src"] = points            # keep the source alive until the copy-in has run
        else:
            h_in_ptr = slot["h_in"].ptr
            check(lib.lidar_host_memcpy(h_in_ptr, src.ctypes.data, n * 16))
            slot["src"] = None
        slot["n"] = n
        o3 = (C.c_double * 3)(*[float(v) for v in origin]) if origin is not None else None
        r4 = (C.c_double * 4)(*[float(v) for v in xy_range]) if xy_range is not None else None
        if self.scan_order is not None:
            want = self._scan_hint if self.scan_order == "auto" else bool(self.scan_order)
            check(lib.lidar_frame_set_fused_scan_order(1 if want else 0))
        entry = lib.lidar_frame_voxel_density_host_begin if self.two_stage else lib.lidar_frame_voxel_density_host
        try:
            check(entry(h_in_ptr, n, self.voxel_size, self.grid_size, o3, r4, _ptr(slot["d_in"]), _ptr(slot["d_vox"]),
                        _ptr(slot["d_out"]), slot["h_out"].ptr, self.flags, C.byref(self.caps), _ptr(slot["ws"]),
                        slot["ws"].numel(), slot["stream"].cuda_stream))
        except Exception:
            with torch.cuda.stream(slot["stream"]):
                check(lib.lidar_frame_workspace_init(_ptr(slot["ws"]), slot["ws"].numel(), C.byref(self.caps), _stream_ptr()))
            raise
        self._pending.append(slot)

    def collect(self, copy: bool = True) -> dict:
        """Wait for the oldest submitted frame and return its results as numpy arrays the caller owns (`copy=False`:
        views of the slot's page-locked block, valid until the slot is reused)."""
        slot = self._pending.pop(0)
        slot["stream"].synchronize()
        n = slot["n"]
        off = self._layout(n)
        raw = slot["h_out"].u8
        desc = FrameDesc.from_buffer_copy(raw[off[6]: off[6] + C.sizeof(FrameDesc)].tobytes())
        if desc.status != 0:
            with torch.cuda.stream(slot["stream"]):
                check(lib.lidar_frame_workspace_init(_ptr(slot["ws"]), slot["ws"].numel(), C.byref(self.caps), _stream_ptr()))
            slot["stream"].synchronize()
            raise _capi.LidarError(int(desc.status), "frame exceeded its capacities")
        v = int(desc.n_voxels)
        nx, ny = (int(desc.nx), int(desc.ny)) if self.grid_size > 0 else (0, 0)
        # owned results (`copy=True`): a recycled page-locked buffer laid out like the staging block (`ResultPool`).
        # Two-stage form: the device writes the arrays straight into it (no host copy); one-copy form: the frame-sized
        # block already sits in the slot's staging memory and is copied out by the worker pool.
        dst = self._pool.take(off[6]) if copy else None
        direct = copy and self.two_stage
        if self.two_stage:
            target = dst.ctypes.data if direct else slot["h_out"].ptr
            check(lib.lidar_frame_host_fetch(n, v, nx, ny, _ptr(slot["d_out"]), target, self.flags,
                                             C.byref(self.caps), slot["stream"].cuda_stream))
            slot["stream"].synchronize()
        elif copy:
            lib.lidar_host_copy_wake()
        self._last_d2h = self.d2h_bytes(n, v, nx, ny)
        self._scan_hint = n > 0 and (2 * ((int(desc.key_space) + 223) // 224) > 3 * n or 2 * v < n)
        jobs: list[tuple[int, int]] = []                 # (offset, bytes) of every array copied out of the staging block

        def arr(k, dtype, count, shape=None):
            if copy:
                view = np.frombuffer(dst, dtype=dtype, count=count, offset=off[k])
                if count and not direct:
                    jobs.append((off[k], view.nbytes))
            else:
                view = np.frombuffer(raw, dtype=dtype, count=count, offset=off[k])
            return view if shape is None else view.reshape(shape)

        out = {
            "centroids": arr(2, np.float32, 4 * v, (v, 4)), "counts": arr(3, np.int32, v),
            "n_voxels": v, "dims": tuple(desc.dims[:3]), "origin": tuple(desc.origin[:3]), "desc": desc,
        }
        if self.unique_keys:
            out["unique_keys"] = arr(4, np.int32, v)
        if self.per_point_outputs:
            out["inverse"] = arr(1, np.int32, n)
            out["voxel_key"] = arr(0, np.int32, n)
        if self.grid_size > 0:
            out["grid_counts"] = arr(5, np.int32, nx * ny, (nx, ny))
        if jobs:
            # ONE job for the copy workers: all arrays of the frame, staging block -> the caller's buffer
            k = len(jobs)
            base_d, base_s = dst.ctypes.data, slot["h_out"].ptr
            check(lib.lidar_host_memcpy_batch(k, (C.c_void_p * k)(*[base_d + o for o, _ in jobs]),
                                              (C.c_void_p * k)(*[base_s + o for o, _ in jobs]),
                                              (C.c_size_t * k)(*[b for _, b in jobs])))
        return out

    def process(self, points, origin=None, xy_range=None) -> dict:
        self.submit(points, origin=origin, xy_range=xy_range)
        return self.collect()


@dataclass
class SortedVoxelResult:
    """Result of the 64-bit-key sort path (`voxel_downsample_sorted`); per-point entries of cropped points are -1."""
    desc: "_capi.SortedDesc"
    voxel_key: torch.Tensor      # (n,) int64
    inverse: torch.Tensor        # (n,) int32
    centroids: torch.Tensor      # (V,4) float32
    counts: torch.Tensor         # (V,) int32
    unique_keys: torch.Tensor    # (V,) int64
    dims: tuple
    origin: tuple

    @property
    def n_voxels(self) -> int:
        return int(self.desc.n_voxels)

    @property
    def n_kept(self) -> int:
        return int(self.desc.n_kept)


def voxel_downsample_sorted(points: torch.Tensor, voxel_size: float, origin=None, roi=None) -> SortedVoxelResult:
    """Voxel downsample with int64 keys, any key space — deterministic radix sort by voxel key +
    segmented reduction (`lidar_voxel_downsample_sorted`), one enqueue and one read-back of the descriptor.
    `roi=(lo3, hi3)` fuses the ROI crop of Appendix B.2 in front: cropped points take no part in the origin / bbox and get
    voxel_key = inverse = -1."""
    if point_format(points) != FMT_F32X4:
        raise ValueError("voxel_downsample_sorted takes (n,4) float32 frames")
    dev, n = points.device, points.shape[0]
    cap = max(n, 1)
    key = torch.empty(cap, dtype=torch.int64, device=dev)
    inv = torch.empty(cap, dtype=torch.int32, device=dev)
    cent = torch.empty((cap, 4), dtype=torch.float32, device=dev)
    cnt = torch.empty(cap, dtype=torch.int32, device=dev)
    ukey = torch.empty(cap, dtype=torch.int64, device=dev)
    d_desc = torch.zeros(C.sizeof(_capi.SortedDesc), dtype=torch.uint8, device=dev)
    ws = _scratch.get("voxel_sorted", lib.lidar_voxel_sorted_workspace_bytes(n), dev)
    o3 = (C.c_double * 3)(*[float(v) for v in origin]) if origin is not None else None
    lo3 = (C.c_double * 3)(*[float(v) for v in roi[0]]) if roi is not None else None
    hi3 = (C.c_double * 3)(*[float(v) for v in roi[1]]) if roi is not None else None
    check(lib.lidar_voxel_downsample_sorted(_ptr(points), n, float(voxel_size), o3, lo3, hi3, _ptr(key), _ptr(inv), _ptr(cent),
                                            _ptr(cnt), _ptr(ukey), _ptr(d_desc), _ptr(ws), ws.numel(), _stream_ptr()))
    desc = _capi.SortedDesc.from_buffer_copy(fetch("sorted_desc", d_desc)[0].tobytes())
    if desc.status != 0:
        raise _capi.LidarError(int(desc.status), "voxel_downsample_sorted: a point lies below the given origin" if desc.status == -1
                               else "voxel_downsample_sorted: the voxel key space does not fit 63 bits")
    v = int(desc.n_voxels)
    return SortedVoxelResult(desc, key[:n], inv[:n], cent[:v], cnt[:v], ukey[:v], tuple(int(d) for d in desc.dims),
                             tuple(float(o) for o in desc.origin))


def voxel_downsample(points: torch.Tensor, voxel_size: float, origin=None, max_key_space: int | None = None):
    """One-shot voxel downsample of an (n,4) float32 CUDA tensor.

    Returns a `FrameResult` (int32 keys, the occupancy-bitmap frame kernel) whose tensors are owned by the caller, or —
    when the key space of the cloud needs more than 31 bits (one far outlier is enough) — a `SortedVoxelResult` with
    int64 keys from the sort path: the same fields, the same values.
    """
    n = points.shape[0]
    if max_key_space is None:
        bb = bbox(points).cpu().numpy()
        org = bb[:3] if origin is None else np.asarray(origin, dtype=np.float64)
        dims = np.floor((bb[4:7] - org) / float(voxel_size)) + 1
        max_key_space = int(max(1, np.prod(np.maximum(dims, 1)))) if np.all(np.isfinite(dims)) else 1
        if max_key_space >= (1 << 31):
            return voxel_downsample_sorted(points, voxel_size, origin=origin)
    pipe = FramePipeline(max(n, 1), voxel_size, 0.0, max_key_space=max_key_space, device=points.device)
    pipe.enqueue(points, origin=origin)
    return pipe.result()


# ------------------------------------------------------------------------------------------------
# K2-K4 preprocess stages, K7 DBSCAN, K8 centroids  (all (n,3) float64 CUDA tensors)
# ------------------------------------------------------------------------------------------------
def _d3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


def _check_f64x3(points: torch.Tensor):
    if point_format(points) != FMT_F64X3:
        raise ValueError("expected an (n,3) float64 CUDA tensor")


def sigma_filter(points: torch.Tensor, mean, thr, tol, zmin: float, zden: float, want_colors: bool = True,
                 want_mask: bool = False):
    """3-sigma inlier filter (utils/data_processing.py:151-157) fused with the height colours (:143-147).

    Returns (inliers (n',3), colors (n',3) | None, mask uint8 | None, guard:int)."""
    _check_f64x3(points)
    dev, n = points.device, points.shape[0]
    out = torch.empty_like(points)
    col = torch.empty_like(points) if want_colors else None
    mask = torch.empty(n, dtype=torch.uint8, device=dev) if want_mask else None
    cnt = torch.zeros(2, dtype=torch.int64, device=dev)   # [count, guard]
    ws = _scratch.get("pre", lib.lidar_preprocess_workspace_bytes(n), dev)
    check(lib.lidar_sigma_filter(_ptr(points), n, _d3(mean), _d3(thr), _d3(tol), float(zmin), float(zden),
                                 _ptr(mask), _ptr(out), _ptr(col), _ptr(cnt), _ptr(cnt[1:]), _ptr(ws), ws.numel(),
                                 _stream_ptr()))
    kept, guard = (int(v) for v in cnt.tolist())
    return out[:kept], (col[:kept] if col is not None else None), mask, guard


def select_kth(column: torch.Tensor, k: int):
    """(x_(k), x_(k+1)) of a 1-D float64 CUDA view (any stride) — exact radix select."""
    if column.dtype != torch.float64 or column.dim() != 1 or not column.is_cuda:
        raise ValueError("select_kth expects a 1-D float64 CUDA tensor")
    n = column.shape[0]
    dev = column.device
    out = torch.empty(2, dtype=torch.float64, device=dev)
    ws = _scratch.get("pre", lib.lidar_preprocess_workspace_bytes(n), dev)
    stride = column.stride(0) if n > 1 else 1
    check(lib.lidar_select_kth(_ptr(column), stride, n, int(k), _ptr(out), _ptr(ws), ws.numel(), _stream_ptr()))
    a, b = out.tolist()
    return a, b


def preprocess_front(points: torch.Tensor, want_colors: bool = True, scaler: bool = False):
    """Everything of the preprocess functions before DBSCAN as ONE enqueue and ONE read-back
    (`lidar_preprocess_front`): bbox / mean / std of the raw cloud, 3-sigma filter (+ colours), the 30th
    percentile of the inlier heights, ground split with the plane sums and both bboxes, and for variant A the
    StandardScaler statistics, the scaled copy and eps (utils/data_processing.py:143-196).

    Returns (desc: _capi.FrontDesc (host copy), inliers, colors | None, non_ground, non_ground_index int32,
    scaled | None); the arrays are views of capacity-n buffers cut to the counts in `desc`."""
    _check_f64x3(points)
    dev, n = points.device, points.shape[0]
    inl = torch.empty_like(points)
    col = torch.empty_like(points) if want_colors else None
    ng = torch.empty_like(points)
    idx = torch.empty(n, dtype=torch.int32, device=dev)
    X = torch.empty_like(points) if scaler else None
    nb = C.sizeof(_capi.FrontDesc)
    d_desc = _scratch.get("front_desc", nb, dev)
    ws = _scratch.get("front", lib.lidar_preprocess_front_workspace_bytes(n), dev)
    flags = (_capi.FRONT_COLORS if want_colors else 0) | (_capi.FRONT_SCALER if scaler else 0)
    check(lib.lidar_preprocess_front(_ptr(points), n, flags, _ptr(inl), _ptr(col), _ptr(ng), _ptr(idx), _ptr(X),
                                     _ptr(d_desc), _ptr(ws), ws.numel(), _stream_ptr()))
    desc = _capi.FrontDesc.from_buffer_copy(fetch("front_desc", d_desc[:nb])[0].tobytes())
    n_in, m = int(desc.n_in), int(desc.n_nonground)
    return (desc, inl[:n_in], (col[:n_in] if col is not None else None), ng[:m], idx[:m],
            (X[:m] if X is not None else None))


def ground_split(points: torch.Tensor, z_threshold: float, center, tol: float = 0.0):
    """z <= thr split (utils/data_processing.py:165-188).

    Returns (non_ground (m,3), non_ground_index int32 (m,), plane_sums (10,) numpy, guard:int)."""
    _check_f64x3(points)
    dev, n = points.device, points.shape[0]
    out = torch.empty_like(points)
    idx = torch.empty(n, dtype=torch.int32, device=dev)
    cnt = torch.zeros(2, dtype=torch.int64, device=dev)
    plane = torch.zeros(10, dtype=torch.float64, device=dev)
    ws = _scratch.get("pre", lib.lidar_preprocess_workspace_bytes(n), dev)
    check(lib.lidar_ground_split(_ptr(points), n, float(z_threshold), _d3(center), float(tol), _ptr(out), _ptr(idx),
                                 _ptr(cnt), _ptr(plane), _ptr(cnt[1:]), _ptr(ws), ws.numel(), _stream_ptr()))
    m, guard = (int(v) for v in cnt.tolist())
    return out[:m], idx[:m], plane.cpu().numpy(), guard


def standardize(points: torch.Tensor, mean, scale) -> torch.Tensor:
    """(x - mean) / scale — StandardScaler.transform (utils/data_processing.py:190-191)."""
    _check_f64x3(points)
    out = torch.empty_like(points)
    check(lib.lidar_standardize(_ptr(points), points.shape[0], _d3(mean), _d3(scale), _ptr(out), _stream_ptr()))
    return out


def scatter_labels(labels: torch.Tensor, index: torch.Tensor, n: int) -> torch.Tensor:
    """int64 (n,) = -1 everywhere, labels at `index` (utils/data_processing.py:203-204)."""
    dev = labels.device
    full = torch.empty(n, dtype=torch.int64, device=dev)
    check(lib.lidar_scatter_labels(_ptr(labels), _ptr(index), labels.shape[0], _ptr(full), n, _stream_ptr()))
    return full


def set_dbscan_dense(on: bool = True) -> None:
    """Process-wide: allow (default) or forbid the dense cell grid of lidar_dbscan (same labels either way)."""
    check(lib.lidar_dbscan_set_dense(1 if on else 0))


def dbscan(points: torch.Tensor, eps: float, min_samples: int = 5, tol: float = 0.0, bounds=None, defer: bool = False):
    """sklearn-identical DBSCAN labels (int32 (m,)), number of clusters, knife-edge guard count.
    `defer=True` skips the read-back: returns (labels, info) with info a 2-element int64 device tensor
    [n_clusters in the low 32 bits, guard] for the caller to fetch together with its other results."""
    _check_f64x3(points)
    dev, m = points.device, points.shape[0]
    labels = torch.empty(m, dtype=torch.int32, device=dev)
    if m == 0:
        return (labels, torch.zeros(2, dtype=torch.int64, device=dev)) if defer else (labels, 0, 0)
    if bounds is None:
        bb = bbox(points).cpu().numpy()
        bounds = (bb[:3], bb[4:7])
    lo, hi = _d3(bounds[0]), _d3(bounds[1])
    nb = lib.lidar_dbscan_workspace_bytes(m, float(eps), lo, hi)
    if nb == 0:
        raise _capi.LidarError(-1, "lidar_dbscan: cannot build a cell grid for this bbox / eps")
    ws = _scratch.get("dbscan", nb, dev)
    info = torch.zeros(2, dtype=torch.int64, device=dev)   # [n_clusters (int32 in low word), guard]
    check(lib.lidar_dbscan(_ptr(points), m, float(eps), int(min_samples), float(tol), lo, hi, _ptr(labels),
                           _ptr(info), _ptr(info[1:]), _ptr(ws), ws.numel(), _stream_ptr()))
    if defer:
        return labels, info
    nc, guard = (int(v) for v in info.tolist())
    return labels, nc & 0xffffffff, guard


def ball_count(points: torch.Tensor, radius: float, bounds=None) -> torch.Tensor:
    """KDTree(points).query_radius(points, r=radius, count_only=True) on the device: int64 (n,) counts, the
    point itself included, inclusive fp64 `rdist <= r*r` (utils/visualization.py:43-45, 167-168).
    `points` is (n,3) float64 CUDA; for a 2-D projection pass the two coordinates and a zero third column."""
    _check_f64x3(points)
    dev, n = points.device, points.shape[0]
    counts = torch.empty(n, dtype=torch.int64, device=dev)
    if n == 0:
        return counts
    if bounds is None:
        bb = bbox(points).cpu().numpy()
        bounds = (bb[:3], bb[4:7])
    lo, hi = _d3(bounds[0]), _d3(bounds[1])
    nb = lib.lidar_ball_count_workspace_bytes(n, float(radius), lo, hi)
    if nb == 0:
        raise _capi.LidarError(-1, "lidar_ball_count: cannot build a cell grid for this bbox / radius")
    ws = _scratch.get("dbscan", nb, dev)
    check(lib.lidar_ball_count(_ptr(points), n, float(radius), lo, hi, _ptr(counts), _ptr(ws), ws.numel(), _stream_ptr()))
    return counts


def cluster_centroids(points: torch.Tensor, labels: torch.Tensor, n_clusters: int):
    """Per-cluster mean of the member points, exact integer accumulation (extract_people_positions,
    utils/data_processing.py:251-280).  Returns ((C,3) float64, (C,) int64 counts)."""
    _check_f64x3(points)
    if labels.dtype not in (torch.int32, torch.int64) or not labels.is_contiguous():
        raise ValueError("labels must be contiguous int32 or int64")
    dev = points.device
    cent = torch.zeros((n_clusters, 3), dtype=torch.float64, device=dev)
    counts = torch.zeros(n_clusters, dtype=torch.int64, device=dev)
    if n_clusters == 0:
        return cent, counts
    ws = _scratch.get("centroid", lib.lidar_centroid_workspace_bytes(n_clusters), dev)
    check(lib.lidar_cluster_centroids(_ptr(points), _ptr(labels), int(labels.dtype == torch.int64), points.shape[0],
                                      n_clusters, _ptr(cent), _ptr(counts), _ptr(ws), ws.numel(), _stream_ptr()))
    return cent, counts


def gather_rows(points, indices):
    """points[indices] on the device (downsample_point_cloud, utils/data_processing.py:247-249).
    numpy in -> numpy out (same dtype); CUDA tensor in -> CUDA tensor out."""
    dev = require_cuda()
    is_np = not isinstance(points, torch.Tensor)
    src = torch.from_numpy(np.ascontiguousarray(points)).to(dev) if is_np else points.contiguous()
    idx = torch.as_tensor(np.asarray(indices, dtype=np.int64)).to(dev) if not isinstance(indices, torch.Tensor) \
        else indices.to(device=dev, dtype=torch.int64).contiguous()
    n = src.shape[0]
    if idx.numel() and (int(idx.min()) < -n or int(idx.max()) >= n):
        raise IndexError("index out of bounds")
    idx = torch.where(idx < 0, idx + n, idx) if idx.numel() else idx
    row_bytes = src[0].numel() * src.element_size() if n else src.element_size()
    if row_bytes % 4:
        raise ValueError("row size must be a multiple of 4 bytes")
    out = torch.empty((idx.numel(),) + tuple(src.shape[1:]), dtype=src.dtype, device=dev)
    check(lib.lidar_gather_rows(_ptr(src), n, row_bytes, _ptr(idx), idx.numel(), _ptr(out), _stream_ptr()))
    return out.cpu().numpy() if is_np else out


# ------------------------------------------------------------------------------------------------
# K9 / K10 flow
# ------------------------------------------------------------------------------------------------
def _f64_dev(a, dev) -> torch.Tensor:
    if isinstance(a, torch.Tensor):
        return a.to(device=dev, dtype=torch.float64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)


def flow_field(x_grid, y_grid, exit_xy, freq: float, amp: float, discs, speed_span: float, clip=None):
    """Lattice flow field (models/crowd_flow_model.py:88-184).  Returns device tensors
    (positions (G,2), vectors (G,2), magnitudes (G,), sums (3,) = [sum|v|, sum vx, sum vy])."""
    dev = require_cuda()
    xg, yg = _f64_dev(x_grid, dev), _f64_dev(y_grid, dev)
    nx, ny = xg.numel(), yg.numel()
    g = nx * ny
    pos = torch.empty((g, 2), dtype=torch.float64, device=dev)
    vec = torch.empty((g, 2), dtype=torch.float64, device=dev)
    mag = torch.empty(g, dtype=torch.float64, device=dev)
    sums = torch.zeros(4, dtype=torch.float64, device=dev)
    flat = [float(v) for d in discs for v in d]
    darr = (C.c_double * max(len(flat), 1))(*flat) if flat else None
    lo, hi = (clip if clip is not None else (0.0, 0.0))
    check(lib.lidar_flow_field(_ptr(xg), nx, _ptr(yg), ny, float(exit_xy[0]), float(exit_xy[1]), float(freq),
                               float(amp), darr, len(discs), float(speed_span), int(clip is not None), float(lo),
                               float(hi), _ptr(pos), _ptr(vec), _ptr(mag), _ptr(sums), _stream_ptr()))
    return pos, vec, mag, sums[:3], (nx, ny)


def flow_bottleneck_severity(nxy, pos, vec, mag) -> torch.Tensor:
    sev = torch.empty_like(mag)
    check(lib.lidar_flow_bottlenecks(nxy[0], nxy[1], _ptr(pos), _ptr(vec), _ptr(mag), _ptr(sev), _stream_ptr()))
    return sev


def flow_box_max(nxy, pos, mag, slow_below: float) -> torch.Tensor:
    out = torch.empty_like(mag)
    check(lib.lidar_flow_box_max(nxy[0], nxy[1], _ptr(pos), _ptr(mag), float(slow_below), _ptr(out), _stream_ptr()))
    return out


def radius_count(centres, qx, qy, radius: float) -> torch.Tensor:
    """int32 (len(qy), len(qx)): #centres within `radius` (inclusive) of every (qx[i], qy[j])."""
    dev = require_cuda()
    c = _f64_dev(centres, dev).reshape(-1, 2)
    x, y = _f64_dev(qx, dev), _f64_dev(qy, dev)
    out = torch.empty((y.numel(), x.numel()), dtype=torch.int32, device=dev)
    check(lib.lidar_radius_count(_ptr(c), c.shape[0], _ptr(x), x.numel(), _ptr(y), y.numel(), float(radius), _ptr(out),
                                 _stream_ptr()))
    return out


def frame_flow_match(prev_xy, cur_xy, dt: float, gate: float = 1.5):
    """B.3 association: (match int32 (C2,), velocity float32 (C2,2)) on the device."""
    dev = require_cuda()

    def f32(a):
        if isinstance(a, torch.Tensor):
            return a.to(device=dev, dtype=torch.float32).contiguous().reshape(-1, 2)
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev).reshape(-1, 2)

    p, c = f32(prev_xy), f32(cur_xy)
    match = torch.empty(c.shape[0], dtype=torch.int32, device=dev)
    vel = torch.zeros((c.shape[0], 2), dtype=torch.float32, device=dev)
    check(lib.lidar_frame_flow_match(_ptr(p), p.shape[0], _ptr(c), c.shape[0], float(dt), float(gate), _ptr(match),
                                     _ptr(vel), _stream_ptr()))
    return match, vel, c


def frame_flow_field(lattice, cur_f32, match, vel, radius: float = 3.0):
    dev = require_cuda()
    lat = _f64_dev(lattice, dev).reshape(-1, 2)
    g = lat.shape[0]
    vec = torch.zeros((g, 2), dtype=torch.float64, device=dev)
    mag = torch.zeros(g, dtype=torch.float64, device=dev)
    check(lib.lidar_frame_flow_field(_ptr(lat), g, _ptr(cur_f32), _ptr(match), _ptr(vel), cur_f32.shape[0],
                                     float(radius), _ptr(vec), _ptr(mag), _stream_ptr()))
    return vec, mag


def frame_flow_step(prev_xy, cur_xy, x_grid, y_grid, dt: float, gate: float = 1.5, radius: float = 3.0, extra=()):
    """The flow step of one sequence frame as ONE staging copy in, ONE C call (`lidar_frame_flow`: lattice, match, field)
    and ONE copy out: (lattice (G,2) f64, vectors (G,2) f64, magnitudes (G,) f64, match (C,) i32, velocity (C,2) f32) as
    numpy arrays cut from one buffer the caller owns, followed by host copies of the device tensors in `extra`.
    `cur_xy`: (C,2) numpy or CUDA tensor (any float dtype; the match runs in float32 as B.3 says).  A dozen torch calls
    per frame (uploads, zeros, views, read-backs) were a third of the interpreter time of a sequence frame."""
    dev = require_cuda()
    xg = np.ascontiguousarray(x_grid, dtype=np.float64)
    yg = np.ascontiguousarray(y_grid, dtype=np.float64)
    prev = np.ascontiguousarray(prev_xy, dtype=np.float32).reshape(-1, 2)
    cur_dev = cur_xy.to(device=dev, dtype=torch.float32).contiguous().reshape(-1, 2) if isinstance(cur_xy, torch.Tensor) else None
    cur = None if cur_dev is not None else np.ascontiguousarray(cur_xy, dtype=np.float32).reshape(-1, 2)
    nx, ny, n_prev = xg.size, yg.size, prev.shape[0]
    n_cur = cur_dev.shape[0] if cur_dev is not None else cur.shape[0]
    g = nx * ny
    # ---- in: [x_grid | y_grid | prev | cur] through this thread's pinned staging, one async copy -------------------
    parts = [xg, yg, prev] + ([cur] if cur is not None else [])
    offs, o = [], 0
    for a in parts:
        offs.append(o)
        o += (a.nbytes + 15) & ~15
    h_in = _pinned.get("flow_in", o)
    hv = h_in.numpy()
    for a, off in zip(parts, offs):
        hv[off:off + a.nbytes] = a.reshape(-1).view(np.uint8)
    d_in = _scratch.get("flow_in", o, dev)
    d_in[:o].copy_(h_in[:o], non_blocking=True)
    base = d_in.data_ptr()
    p_cur = cur_dev.data_ptr() if cur_dev is not None else base + offs[3]
    # ---- out: [lattice | vectors | magnitudes | match | velocity] in one device block -------------------------------
    sizes = [g * 16, g * 16, g * 8, n_cur * 4, n_cur * 8]
    ooffs, total = [], 0
    for b in sizes:
        ooffs.append(total)
        total += (b + 15) & ~15
    d_out = _scratch.get("flow_out", total, dev)
    ob = d_out.data_ptr()
    check(lib.lidar_frame_flow(base + offs[2], n_prev, p_cur, n_cur, float(dt), float(gate), base + offs[0], nx,
                               base + offs[1], ny, float(radius), ob + ooffs[3], ob + ooffs[4], ob + ooffs[0],
                               ob + ooffs[1], ob + ooffs[2], _stream_ptr()))
    got = fetch("flow_out", d_out[:total], *extra)
    own = got[0].copy()                                   # one memcpy: the results live in a buffer the caller owns
    lattice = own[ooffs[0]:ooffs[0] + sizes[0]].view(np.float64).reshape(g, 2)
    vec = own[ooffs[1]:ooffs[1] + sizes[1]].view(np.float64).reshape(g, 2)
    mag = own[ooffs[2]:ooffs[2] + sizes[2]].view(np.float64)
    match = own[ooffs[3]:ooffs[3] + sizes[3]].view(np.int32)
    vel = own[ooffs[4]:ooffs[4] + sizes[4]].view(np.float32).reshape(n_cur, 2)
    return (lattice, vec, mag, match, vel) + tuple(a.copy() for a in got[1:])


def nearest_grid_cell(nodes_xy, grid_x, grid_y, density_flat=None, speed=None):
    """cKDTree(cell centres).query(nodes, k=1) on the rectilinear density grid (utils/visualization.py:306-314): returns
    device tensors (index int64 (G,), distance (G,)) and, with `density_flat` / `speed`, also (density_at, risk, risk_max)."""
    dev = require_cuda()
    nodes = _f64_dev(nodes_xy, dev).reshape(-1, 2)
    gx, gy = _f64_dev(grid_x, dev), _f64_dev(grid_y, dev)
    g = nodes.shape[0]
    index = torch.empty(g, dtype=torch.int64, device=dev)
    dist = torch.empty(g, dtype=torch.float64, device=dev)
    if density_flat is None:
        check(lib.lidar_nearest_grid_cell(_ptr(nodes), g, _ptr(gx), gx.numel(), _ptr(gy), gy.numel(), None, None, _ptr(index),
                                          _ptr(dist), None, None, None, _stream_ptr()))
        return index, dist
    dens, spd = _f64_dev(density_flat, dev).reshape(-1), _f64_dev(speed, dev).reshape(-1)
    if dens.numel() != gx.numel() * gy.numel() or spd.numel() != g:
        raise ValueError("density_flat must have nx*ny entries and speed one entry per node")
    at = torch.empty(g, dtype=torch.float64, device=dev)
    risk = torch.empty(g, dtype=torch.float64, device=dev)
    rmax = torch.zeros(1, dtype=torch.float64, device=dev)
    check(lib.lidar_nearest_grid_cell(_ptr(nodes), g, _ptr(gx), gx.numel(), _ptr(gy), gy.numel(), _ptr(dens), _ptr(spd),
                                      _ptr(index), _ptr(dist), _ptr(at), _ptr(risk), _ptr(rmax), _stream_ptr()))
    return index, dist, at, risk, rmax
