"""Multi-GPU modes of the hot path.  One process per GPU; torch.distributed is the
plumbing (rendezvous, symmetric-memory addresses), the data path is the C ABI.

  frames   independent frames shard across ranks as contiguous ranges — NO data-path collective.  The
           reference's analyze() calls are stateless per frame (its prev_positions slot is dead code,
           models/crowd_flow_model.py:16-17); the NEW frame-to-frame flow needs the centroids of the
           frame before the shard's first one, which the shard recomputes locally (1-frame halo).
  points   one oversized scan shards by points (`sharded_grid_density` / `ScanDensity`):
           calculate_grid_density (utils/data_processing.py:282-328) with the bbox MAX-reduced and the
           integer grid SUM-reduced across the ranks — bit-identical to the single-GPU grid.  ONE enqueue
           per call, no host round trip before the result is read back:
             "fused"  lidar_scan_density: a persistent cooperative kernel per rank does bbox, the bbox
                      exchange, the device-side np.arange parameters, the histogram, the two-shot grid
                      all-reduce over NVLink (multimem.ld_reduce / multimem.st through the NVSwitch when
                      the symmetric buffer has a multicast mapping, peer loads / stores otherwise) and
                      the density conversion.  torch symmetric memory only hands out the addresses.
             "nccl"   the checked fallback: lidar_scan_bbox_packed -> lidar_nccl_allreduce(MAX) ->
                      lidar_scan_hist -> lidar_nccl_allreduce(SUM) -> lidar_scan_finish on one stream.
Voxel downsample, DBSCAN, FPS and ball query do not shard by points (global neighbourhoods / sequential
dependence): they run as replicas over frames or batch elements.
"""
from __future__ import annotations

import ctypes as C
import time
from typing import Callable, Sequence

import numpy as np
import torch
import torch.distributed as dist


def frame_range(n_frames: int, rank: int, world: int) -> range:
    """Contiguous frame range of `rank` (sizes differ by at most one, earlier ranks get the extra)."""
    base, extra = divmod(n_frames, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def frame_range_with_halo(n_frames: int, rank: int, world: int) -> tuple[range, int | None]:
    """The shard's frames plus the index of the halo frame whose centroids seed frame_flow (None for
    the shard that starts the sequence)."""
    r = frame_range(n_frames, rank, world)
    halo = r.start - 1 if len(r) and r.start > 0 else None
    return r, halo


def _world(group=None) -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def allreduce_bbox(local_min: torch.Tensor, local_max: torch.Tensor, group=None):
    """Global per-axis min / max with ONE collective: MAX over the concatenation [-min, max]."""
    packed = torch.cat([-local_min, local_max])
    if _world(group)[1] > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.MAX, group=group)
    k = local_min.numel()
    return -packed[:k], packed[k:]


def allreduce_grid(counts: torch.Tensor, group=None) -> torch.Tensor:
    """SUM all-reduce of the integer density grid (in place)."""
    if _world(group)[1] > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts


def _host_logic_grid_density(points_shard, grid_size, group, local_bbox: Callable, local_hist: Callable):
    """The sharded algorithm with caller-supplied per-rank kernels and torch.distributed collectives: the form the
    world_size-2 gloo tests run on CPU tensors (tests/test_sharding_cpu.py).  The product path is ScanDensity."""
    lo, hi = local_bbox(points_shard)                       # +inf / -inf for an empty shard
    gmin, gmax = allreduce_bbox(lo.to(torch.float64), hi.to(torch.float64), group)
    packed = torch.cat([gmin, gmax]).cpu().numpy()
    gmin, gmax = packed[:2], packed[2:]
    if not np.all(np.isfinite(packed)):
        return None, None, None
    margin = grid_size * 2
    x_edges = np.arange(gmin[0] - margin, (gmax[0] + margin) + grid_size, grid_size)
    y_edges = np.arange(gmin[1] - margin, (gmax[1] + margin) + grid_size, grid_size)
    counts = allreduce_grid(local_hist(points_shard, x_edges, y_edges), group)
    density = (counts.to(torch.float64) / (grid_size * grid_size)).cpu().numpy()
    return (x_edges[:-1] + x_edges[1:]) / 2, (y_edges[:-1] + y_edges[1:]) / 2, density


class ScanDensity:
    """Reusable context of the point-sharded calculate_grid_density on CUDA: workspace, (symmetric) grid buffer,
    host-mapped descriptor, page-locked result staging.  Every rank of `group` constructs one (collective) and then
    calls it the same number of times.

    backend: "auto" (fused over symmetric memory when the rendezvous succeeds, else "nccl"), "fused", "nccl".
    solo=True builds a single-rank context inside a multi-rank job (measurements: the shard without its peers).
    """

    def __init__(self, device: torch.device | None = None, group=None, backend: str = "auto",
                 max_nx: int = 4096, max_ny: int = 4096, cap_cells: int = 1 << 20, solo: bool = False):
        from . import _capi, ops
        self._capi = _capi
        lib = _capi.lib
        self.device = device or ops.require_cuda()
        self.group = group
        self.rank, self.world = (0, 1) if solo else _world(group)     # solo: this rank's shard alone, no peers
        if self.world > 16:
            raise ValueError("ScanDensity: at most 16 ranks (one NVLink domain)")
        self.max_nx, self.max_ny = int(max_nx), int(max_ny)
        self.cap_cells = (int(cap_cells) + 3) & ~3
        dev = self.device
        self.ws = torch.zeros(lib.lidar_scan_workspace_bytes(), dtype=torch.uint8, device=dev)
        self.density = torch.empty(self.cap_cells, dtype=torch.float64, device=dev)
        self.gx = torch.empty(self.max_nx, dtype=torch.float64, device=dev)
        self.gy = torch.empty(self.max_ny, dtype=torch.float64, device=dev)
        self.desc_dev = torch.zeros(C.sizeof(_capi.ScanDesc), dtype=torch.uint8, device=dev)
        self.packed = torch.empty(4, dtype=torch.float64, device=dev)
        # page-locked + device-mapped: [ScanDesc (256 B) | gx | gy]
        self._h_bytes = 256 + 8 * (self.max_nx + self.max_ny)
        p = C.c_void_p()
        _capi.check(lib.lidar_host_alloc(self._h_bytes, C.byref(p)))
        self._h_ptr = p.value
        self._h_desc = _capi.ScanDesc.from_address(self._h_ptr)
        self._h_arr = np.ctypeslib.as_array((C.c_double * ((self._h_bytes - 256) // 8)).from_address(self._h_ptr + 256))
        self.epoch = 0
        self._pool = ops.ResultPool()
        self.comm = None          # ScanComm of the fused multi-rank path
        self.nccl = None          # lidar_nccl communicator of the fallback
        self.grid = None
        self.backend = "fused"
        if self.world > 1:
            want = backend
            if want in ("auto", "fused"):
                try:
                    self._init_symmetric()
                except Exception as e:
                    if want == "fused":
                        raise
                    self._symm_error = repr(e)
                    want = "nccl"
            # all ranks must agree (a rendezvous can fail on one rank only if it fails on all, but be explicit)
            flag = torch.tensor([1 if self.comm is not None else 0], device=dev, dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag.item()) == 0:
                self.comm = None
                want = "nccl" if backend == "auto" else want
            if self.comm is None:
                if want != "nccl":
                    raise RuntimeError("ScanDensity: symmetric memory is not available on every rank")
                self._init_nccl()
                self.backend = "nccl"
        if self.comm is None:
            self.grid = torch.zeros(self.cap_cells, dtype=torch.int32, device=dev)
        torch.cuda.synchronize(dev)

    # ---- plumbing ------------------------------------------------------------------------------
    def _init_symmetric(self):
        import torch.distributed._symmetric_memory as symm
        lib, _capi = self._capi.lib, self._capi
        nbytes = int(lib.lidar_scan_symm_bytes(self.cap_cells))
        t = symm.empty(nbytes, dtype=torch.uint8, device=self.device)
        t.zero_()
        grp = self.group if self.group is not None else dist.group.WORLD
        hdl = symm.rendezvous(t, group=grp)
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)          # every rank's buffer is zero before anybody's first kernel runs
        comm = _capi.ScanComm()
        comm.rank, comm.world, comm.symm_bytes = self.rank, self.world, nbytes
        ptrs = list(hdl.buffer_ptrs)
        for r in range(self.world):
            comm.peer_ptrs[r] = int(ptrs[r])
        mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
        comm.multicast_ptr = mc if mc else None
        self._symm_tensor, self._symm_handle = t, hdl
        self.comm = comm
        self.multicast = bool(mc)

    def _init_nccl(self):
        lib, _capi = self._capi.lib, self._capi
        if not lib.lidar_nccl_available():
            raise RuntimeError("ScanDensity: libnccl.so.2 not found")
        ident = (C.c_char * 128)()
        if self.rank == 0:
            _capi.check(lib.lidar_nccl_unique_id(ident))
        box = [bytes(ident.raw)]
        src = dist.get_global_rank(self.group, 0) if self.group is not None else 0
        dist.broadcast_object_list(box, src=src, group=self.group)
        ident = (C.c_char * 128).from_buffer_copy(box[0])
        h = C.c_void_p()
        _capi.check(lib.lidar_nccl_comm_init(ident, self.rank, self.world, C.byref(h)))
        self.nccl = h

    def close(self):
        lib = self._capi.lib
        if self.nccl is not None:
            lib.lidar_nccl_comm_destroy(self.nccl)
            self.nccl = None
        if self._h_ptr:
            self._h_arr = None
            self._h_desc = None
            lib.lidar_host_free(self._h_ptr)
            self._h_ptr = 0

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # ---- one call ------------------------------------------------------------------------------
    def enqueue(self, points: torch.Tensor, grid_size: float) -> None:
        """Enqueue the whole sharded density on the current stream (asynchronous; `result()` reads it back)."""
        from . import ops
        _capi, lib = self._capi, self._capi.lib
        fmt = ops.point_format(points)
        st = ops._stream_ptr()
        self.epoch += 1
        self._g = float(grid_size)
        n = points.shape[0]
        if self.backend == "fused":
            comm_ref = C.byref(self.comm) if self.comm is not None else None
            _capi.check(lib.lidar_scan_density(
                ops._ptr(points), fmt, n, self._g, self.max_nx, self.max_ny, self.cap_cells, ops._ptr(self.grid),
                ops._ptr(self.density), ops._ptr(self.gx), ops._ptr(self.gy), ops._ptr(self.desc_dev), self._h_ptr,
                comm_ref, self.epoch, ops._ptr(self.ws), self.ws.numel(), st))
        else:
            _capi.check(lib.lidar_scan_bbox_packed(ops._ptr(points), fmt, n, ops._ptr(self.packed), ops._ptr(self.ws),
                                                   self.ws.numel(), st))
            _capi.check(lib.lidar_nccl_allreduce(self.nccl, ops._ptr(self.packed), 4, _capi.NCCL_MAX_F64, st))
            _capi.check(lib.lidar_scan_hist(ops._ptr(points), fmt, n, ops._ptr(self.packed), self._g, self.max_nx,
                                            self.max_ny, self.cap_cells, ops._ptr(self.grid), ops._ptr(self.desc_dev), st))
            _capi.check(lib.lidar_nccl_allreduce(self.nccl, ops._ptr(self.grid), self.cap_cells, _capi.NCCL_SUM_I32, st))
            _capi.check(lib.lidar_scan_finish(ops._ptr(self.grid), ops._ptr(self.desc_dev), ops._ptr(self.density),
                                              ops._ptr(self.gx), ops._ptr(self.gy), st))

    def _wait_desc(self):
        """nx / ny / status of the call in flight.  Fused: the kernel publishes them in host-mapped memory as soon as
        the edges are known, long before it ends, so the read-back below queues up BEHIND the running kernel."""
        _capi = self._capi
        stream = torch.cuda.current_stream(self.device)
        if self.backend == "fused":
            h = self._h_desc
            t0 = time.perf_counter()
            spins = 0
            while h.pad != self.epoch:
                spins += 1
                if (spins & 0x3ff) == 0:
                    if stream.query():              # the kernel is gone: either it published, or it failed
                        if h.pad == self.epoch:
                            break
                        stream.synchronize()        # raises the CUDA error, if any
                        raise RuntimeError("lidar_scan_density finished without publishing its descriptor")
                    if time.perf_counter() - t0 > 30.0:
                        raise TimeoutError("lidar_scan_density: no descriptor after 30 s (is every rank calling?)")
            return _capi.ScanDesc.from_buffer_copy(bytes(C.string_at(self._h_ptr, C.sizeof(_capi.ScanDesc))))
        # three-enqueue form: the descriptor comes back with a copy of its own (first of two waits)
        nb = C.sizeof(_capi.ScanDesc)
        _capi.check(_capi.lib.lidar_copy_async(self._h_ptr, self.desc_dev.data_ptr(), nb, 0, stream.cuda_stream))
        stream.synchronize()
        return _capi.ScanDesc.from_buffer_copy(bytes(C.string_at(self._h_ptr, nb)))

    def result(self, fetch: bool = True):
        """(grid_x, grid_y, density) of the enqueued call as owned numpy arrays — the reference's return value
        (utils/data_processing.py:324-328); (None, None, None) for an empty scan (:297-298).
        `fetch=False` (a rank that does not need the arrays on its host: every rank's DEVICE copy is complete and
        identical) only waits for the call and checks its status; returns None."""
        _capi = self._capi
        d = self._wait_desc()
        stream = torch.cuda.current_stream(self.device)
        if not fetch:
            stream.synchronize()
            if d.status not in (0, _capi.SCAN_EMPTY):
                raise _capi.LidarError(int(d.status), "scan density grid exceeds the capacities")
            return None
        if d.status == _capi.SCAN_EMPTY:
            stream.synchronize()
            return None, None, None
        if d.status != 0:
            stream.synchronize()
            raise _capi.LidarError(int(d.status), f"scan density grid exceeds the capacities "
                                                  f"(max {self.max_nx} x {self.max_ny}, {self.cap_cells} cells)")
        nx, ny = int(d.nx), int(d.ny)
        # the density lands directly in a recycled page-locked buffer the caller then owns (ops.ResultPool: reused only
        # when no array referencing it is alive); the two short coordinate vectors go through the staging block
        dens = np.frombuffer(self._pool.take(8 * nx * ny), dtype=np.float64, count=nx * ny).reshape(nx, ny)
        base, st, cp = self._h_ptr + 256, stream.cuda_stream, _capi.lib.lidar_copy_async
        _capi.check(cp(base, self.gx.data_ptr(), 8 * nx, 0, st))
        _capi.check(cp(base + 8 * self.max_nx, self.gy.data_ptr(), 8 * ny, 0, st))
        _capi.check(cp(dens.ctypes.data, self.density.data_ptr(), 8 * nx * ny, 0, st))
        stream.synchronize()
        a = self._h_arr
        return a[:nx].copy(), a[self.max_nx:self.max_nx + ny].copy(), dens

    def __call__(self, points: torch.Tensor, grid_size: float):
        self.enqueue(points, grid_size)
        return self.result()


_contexts: dict = {}


def scan_context(device: torch.device, group=None, backend: str = "auto") -> ScanDensity:
    """The cached ScanDensity of (device, group, backend); constructing one is a collective over `group`."""
    key = (device.index, id(group) if group is not None else None, backend)
    ctx = _contexts.get(key)
    if ctx is None:
        ctx = _contexts[key] = ScanDensity(device, group, backend)
    return ctx


def sharded_grid_density(points_shard: torch.Tensor, grid_size: float, group=None,
                         local_bbox: Callable | None = None, local_hist: Callable | None = None,
                         backend: str = "auto"):
    """calculate_grid_density (utils/data_processing.py:282-328) of a scan whose points are spread over the
    ranks of `group`; `points_shard` is this rank's part ((n,4) float32 or (n,3) float64, CUDA).

    Returns (grid_x, grid_y, density) exactly like the reference, identical on every rank; density is
    counts / g² with bit-exact integer counts.  An empty shard is fine; an empty scan returns
    (None, None, None).  With `local_bbox` / `local_hist` the per-rank kernels are the caller's and the collectives
    are torch.distributed's (the CPU form of the algorithm, used by the gloo tests)."""
    if local_bbox is not None or local_hist is not None:
        return _host_logic_grid_density(points_shard, grid_size, group, local_bbox, local_hist)
    return scan_context(points_shard.device, group, backend)(points_shard, grid_size)


def run_frames_sharded(frames: Sequence, process: Callable, rank: int | None = None, world: int | None = None):
    """Apply `process(frame_index, frame)` to this rank's contiguous share of `frames`; returns
    {frame_index: result}.  No collective: results are gathered by the caller if it wants them."""
    r, w = _world()
    rank = r if rank is None else rank
    world = w if world is None else world
    return {i: process(i, frames[i]) for i in frame_range(len(frames), rank, world)}


def gather_results(local: dict, group=None) -> dict | None:
    """Gather the per-frame result dicts on rank 0 (small python objects, off the hot path)."""
    rank, world = _world(group)
    if world == 1:
        return dict(local)
    bucket = [None] * world if rank == 0 else None
    dist.gather_object(local, bucket, dst=0, group=group)
    if rank != 0:
        return None
    out: dict = {}
    for part in bucket:
        out.update(part)
    return out
