"""Multi-GPU modes of the hot path.  One process per GPU, torch.distributed for the
plumbing (NCCL over NVLink on the B200 box; gloo in the CPU tests).

  frames   independent frames shard across ranks as contiguous ranges — NO data-path collective.  The
           reference's analyze() calls are stateless per frame (its prev_positions slot is dead code,
           models/crowd_flow_model.py:16-17); the NEW frame-to-frame flow needs the centroids of the
           frame before the shard's first one, which the shard recomputes locally (1-frame halo).
  points   one oversized scan shards by points: (1) bbox = one MAX all-reduce over [-min, max],
           (2) every rank bins its points into the SAME edges, (3) one SUM all-reduce of the int32 grid.
           Integer sums are order independent, so the result is bit-identical to the single-GPU grid.
Voxel downsample, DBSCAN, FPS and ball query do not shard by points (global neighbourhoods / sequential
dependence): they run as replicas over frames or batch elements.
"""
from __future__ import annotations

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


def _cuda_bbox(points: torch.Tensor):
    from . import ops
    bb = ops.bbox(points)
    return bb[:2].clone(), bb[4:6].clone()


def _cuda_hist(points: torch.Tensor, x_edges: np.ndarray, y_edges: np.ndarray) -> torch.Tensor:
    from . import ops
    return ops.hist2d_points_counts(points, x_edges, y_edges)


def sharded_grid_density(points_shard: torch.Tensor, grid_size: float, group=None,
                         local_bbox: Callable = _cuda_bbox, local_hist: Callable = _cuda_hist):
    """calculate_grid_density (utils/data_processing.py:282-328) of a scan whose points are spread over
    the ranks of `group`; `points_shard` is this rank's part ((n,4) float32 or (n,3) float64, CUDA).

    Returns (grid_x, grid_y, density) exactly like the reference, identical on every rank; density is
    counts / g² with bit-exact integer counts.  An empty shard is fine; an empty scan returns
    (None, None, None)."""
    lo, hi = local_bbox(points_shard)                       # +inf / -inf for an empty shard
    # an empty scan needs no extra collective: its global max stays -inf
    gmin, gmax = allreduce_bbox(lo.to(torch.float64), hi.to(torch.float64), group)
    packed = torch.cat([gmin, gmax]).cpu().numpy()          # host round trip 1 of 2: the edges are np.arange
    gmin, gmax = packed[:2], packed[2:]
    if not np.all(np.isfinite(packed)):
        return None, None, None
    margin = grid_size * 2
    x_edges = np.arange(gmin[0] - margin, (gmax[0] + margin) + grid_size, grid_size)
    y_edges = np.arange(gmin[1] - margin, (gmax[1] + margin) + grid_size, grid_size)
    counts = local_hist(points_shard, x_edges, y_edges)
    counts = allreduce_grid(counts, group)
    # counts / g² in float64 is exact-rounded per element wherever it is evaluated; on the device it is one
    # kernel and one page-locked copy (round trip 2 of 2) instead of three host passes over the grid
    density_dev = counts.to(torch.float64) / (grid_size * grid_size)
    if density_dev.is_cuda:
        host = torch.empty(density_dev.shape, dtype=torch.float64, pin_memory=True)
        host.copy_(density_dev, non_blocking=True)
        torch.cuda.current_stream(density_dev.device).synchronize()
        density = host.numpy()
    else:
        density = density_dev.numpy()
    return (x_edges[:-1] + x_edges[1:]) / 2, (y_edges[:-1] + y_edges[1:]) / 2, density


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
