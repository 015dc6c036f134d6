"""Frame sequences (BASELINE configs[3]): preprocess of frame f+1 overlaps the host-side steps of frame f.

One frame = variant-A/B preprocess (device-resident mode of `preprocess.run`) followed by
`CrowdFlowModel.analyze_sequence_frame`, which needs the frames IN ORDER (it matches people against the
previous frame, NEW op B.3).  The preprocess of different frames is independent; a call still waits on the
device twice (the descriptor of the chained front, DBSCAN's counters) and copies 37 MB in, so
`SequenceRunner` runs it on a few worker threads, each with its own CUDA stream (ctypes and torch release the
GIL while they wait): the copy-in of one frame, the kernels of another and the host-side steps of a third
overlap, and the results are fed to the flow model in frame order on the caller's thread.
Results are identical to the serial loop.
"""
from __future__ import annotations

from collections import deque
from concurrent.futures import ThreadPoolExecutor
from typing import Iterable, Iterator

import torch

from . import ops
from . import preprocess as _pre
from .models.crowd_flow_model import CrowdFlowModel


class SequenceRunner:
    def __init__(self, variant: str = "B", workers: int = 2, dt: float = 0.1, gate: float = 1.5):
        self.variant, self.dt, self.gate = variant, float(dt), float(gate)
        self.workers = max(1, int(workers))
        self.model = CrowdFlowModel()
        self._dev = ops.require_cuda()
        self._streams = [torch.cuda.Stream(device=self._dev) for _ in range(self.workers)]
        # The ordered stage (centroids, match, flow field: ~0.1 ms of small kernels and two waits per frame) runs on a
        # HIGH-PRIORITY stream: on an ordinary one its kernels queue behind whatever the workers have in flight -- each of
        # their kernels fills the device -- and the stage, which is serial, then paces the whole pipeline.
        self._flow_stream = torch.cuda.Stream(device=self._dev, priority=-1)
        self._pool = ThreadPoolExecutor(max_workers=self.workers) if self.workers > 1 else None

    def _preprocess(self, frame, slot: int) -> dict:
        with torch.cuda.device(self._dev), torch.cuda.stream(self._streams[slot]):
            if isinstance(frame, torch.Tensor) and not frame.is_cuda:
                frame = frame.to(self._dev, non_blocking=True)      # pinned host tensor -> device
            out = _pre.run(frame, variant=self.variant, host_arrays=False)
            self._streams[slot].synchronize()
        return out

    def run(self, frames: Iterable) -> Iterator[tuple[dict, dict]]:
        """Yields (processed_data, flow_result) per frame, in order.  `frames` yields (n,3) float64 arrays
        (numpy, pinned host tensors or CUDA tensors)."""
        if self._pool is None:
            for f in frames:
                pd = self._preprocess(f, 0)
                yield pd, self.model.analyze_sequence_frame(pd, dt=self.dt, gate=self.gate)
            return
        pending: deque = deque()
        k = 0
        for f in frames:
            pending.append(self._pool.submit(self._preprocess, f, k % self.workers))
            k += 1
            if len(pending) >= self.workers:
                yield self._flow(pending.popleft().result())
        while pending:
            yield self._flow(pending.popleft().result())

    def _flow(self, pd: dict) -> tuple[dict, dict]:
        # the worker has synchronised its stream: everything in `pd` is complete and may be read from any stream
        with torch.cuda.stream(self._flow_stream):
            return pd, self.model.analyze_sequence_frame(pd, dt=self.dt, gate=self.gate)

    def close(self) -> None:
        if self._pool is not None:
            self._pool.shutdown(wait=True)
            self._pool = None
