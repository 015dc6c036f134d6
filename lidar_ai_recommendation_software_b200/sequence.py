"""Frame sequences (BASELINE configs[3]): preprocess of frame f+1 overlaps the host-side steps of frame f.

One frame = variant-A/B preprocess (device-resident mode of `preprocess.run`) followed by the flow step of
`CrowdFlowModel` (it matches people against the previous frame, NEW op B.3).  A call waits on the device a few times
(the descriptor of the chained front, DBSCAN's counters, the centroids, the flow field) and copies 37 MB in, so
`SequenceRunner` runs whole frames on a few worker threads, each with its own CUDA stream (ctypes and torch release the
GIL while they wait): the copy-in of one frame, the kernels of another and the host-side steps of a third overlap.
The flow step of frame f needs the people POSITIONS of frame f-1 and nothing else of it, so it runs on frame f's
worker too, as soon as the worker of f-1 has published its positions; only the hand-over of the results to the caller
is in frame order.  (With the flow step on the caller's thread, that thread -- 1.0 ms per frame under load -- paced the
pipeline.)  Results are identical to the serial loop.
"""
from __future__ import annotations

from collections import deque
from concurrent.futures import Future, ThreadPoolExecutor
from typing import Iterable, Iterator

import torch

from . import ops
from . import preprocess as _pre
from .models.crowd_flow_model import CrowdFlowModel
from .utils.data_processing import extract_people_positions


class SequenceRunner:
    def __init__(self, variant: str = "B", workers: int = 2, dt: float = 0.1, gate: float = 1.5):
        self.variant, self.dt, self.gate = variant, float(dt), float(gate)
        self.workers = max(1, int(workers))
        self.model = CrowdFlowModel()
        self._dev = ops.require_cuda()
        self._streams = [torch.cuda.Stream(device=self._dev) for _ in range(self.workers)]
        self._pool = ThreadPoolExecutor(max_workers=self.workers) if self.workers > 1 else None

    def _preprocess(self, frame, slot: int) -> dict:
        with torch.cuda.device(self._dev), torch.cuda.stream(self._streams[slot]):
            if isinstance(frame, torch.Tensor) and not frame.is_cuda:
                frame = frame.to(self._dev, non_blocking=True)      # pinned host tensor -> device
            if self.variant == "B" and isinstance(frame, torch.Tensor) and frame.dtype == torch.float64:
                out = _pre.run_sequence_frame(frame)        # the whole frame behind one C call, outside the interpreter lock
            else:
                out = _pre.run(frame, variant=self.variant, host_arrays=False)
                self._streams[slot].synchronize()
        return out

    def _frame(self, frame, slot: int, mine: Future, before: Future | None, first_prev):
        """One frame on a worker thread: preprocess, the people positions (published for the NEXT frame's worker as soon
        as they are on the host), then the flow result against the PREVIOUS frame's positions.  The flow of a frame needs
        its predecessor's positions, not its predecessor's flow (CrowdFlowModel.sequence_step), so nothing here is serial
        except the order in which the caller is handed the results."""
        try:
            with torch.cuda.device(self._dev), torch.cuda.stream(self._streams[slot]):
                pd = self._preprocess(frame, slot)
                people = extract_people_positions(pd)
                mine.set_result(people)
        except BaseException as e:
            mine.set_exception(e)
            raise
        prev = before.result() if before is not None else first_prev
        with torch.cuda.device(self._dev), torch.cuda.stream(self._streams[slot]):
            res, _ = self.model.sequence_step(pd, prev, dt=self.dt, gate=self.gate, people_positions=people)
        return pd, res

    def run(self, frames: Iterable) -> Iterator[tuple[dict, dict]]:
        """Yields (processed_data, flow_result) per frame, in order.  `frames` yields (n,3) float64 arrays
        (numpy, pinned host tensors or CUDA tensors).  Identical to calling `preprocess.run` and
        `self.model.analyze_sequence_frame` frame by frame; the model's state (`prev_positions`) is carried across calls."""
        if self._pool is None:
            for f in frames:
                pd = self._preprocess(f, 0)
                yield pd, self.model.analyze_sequence_frame(pd, dt=self.dt, gate=self.gate)
            return
        pending: deque = deque()
        before, k = None, 0
        first_prev = self.model.prev_positions
        for f in frames:
            mine: Future = Future()
            pending.append((self._pool.submit(self._frame, f, k % self.workers, mine, before, first_prev), mine))
            before = mine
            k += 1
            if len(pending) >= self.workers:
                yield self._hand_over(pending.popleft())
        while pending:
            yield self._hand_over(pending.popleft())

    def _hand_over(self, item):
        job, positions = item
        pd, res = job.result()
        people = positions.result()
        self.model.prev_positions = people if len(people) else None      # as analyze_sequence_frame leaves it
        if "matches" in res:
            self.model.flow_vectors = res["flow_vectors"]
        return pd, res

    def close(self) -> None:
        if self._pool is not None:
            self._pool.shutdown(wait=True)
            self._pool = None
