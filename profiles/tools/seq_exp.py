"""What bounds the sequence pipeline: preprocess workers alone (with / without the 37 MB copy-in), the ordered flow stage
alone, and both, for several worker counts."""
import sys, time, numpy as np, torch
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, '.')
from lidar_ai_recommendation_software_b200 import ops, preprocess, synth
from lidar_ai_recommendation_software_b200.sequence import SequenceRunner
dev = torch.device('cuda', 0)
pool = [synth.ring_sequence_frame(i) for i in range(4)]
pinned = [torch.from_numpy(np.ascontiguousarray(p[:, :3], dtype=np.float64)).pin_memory() for p in pool]
resident = [p.to(dev) for p in pinned]
N = 240
def run_workers(frames, workers):
    r = SequenceRunner(variant="B", workers=workers)
    futs = [r._pool.submit(r._preprocess, frames[i % 4], i % workers) for i in range(2 * workers)]
    [f.result() for f in futs]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    futs = [r._pool.submit(r._preprocess, frames[i % 4], i % workers) for i in range(N)]
    outs = [f.result() for f in futs]
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    r.close()
    return N / dt, outs
for w in (2, 3, 4, 6):
    a, _ = run_workers(pinned, w)
    b, outs = run_workers(resident, w)
    print(f"preprocess only, {w} workers: {a:7.1f} fps with copy-in, {b:7.1f} fps resident", flush=True)
# the ordered stage alone, on finished preprocess results
r = SequenceRunner(variant="B", workers=2)
pds = outs[:8]
for pd in pds[:4]: r.model.analyze_sequence_frame(pd, dt=0.1)
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(N): r.model.analyze_sequence_frame(pds[i % 8], dt=0.1)
torch.cuda.synchronize(); print(f"flow stage alone: {N / (time.perf_counter() - t0):7.1f} fps", flush=True)
r.close()
for w in (3, 4, 6):
    r = SequenceRunner(variant="B", workers=w)
    list(r.run([pinned[i % 4] for i in range(2 * w)]))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in r.run(pinned[i % 4] for i in range(N)): pass
    torch.cuda.synchronize(); print(f"whole pipeline, {w} workers: {N / (time.perf_counter() - t0):7.1f} fps", flush=True)
    r.close()
