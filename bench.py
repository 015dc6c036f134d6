#!/usr/bin/env python
"""Benchmark of the LiDAR hot path on B200 — BASELINE.json's metric:
Mpoints/s through voxelise + density, % of the HBM roofline, CPU reference alongside.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--points P]

Workload (BASELINE.json configs[1]): a 1 M-point synthetic crowd frame (float4 x,y,z,intensity),
0.05 m voxel downsample + 0.5 m calculate_grid_density histogram of the same points.  One step =
one batch of --frames-per-step frames per GPU (512: 20 steps keep the timed region above 0.5 s).
N > 1: independent frames shard across ranks (weak scaling, no data-path collective); launched by
torchrun, one rank per GPU.

The JSON line carries
  value      whole-job Mpoints/s, frames resident in HBM when the timed region starts (CUDA events)
  e2e        the same through the host-buffer API under the drop-in contract: PAGEABLE numpy in ->
             numpy arrays the caller OWNS out, every copy timed; e2e.streaming = page-locked in,
             views out (a sensor driver's zero-copy mode); e2e.streaming_voxels_only without the
             per-point outputs
  roofline   dominant kernel: algorithmic bytes / measured launch duration vs the measured HBM peak
  cpu_baseline  the CPU oracle (numpy restatement; the reference has no voxel op — "port") timed on
             the host cores of this box, bounded sample
  extra      extra.pcie = host <-> device copy rates with all ranks copying at once; the other named configs in the same run: extra.scan50m (configs[4], 50 M points sharded
             by points, fused NVLink all-reduce), extra.seq (configs[3], 300-frame 128-beam
             sequence, frames sharded), extra.sa (configs[2], set abstraction, N = 1 only)
`--impl reference` times that CPU path as the line's own value (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
# The contract is ONE JSON line on stdout, but libraries write there too (torch's NCCL process group announces
# "NCCL version ..." on fd 1 at the first collective).  Everything that is not the result line goes to stderr.
_RESULT_FD = os.dup(1)
os.dup2(2, 1)


def emit(line: str) -> None:
    os.write(_RESULT_FD, (line + "\n").encode())

VOXEL = 0.05
GRID = 0.5
EXTENT = 50.0
POOL = 16          # distinct frames cycled through: 16 x 16 MB = 256 MB > 126 MB of L2
KERNELS = ["k_frame_prep", "k_frame_mark", "k_frame_scan", "k_frame_rank", "k_frame_finalize"]


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel: str):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/ncu_traffic.json)."""
    p = ROOT / "profiles" / "ncu_traffic.json"
    if not p.exists():
        return None
    d = json.loads(p.read_text()).get(kernel)
    return None if d is None else d["dram_read_bytes_per_launch"] + d["dram_write_bytes_per_launch"]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_oracle_step(frame: np.ndarray):
    """The CPU restatement of one step: numpy voxel downsample (Appendix B.1) + the reference's
    calculate_grid_density arithmetic (utils/data_processing.py:282-328) on the same points."""
    from oracle import new_ops, ref_path
    xyz = frame[:, :3].astype(np.float64)
    new_ops.voxel_downsample(frame, VOXEL)
    xr = (xyz[:, 0].min(), xyz[:, 0].max())
    yr = (xyz[:, 1].min(), xyz[:, 1].max())
    ref_path.calculate_grid_density(xyz[:, :2], xr, yr, GRID)


def time_cpu(frame: np.ndarray, budget_s: float = 12.0, max_reps: int = 5):
    reps, best, t_all = 0, float("inf"), time.perf_counter()
    while reps < max_reps and (time.perf_counter() - t_all) < budget_s:
        t0 = time.perf_counter()
        cpu_oracle_step(frame)
        best = min(best, time.perf_counter() - t0)
        reps += 1
    return best, reps


def _ref_worker(seed, n, steps, warmup, barrier, q):
    """One host core: its own frame, `warmup` untimed + `steps` timed oracle steps after a common start."""
    from lidar_ai_recommendation_software_b200 import synth
    frame = synth.crowd_frame(n, seed=seed, extent=EXTENT)
    for _ in range(warmup):
        cpu_oracle_step(frame)
    barrier.wait()
    for _ in range(steps):
        cpu_oracle_step(frame)
    q.put(time.time())


def run_reference(args, rank: int):
    """--impl reference: the reference-side CPU path on ALL of this box's host cores (rank 0 only).

    The reference has no voxel op, so the CPU arm is the oracle port (numpy restatement of Appendix B.1)
    plus the reference's own calculate_grid_density arithmetic; numpy runs it on one thread, so every
    core gets its own frame (independent frames, exactly how the GPU arm shards them)."""
    if rank != 0:
        return
    import multiprocessing as mp
    n = args.points
    procs = max(1, os.cpu_count() or 1)
    steps = max(1, min(args.steps, 4))
    warmup = min(args.warmup, 1)
    ctx = mp.get_context("fork")
    barrier = ctx.Barrier(procs + 1)
    q = ctx.Queue()
    ws = [ctx.Process(target=_ref_worker, args=(1000 + i, n, steps, warmup, barrier, q)) for i in range(procs)]
    for w in ws:
        w.start()
    barrier.wait()
    t0 = time.time()
    ends = [q.get() for _ in ws]
    for w in ws:
        w.join()
    dt = max(ends) - t0
    val = n * steps * procs / dt / 1e6
    line = {
        "impl": "reference", "metric": "Mpoints/s voxelize+density", "value": val, "unit": "Mpoints/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": dt / steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{n}-point crowd frame, {VOXEL} m voxel downsample + {GRID} m "
                               "calculate_grid_density histogram (BASELINE configs[1])",
                   "points_per_frame": n, "voxel_m": VOXEL, "grid_m": GRID,
                   "frames_per_step": procs},
        "cpu_baseline": {"value": val, "unit": "Mpoints/s", "cores": procs, "kind": "port",
                         "sample": f"{steps} steps x {procs} frames of {n} points, one frame per host core in "
                                   "parallel processes; numpy restatement of voxel downsample (absent upstream) "
                                   "+ the reference's np.histogram2d grid density"},
        "e2e": {"value": val, "unit": "Mpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(json.dumps(line))


def e2e_leg(ops, torch, dist, world, dev, n, frames, steps, threads, slots, mode):
    """End to end through the public host-buffer API, host wall clock, whole-job throughput (max over ranks).

    mode "api"        the drop-in contract: caller-owned PAGEABLE numpy frame in, fresh numpy arrays the
                      caller owns out.  `threads` Python threads drive one HostFramePipeline each, the way Streamlit
                      sessions drive the reference (one script thread per session); the C ABI stages with its
                      parallel host memcpy and reads back exactly what each frame produced (two-stage).
    mode "streaming"  a sensor driver that owns page-locked buffers: pinned frame in, views of the slot's pinned
                      result block out (`collect(copy=False)`), one thread, `slots` frames in flight.
    mode "streaming_voxels_only"  the same without the per-point outputs (LIDAR_HOST_NO_PER_POINT)."""
    api = mode == "api"
    kw = dict(max_points=n, voxel_size=VOXEL, grid_size=GRID, max_key_space=1 << 28, max_nx=256, max_ny=256)
    if api:
        pipes = [ops.HostFramePipeline(slots=2, two_stage=True, **kw) for _ in range(threads)]
        srcs = frames                                               # plain numpy arrays (pageable)
    else:
        pipes = [ops.HostFramePipeline(slots=slots, two_stage=False, per_point_outputs=(mode == "streaming"), **kw)]
        srcs = [torch.from_numpy(f).pin_memory() for f in frames]
        threads = 1
    torch.cuda.synchronize()
    for p_ in pipes:
        for w in range(2):
            p_.process(srcs[w % len(srcs)])
    d2h = pipes[0]._last_d2h if api else pipes[0].d2h_bytes(n)
    per_thread = max(2, steps // threads)
    total = per_thread * threads
    checks = [0] * threads

    def drive(t):
        hp = pipes[t]
        torch.cuda.set_device(dev)
        inflight, out = 0, None
        depth = len(hp.slots)
        for s in range(per_thread):
            if inflight == depth:
                out = hp.collect(copy=api)
                inflight -= 1
            hp.submit(srcs[(t + s) % len(srcs)])
            inflight += 1
        while inflight:
            out = hp.collect(copy=api)
            inflight -= 1
        checks[t] = int(out["counts"].sum())

    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if threads == 1:
        drive(0)
    else:
        ths = [threading.Thread(target=drive, args=(t,)) for t in range(threads)]
        for th in ths:
            th.start()
        for th in ths:
            th.join()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert all(c == n for c in checks), checks
    te = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    for p_ in pipes:
        p_.close()
    return {"value": n * total * world / float(te.item()) / 1e6, "unit": "Mpoints/s", "h2d_bytes_per_step": 16 * n,
            "d2h_bytes_per_step": int(d2h), "steps": total, "threads": threads}


def measure_pcie(torch, dist, dev, rank, world):
    """Host <-> device copy rates with EVERY rank copying at the same time (page-locked memory from the C ABI, CUDA
    events): what the end-to-end legs can at best move on this box.  Whole-job GB/s = sum over ranks."""
    from lidar_ai_recommendation_software_b200 import _capi, ops
    nb = 128 << 20
    pin_a, pin_b = ops._PinnedBlock(nb), ops._PinnedBlock(nb)
    d_a = torch.empty(nb, dtype=torch.uint8, device=dev)
    d_b = torch.empty(nb, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    cp = _capi.lib.lidar_copy_async

    def timed(h2d, d2h, reps=4):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_event(e0)
        s2.wait_event(e0)
        for _ in range(reps):
            if h2d:
                cp(d_a.data_ptr(), pin_a.ptr, nb, 1, s1.cuda_stream)
            if d2h:
                cp(pin_b.ptr, d_b.data_ptr(), nb, 0, s2.cuda_stream)
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)
        e1.record()
        torch.cuda.synchronize()
        return reps * nb / (e0.elapsed_time(e1) * 1e-3) / 1e9      # GB/s per direction

    timed(True, True, 1)
    rates = torch.tensor([timed(True, False), timed(False, True), timed(True, True)], dtype=torch.float64, device=dev)
    lo = rates.clone()
    if world > 1:
        dist.all_reduce(rates, op=dist.ReduceOp.SUM)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    pin_a.free()
    pin_b.free()
    if rank != 0:
        return None
    r, m = rates.tolist(), lo.tolist()
    return {"h2d_GBs_total": r[0], "d2h_GBs_total": r[1], "both_GBs_per_direction_total": r[2],
            "h2d_GBs_slowest_rank": m[0], "d2h_GBs_slowest_rank": m[1], "both_GBs_per_direction_slowest_rank": m[2],
            "note": f"{world} rank(s) copying concurrently, 4 x 128 MB per direction per rank, page-locked host memory"}


def run_extras(args, torch, dist, dev, rank, world):
    """The other named shapes of BASELINE.json (configs[2], [3], [4]) in the same run, so that the driver's N = 1/2/4/8
    records carry them: extra.sa (N = 1 only), extra.seq (frames sharded), extra.scan50m (points sharded)."""
    import types

    import bench_configs as bc
    extra = {}

    def guarded(name, fn):
        t0 = time.perf_counter()
        try:
            out = fn()
        except Exception as e:          # an extra must never take the headline line down with it
            out = {"error": f"{type(e).__name__}: {e}"[:400]}
            if world > 1:
                raise                    # ... but a rank that skips a collective would hang the others: fail loudly
        if rank == 0 and out is not None:
            out["bench_wall_s"] = time.perf_counter() - t0
            extra[name] = out

    if "pcie" in args.extras:
        guarded("pcie", lambda: measure_pcie(torch, dist, dev, rank, world))
    if "scan" in args.extras:
        a = types.SimpleNamespace(points=args.scan_points, host_shards=8, reps=10)
        guarded("scan50m", lambda: bc.run_scan(a, torch, dev, rank, world, dist))
    if "seq" in args.extras:
        # preprocess worker threads per rank: three when the rank has the CPUs for them (each worker drives a stream and
        # waits on the device twice per frame), two when eight ranks share a 32-vCPU host
        local_world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))
        per_rank = (os.cpu_count() or 1) // local_world      # the per-frame path is bound by its host side: a worker per ~3 vCPUs
        workers = 5 if per_rank >= 12 else 4 if per_rank >= 8 else 3 if per_rank >= 6 else 2
        a = types.SimpleNamespace(frames=args.seq_frames, pool=4, rings=128, azimuth=20480, dropin=False, workers=workers)
        guarded("seq", lambda: bc.run_seq(a, torch, dev, rank, world, dist))
    if "sa" in args.extras and world == 1:
        a = types.SimpleNamespace(reps=20)
        guarded("sa", lambda: bc.run_sa(a, torch, dev, rank, world))
    return extra


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--points", type=int, default=1_000_000)
    ap.add_argument("--frames-per-step", type=int, default=512,
                    help="frames per GPU in one step (one step = one batch of frames): 20 steps x 512 frames keep the "
                         "timed region above half a second, long enough for the clock sampler")
    ap.add_argument("--e2e-steps", type=int, default=240, help="frames per rank in each end-to-end leg")
    ap.add_argument("--e2e-slots", type=int, default=3, help="frames in flight in the streaming end-to-end legs")
    ap.add_argument("--e2e-threads", type=int, default=3, help="driver threads of the drop-in end-to-end leg")
    ap.add_argument("--extras", default="pcie,scan,seq,sa", help="comma list of the extra configs to measure ('' = none)")
    ap.add_argument("--scan-points", type=int, default=50_000_000)
    ap.add_argument("--seq-frames", type=int, default=300)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--ctas-per-sm", type=int, default=0, help="frame kernel grid cap (0 = library default)")
    ap.add_argument("--streams", type=int, default=0,
                    help="independent frame pipelines in flight (frames round-robin over CUDA streams)")
    ap.add_argument("--mode", default="fused", choices=["fused", "partitioned", "multikernel"],
                    help="frame back end: the persistent fused kernel (occupancy bitmap in L2), the persistent partitioned "
                         "kernel (MSD radix partition by voxel key, occupancy bits in shared memory), or five dependent kernels")
    ap.add_argument("--fused-threads", type=int, default=0)
    ap.add_argument("--fused-ctas-per-sm", type=int, default=0)
    ap.add_argument("--fused-smem-kb", type=int, default=0)
    ap.add_argument("--fused-plain-launch", action="store_true",
                    help="experiment: ordinary instead of cooperative launch (lets frames of different streams overlap)")
    ap.add_argument("--pdl", type=int, default=-1,
                    help="programmatic dependent launch of the fused kernel (2/1/0; default: library setting)")
    ap.add_argument("--l2-persist-mb", type=int, default=24,
                    help="pin this many MB of the occupancy groups in L2 (access policy window); 0 = off")
    ap.add_argument("--scan-order", type=int, default=0,
                    help="1: the scan-order variant of the fused kernel (run-length aggregation, summary-bitmap scan)")
    ap.add_argument("--streaming", type=int, default=1,
                    help="fused back end, one stream: ordinary launch + programmatic dependent launch for the "
                         "device-resident leg (frames back to back on one stream); 0 = cooperative launches")
    args = ap.parse_args()
    args.extras = [x for x in args.extras.split(",") if x]
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    from lidar_ai_recommendation_software_b200 import ops, synth

    assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # one rank per GPU: each rank lives on the CPUs / memory of ITS GPU's NUMA node before anything page-locked exists
    numa = ops.bind_to_device_numa(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n = args.points
    F = max(1, args.frames_per_step)
    # every rank owns its own frames (frames shard across GPUs: weak scaling, no collective)
    host_frames = [synth.crowd_frame(n, seed=rank * 1000 + s, extent=EXTENT) for s in range(POOL)]
    frames = [torch.from_numpy(f).to(dev) for f in host_frames]
    from lidar_ai_recommendation_software_b200 import _capi
    if args.ctas_per_sm:
        _capi.check(_capi.lib.lidar_frame_set_ctas_per_sm(args.ctas_per_sm))
    fused = args.mode in ("fused", "partitioned")
    if args.fused_plain_launch:
        _capi.check(_capi.lib.lidar_frame_set_fused_plain_launch(1))
    if args.scan_order:
        ops.set_frame_scan_order(True)
    if args.pdl >= 0:
        _capi.check(_capi.lib.lidar_frame_set_fused_pdl(args.pdl))
    if args.l2_persist_mb > 0:
        try:        # an optimisation, not a requirement: a device that refuses the carve-out runs without it
            _capi.check(_capi.lib.lidar_frame_set_fused_l2_persist(args.l2_persist_mb << 20))
        except _capi.LidarError as e:
            print(f"bench: L2 persistence not available ({e}); continuing without it", file=sys.stderr)
            args.l2_persist_mb = 0
    frame_mode = {"fused": ops.FRAME_FUSED, "partitioned": ops.FRAME_PARTITIONED, "multikernel": ops.FRAME_MULTIKERNEL}[args.mode]
    ops.set_frame_mode(frame_mode, args.fused_threads, args.fused_ctas_per_sm, args.fused_smem_kb)
    # the fused kernel fills the device by itself (frames of other streams would only queue behind it);
    # the five-kernel path leaves gaps that frames on other streams fill
    S = args.streams if args.streams > 0 else (1 if fused else 4)
    pipes = [ops.FramePipeline(max_points=n, voxel_size=VOXEL, grid_size=GRID, max_key_space=1 << 28,
                               max_nx=256, max_ny=256, device=dev, scan_order=bool(args.scan_order)) for _ in range(S)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(S)]
    pipe = pipes[0]
    streaming = bool(fused and S == 1 and args.streaming and not args.fused_plain_launch and args.pdl < 0)
    if streaming:
        # the frames are resident (uploaded and synchronised above) long before they are enqueued
        ops.set_frame_streaming(True, inputs_complete=True)
    torch.cuda.synchronize()

    def run_frames(count):
        """`count` frames, round-robin over the S pipelines/streams; returns after enqueueing, with the
        current stream made to wait for all of them."""
        cur = torch.cuda.current_stream()
        if S == 1:
            for s_ in range(count):
                pipe.enqueue(frames[s_ % POOL])
            return
        for st in streams:
            st.wait_stream(cur)
        for s_ in range(count):
            k = s_ % S
            with torch.cuda.stream(streams[k]):
                pipes[k].enqueue(frames[s_ % POOL])
        for st in streams:
            cur.wait_stream(st)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput: K steps of F frames each --------------------------------------
    clocks = ClockSampler(local_rank)
    clocks.start()
    for _ in range(args.warmup):
        run_frames(F)
    torch.cuda.synchronize()
    res = pipe.result()
    v_over_n = res.n_voxels / n
    barrier()
    clocks.rows.clear()   # keep only samples taken during the timed region
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        run_frames(F)
    ev1.record()
    barrier()
    clk = clocks.stop()
    ms = ev0.elapsed_time(ev1)
    for p_ in pipes:
        p_.result()  # raises if any frame overflowed its capacities
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = n * F * args.steps * world / (ms * 1e-3) / 1e6

    if streaming:
        ops.set_frame_streaming(False)   # the remaining legs (several pipelines / streams) use cooperative launches
    # ---- per-kernel device time (CUDA events on the launching stream) --------------------------
    ksteps = 50
    per_kernel = np.zeros(5)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(6)] for _ in range(ksteps)]
    for s in range(ksteps):
        pipe.enqueue(frames[s % POOL], events=evs[s])
    torch.cuda.synchronize()
    for s in range(ksteps):
        for k in range(5):
            per_kernel[k] += evs[s][k].elapsed_time(evs[s][k + 1])
    per_kernel /= ksteps  # ms
    hbm_peak, peak_src = peaks()
    step_ms = ms / args.steps
    frame_ms = step_ms / F
    frame_bytes = (20.0 + 20.0 * v_over_n) * n      # 20 + 20*V/N bytes per point
    phases = None
    if fused:
        # one launch per frame: the kernel's algorithmic bytes are the frame's; phases from %globaltimer
        kms = float(per_kernel.sum())
        last = pipe.result()
        ph = [int(x) for x in last.desc.trace_ns]
        assert ph[15] > 0, "the fused kernel did not run"
        names = (["load_bbox", "bar1", "desc", "keys_slots_claim", "bar2", "sort_scatter", "bar3", "owner_pass", "bar4",
                  "ranks_records", "bar5", "finalize"] if args.mode == "partitioned" else
                 ["load_bbox", "bar1", "desc", "mark", "bar2", "scan_popc", "scan_wait", "scan_prefix", "bar3",
                  "rank", "bar4", "clean_finalize"])
        phases = {k + "_us": ph[i] / 1e3 for i, k in enumerate(names)}
        phases["ctas"] = ph[15]
        gbs = frame_bytes / (frame_ms * 1e-3) / 1e9
        kname = "k_frame_part" if args.mode == "partitioned" else "k_frame_fused"
        roofline = {"bound": "hbm", "kernel": kname, "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                    "frac": gbs / hbm_peak, "traffic": ncu_traffic(kname), "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": frame_bytes, "launch_ms": frame_ms,
                    "launch_ms_alone": kms,
                    "note": "launch_ms = average duration of the one launch per frame over the timed region (CUDA "
                            "events around all steps / frames); launch_ms_alone = one frame alone on an idle device, "
                            "CUDA events around a single cooperative launch (includes the launch gap)"}
        roofline_step = {"bytes_per_point": frame_bytes / n, "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                         "frac": gbs / hbm_peak, "kernel_ms": {kname: kms}, "phases_of_one_frame": phases}
    dom = int(np.argmax(per_kernel))
    # algorithmic bytes per point of each kernel
    n_groups = float(res.desc.key_space) / 224.0          # one 32 B occupancy group per 224 voxel cells
    alg = {
        "k_frame_prep": 16.0 * n,                               # read every point once
        "k_frame_mark": (16.0 + 4.0) * n,                       # read point, write voxel key
        "k_frame_scan": n_groups * (32.0 + 4.0),                # read every group, write its prefix
        "k_frame_rank": (16.0 + 4.0 + 4.0) * n + 32.0 * res.n_voxels,   # point, key, inverse + voxel record
        "k_frame_finalize": 4.0 * res.n_voxels,                 # sweep of the member counters
    }
    if not fused:
        dom_name = KERNELS[dom]
        dom_gbs = alg[dom_name] / (per_kernel[dom] * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": dom_name, "achieved": dom_gbs, "peak": hbm_peak, "unit": "GB/s",
                    "frac": dom_gbs / hbm_peak, "traffic": ncu_traffic(dom_name), "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg[dom_name], "launch_ms": float(per_kernel[dom])}
        roofline_step = {"bytes_per_point": frame_bytes / n, "achieved": frame_bytes / (frame_ms * 1e-3) / 1e9,
                         "peak": hbm_peak, "unit": "GB/s", "frac": frame_bytes / (frame_ms * 1e-3) / 1e9 / hbm_peak,
                         "kernel_ms": {k: float(v) for k, v in zip(KERNELS, per_kernel)}}
    for p_ in pipes:
        del p_
    pipes.clear()

    # ---- end to end through the host-buffer API ------------------------------------------------
    e2e_steps = max(8, args.e2e_steps)
    e2e = e2e_leg(ops, torch, dist, world, dev, n, host_frames[:4], e2e_steps, max(1, args.e2e_threads), args.e2e_slots, "api")
    e2e["api"] = (f"pageable numpy in, owned numpy out: ops.HostFramePipeline.submit / collect(copy=True), "
                  f"{e2e['threads']} caller threads with one pipeline (2 slots) each; per frame one parallel staging "
                  "memcpy, lidar_frame_voxel_density_host_begin, lidar_frame_host_fetch (exact sizes), parallel copy-out")
    e2e["numa_node"], e2e["cpus"] = numa
    stream_leg = e2e_leg(ops, torch, dist, world, dev, n, host_frames[:4], e2e_steps, 1, args.e2e_slots, "streaming")
    stream_leg["api"] = (f"pinned numpy in, views of the pinned result block out: submit / collect(copy=False), one thread, "
                         f"{args.e2e_slots} slots in flight, one frame-sized copy-out (the round-1 e2e figure)")
    e2e["streaming"] = stream_leg
    vox_leg = e2e_leg(ops, torch, dist, world, dev, n, host_frames[:4], e2e_steps, 1, args.e2e_slots, "streaming_voxels_only")
    vox_leg["api"] = "as streaming, LIDAR_HOST_NO_PER_POINT: voxel_key / inverse stay on the device"
    e2e["streaming_voxels_only"] = vox_leg

    # ---- the other named configs -----------------------------------------------------------------
    del frames
    torch.cuda.empty_cache()
    extra = run_extras(args, torch, dist, dev, rank, world) if args.extras else {}

    # ---- CPU baseline (rank 0, N=1 only) -------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        best, reps = time_cpu(host_frames[0])
        cpu = {"value": n / best / 1e6, "unit": "Mpoints/s", "cores": 1, "kind": "port",
               "sample": f"best of {reps} x one {n}-point frame: numpy restatement of voxel downsample "
                         "(op absent upstream) + the reference's calculate_grid_density arithmetic; "
                         f"single-threaded numpy, host has {os.cpu_count()} logical cores"}

    if rank == 0:
        line = {
            "metric": "Mpoints/s voxelize+density", "value": value, "unit": "Mpoints/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64 decisions / i64 fixed-point sums on f32 points",
            "data": "synthetic",
            "config": {"workload": f"{n}-point crowd frame, {VOXEL} m voxel downsample + "
                                   f"{GRID} m calculate_grid_density histogram (BASELINE configs[1])",
                       "points_per_frame": n, "voxel_m": VOXEL, "grid_m": GRID, "frames_per_step": F,
                       "ms_per_frame": frame_ms, "timed_region_s": ms * 1e-3,
                       "voxels_per_point": v_over_n, "key_space": int(res.desc.key_space),
                       "l2": f"inputs rotate over {POOL} distinct frames ({POOL * n * 16 / 1e6:.0f} MB > 126 MB L2)",
                       "streams": S, "backend": args.mode,
                       "launch": ("ordinary launch + programmatic dependent launch (one pipeline, one stream, frames "
                                  "resident before they are enqueued)"
                                  if streaming else "cooperative launch" if fused else "five ordinary launches"),
                       "l2_persist_mb": args.l2_persist_mb,
                       "sharding": "independent frames per rank, no collective"},
            "roofline": roofline, "roofline_step": roofline_step, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": (1 if fused else 5) * args.steps * F, "clocks": clk, "extra": extra,
        }
        emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
