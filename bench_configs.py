#!/usr/bin/env python
"""Secondary measurements: BASELINE.json configs[2], [3], [4].

bench.py is the contract line (configs[1]).  This script times the other named shapes on B200 and
prints one JSON line per config; results are kept under profiles/.

    python bench_configs.py --config surfaces  # cfg 1 shapes through the drop-in call surfaces A and B
    python bench_configs.py --config sa      # cfg 3: FPS 16384->1024, ball query r=0.2 k=32, MLP 64-64-128, batch 16
    python bench_configs.py --config seq     # cfg 4: 128-beam sequence (~2.6 M pts/frame), frames shard over ranks
    python bench_configs.py --config scan    # cfg 5: 50 M-point venue scan, points shard over ranks, allreduce of the grid
    torchrun --nproc-per-node N bench_configs.py --config seq|scan     (N = 2, 4, 8)

All timing is CUDA events on the launching stream, after warm-up; multi-rank numbers are the max over
ranks.  Nothing here reads /root/reference or executes oracle/ (parity lives in tests/).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
# The contract is ONE JSON line on stdout, but libraries write there too (torch's NCCL process group announces
# "NCCL version ..." on fd 1 at the first collective).  Everything that is not the result line goes to stderr.
_RESULT_FD = None


def _redirect_stdout() -> None:
    global _RESULT_FD
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line: str) -> None:
    os.write(_RESULT_FD if _RESULT_FD is not None else 1, (line + "\n").encode())


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), float(d["bf16_tflops"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1600.0, "fallback (B200_PROFILING.md)"


def ev_time(fn, reps, warm, torch):
    """median device time of fn() in ms"""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


def run_surfaces(args, torch, dev):
    """cfg 1 shapes through the drop-in call surfaces (numpy in -> numpy / dict out, copies included): the calls
    the Streamlit apps make.  The reference-side CPU times of the same calls are in SURVEY.md §6 (measured with the
    unmodified reference in the build container; /root/reference does not exist on the GPU box)."""
    from lidar_ai_recommendation_software_b200 import apps, synth
    from lidar_ai_recommendation_software_b200.models.crowd_density_model import CrowdDensityModel
    from lidar_ai_recommendation_software_b200.models.crowd_flow_model import CrowdFlowModel
    from lidar_ai_recommendation_software_b200.utils import data_processing as dp

    def wall(fn, reps=5, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        return float(np.min(ts)) * 1e3

    out = {}
    for n in (100_000, 1_000_000):
        f = synth.crowd_frame(n, seed=0, extent=50.0)
        pts = np.ascontiguousarray(f[:, :3], dtype=np.float64)
        tag = f"{n // 1000}k"
        out[f"preprocess_lidar_data_{tag}_ms"] = wall(lambda: dp.preprocess_lidar_data(pts))
        out[f"preprocess_point_cloud_{tag}_ms"] = wall(lambda: apps.preprocess_point_cloud(pts))
        pa = dp.preprocess_lidar_data(pts)
        pb = apps.preprocess_point_cloud(pts)
        out[f"clusters_A_{tag}"] = int(pa["clusters"].max() + 1)
        out[f"clusters_B_{tag}"] = int(pb["clusters"].max() + 1)
        out[f"extract_people_positions_B_{tag}_ms"] = wall(lambda: dp.extract_people_positions(pb))
        out[f"CrowdDensityModel.analyze_{tag}_ms"] = wall(lambda: CrowdDensityModel(grid_size=1.0).analyze(pb))
        out[f"CrowdFlowModel.analyze_{tag}_ms"] = wall(lambda: CrowdFlowModel().analyze(pb))
        out[f"analyze_crowd_density_B_{tag}_ms"] = wall(lambda: apps.analyze_crowd_density(pb))
        out[f"analyze_crowd_flow_B_{tag}_ms"] = wall(lambda: apps.analyze_crowd_flow(pb))
        out[f"density_heatmap_counts_{tag}_ms"] = wall(lambda: apps.density_heatmap_counts(pb))
        xr = (pts[:, 0].min(), pts[:, 0].max())
        yr = (pts[:, 1].min(), pts[:, 1].max())
        out[f"calculate_grid_density_points_0.5m_{tag}_ms"] = wall(lambda: dp.calculate_grid_density(pts[:, :2], xr, yr, 0.5))
        out[f"downsample_point_cloud_0.1_{tag}_ms"] = wall(lambda: dp.downsample_point_cloud(pts, 0.1))
    line = {"config": "configs[0] shapes (and 1 M) through the drop-in surfaces A and B, numpy in -> numpy out, host wall clock, best of 5",
            "n_gpus": 1, "data": "synthetic (Appendix C.1)", "ms": out,
            "reference_cpu_ms_SURVEY_6": {"preprocess_lidar_data_100k": 4770.0, "preprocess_point_cloud_100k": 1060.0,
                                          "CrowdDensityModel.analyze": 3.9, "analyze_crowd_density_B_100x100": 1010.0,
                                          "CrowdFlowModel.analyze": "130-290", "calculate_grid_density_points_100k": 16.1,
                                          "calculate_grid_density_points_1000k": 166.0, "histogram2d_bins100_1000k": 152.0,
                                          "downsample_point_cloud_1000k": 38.0,
                                          "preprocess_lidar_data_1000k": "infeasible (sklearn neighbour lists ~50 GB)"}}
    emit(json.dumps(line))


def run_sa(args, torch, dev, rank, world):
    from lidar_ai_recommendation_software_b200 import pointnet2 as pn, synth
    B, N, M, K, R = 16, 16384, 1024, 32, 0.2
    xyz = torch.from_numpy(synth.sa_batch(B, N, seed=rank)).to(dev)
    ws, bs = synth.sa_weights(seed=1)
    ws = [torch.from_numpy(w).to(dev) for w in ws]
    bs = [torch.from_numpy(b).to(dev) for b in bs]
    hbm, tf, src = peaks()
    fps_idx = pn.furthest_point_sample(xyz, M)
    new_xyz = pn.gather_points(xyz, fps_idx)
    idx = pn.ball_query(xyz, new_xyz, R, K)
    out = {}
    t_fps, _ = ev_time(lambda: pn.furthest_point_sample(xyz, M), args.reps, 3, torch)
    t_gat, _ = ev_time(lambda: pn.gather_points(xyz, fps_idx), args.reps, 3, torch)
    t_bq, _ = ev_time(lambda: pn.ball_query(xyz, new_xyz, R, K), args.reps, 3, torch)
    t_grp, _ = ev_time(lambda: pn.group_points(xyz, None, idx, new_xyz), args.reps, 3, torch)
    t_tc, _ = ev_time(lambda: pn.shared_mlp_maxpool(ws, bs, xyz=xyz, idx=idx, new_xyz=new_xyz, impl=pn.MLP_TCGEN05),
                      args.reps, 3, torch)
    t_simt, _ = ev_time(lambda: pn.shared_mlp_maxpool(ws, bs, xyz=xyz, idx=idx, new_xyz=new_xyz, impl=pn.MLP_SIMT),
                        args.reps, 3, torch)
    sa = pn.SetAbstraction(M, R, K, ws, bs)
    t_all, t_all_min = ev_time(lambda: sa(xyz), args.reps, 3, torch)
    flops = 2.0 * B * M * K * (3 * 64 + 64 * 64 + 64 * 128)
    # the tcgen05 path issues 3 bf16 MMAs per product (hi*hi + hi*lo + lo*hi) for layers 2 and 3
    mma_flops = 2.0 * B * M * K * 3 * (64 * 64 + 64 * 128)
    tests = float(B) * M * N
    line = {
        "config": "configs[2]: PointNet++ SSG set abstraction, FPS 16384->1024, ball query r=0.2 k=32, MLP 64-64-128, batch 16",
        "n_gpus": 1, "data": "synthetic (Appendix C.2)", "reps": args.reps,
        "ms": {"fps": t_fps, "gather_centres": t_gat, "ball_query": t_bq, "group_points_materialised": t_grp,
               "shared_mlp_tcgen05_fused_group": t_tc, "shared_mlp_simt_fused_group": t_simt,
               "set_abstraction_total": t_all},
        "fps_us_per_iteration": t_fps * 1e3 / M,
        "ball_query_Gtests_per_s": tests / (t_bq * 1e-3) / 1e9,
        "mlp": {"algorithmic_GFLOP": flops / 1e9, "tcgen05_TFLOPs_algorithmic": flops / (t_tc * 1e-3) / 1e12,
                "tcgen05_TFLOPs_issued_bf16": mma_flops / (t_tc * 1e-3) / 1e12,
                "simt_TFLOPs": flops / (t_simt * 1e-3) / 1e12,
                "roofline": {"bound": "tensor", "achieved": mma_flops / (t_tc * 1e-3) / 1e12, "peak": tf,
                             "unit": "TFLOP/s", "frac": mma_flops / (t_tc * 1e-3) / 1e12 / tf, "peak_source": src,
                             "note": "issued bf16 MMA flops (3 per fp32-accurate product); the kernel also runs layer 1, "
                                     "the bf16 hi/lo split, the TMEM epilogue and the max-pool"}},
        "Mpoints_per_s_input": B * N / (t_all * 1e-3) / 1e6,
        "out_bytes": B * 128 * M * 4,
    }
    return line


def run_seq(args, torch, dev, rank, world, dist):
    """cfg 4: per frame = variant-B preprocess (3 sigma, ground split, DBSCAN eps 0.3, labels) + people
    centroids + flow (first frame: simulated field as the reference; later: measured displacement)
    + the frame pipeline (voxel downsample + density) on the same points."""
    from lidar_ai_recommendation_software_b200 import apps, ops, preprocess, sharding, synth
    from lidar_ai_recommendation_software_b200.models.crowd_flow_model import CrowdFlowModel
    n_frames = args.frames
    mine, halo = sharding.frame_range_with_halo(n_frames, rank, world)
    pool_n = min(args.pool, len(mine) + (1 if halo is not None else 0))
    t0 = time.perf_counter()
    first = halo if halo is not None else (mine[0] if len(mine) else 0)
    pool = [synth.ring_sequence_frame(first + i, rings=args.rings, azimuth_steps=args.azimuth) for i in range(pool_n)]
    gen_s = time.perf_counter() - t0
    pts_per_frame = float(np.mean([p.shape[0] for p in pool]))
    pool64 = [torch.from_numpy(np.ascontiguousarray(p[:, :3], dtype=np.float64)).pin_memory() for p in pool]
    dpool = [torch.from_numpy(p).to(dev) for p in pool]
    nmax = max(p.shape[0] for p in pool)
    if getattr(args, "seq_frame_mode", "fused") == "multikernel":
        ops.set_frame_mode(ops.FRAME_MULTIKERNEL)
    pipe = ops.FramePipeline(max_points=nmax, voxel_size=0.05, grid_size=0.5, max_key_space=(1 << 31) - 1,
                             max_nx=1024, max_ny=1024, device=dev)
    model = CrowdFlowModel()

    def one(i):
        k = i % pool_n
        pipe.enqueue(dpool[k])
        if args.dropin:
            pd = apps.preprocess_point_cloud(pool64[k].numpy())          # numpy in, numpy dict out
        else:
            d64 = pool64[k].to(dev, non_blocking=True)                   # pinned host -> device
            pd = preprocess.run(d64, variant="B", host_arrays=False)      # per-point outputs stay on the device
        res = model.analyze_sequence_frame(pd, dt=0.1)
        return pd, res

    # warm-up (also the halo frame of this shard: recomputed locally)
    pd, res = one(0)
    fr = pipe.result()
    torch.cuda.synchronize()
    people = 0
    matched = 0
    if args.dropin or args.workers <= 1:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for j, f in enumerate(mine):
            pd, res = one(j + 1)
            people += pd[preprocess.DEVICE_KEY].n_clusters
            if "matches" in res:
                matched += int((np.asarray(res["matches"]) >= 0).sum())
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    else:
        from lidar_ai_recommendation_software_b200.sequence import SequenceRunner
        runner = SequenceRunner(variant="B", workers=args.workers, dt=0.1)
        runner.model = model                      # carries the halo frame's people positions
        # warm the worker threads: every (stream, frame size) pair once, so that the caching allocator of each
        # stream already owns its blocks (a cudaMalloc inside the timed region costs a device synchronisation)
        for _ in runner.run([pool64[i % pool_n] for i in range(pool_n * args.workers)]):
            pass
        model.prev_positions = None
        list(runner.run([pool64[0]]))             # halo / first frame again, so the timed frames all match
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()

        def feed():
            for j in range(len(mine)):
                k = (j + 1) % pool_n
                pipe.enqueue(dpool[k])            # voxel downsample + density of the float4 frame (main stream)
                yield pool64[k]

        for pd, res in runner.run(feed()):
            people += pd[preprocess.DEVICE_KEY].n_clusters
            if "matches" in res:
                matched += int((np.asarray(res["matches"]) >= 0).sum())
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        runner.close()
    # device-only share: the frame pipeline alone on the same frames
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for j in range(len(mine)):
        pipe.enqueue(dpool[j % pool_n])
    e1.record()
    torch.cuda.synchronize()
    vox_ms = e0.elapsed_time(e1) / max(1, len(mine))
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    if rank == 0:
        line = {
            "config": f"configs[3]: crowd_flow over a {n_frames}-frame {args.rings}-beam sequence, frames sharded over {world} GPU(s)",
            "n_gpus": world, "frames": n_frames, "frames_per_rank": len(mine), "halo_frames_recomputed": 0 if halo is None else 1,
            "points_per_frame": pts_per_frame, "distinct_frames_cycled_per_rank": pool_n,
            "wall_s": dt, "frames_per_s": n_frames / dt, "Mpoints_per_s": n_frames * pts_per_frame / dt / 1e6,
            "per_frame": ("apps.preprocess_point_cloud (pageable numpy f64 in, numpy dict out" if args.dropin else
                          "preprocess.run(host_arrays=False) (pinned f64 in, per-point outputs stay on the device") +
                         ": 3 sigma, 30th-pct ground split, DBSCAN eps 0.3, labels) + extract_people_positions + "
                         "CrowdFlowModel.analyze_sequence_frame (NEW frame_flow after the first frame) + FramePipeline "
                         "voxel 0.05 m / density 0.5 m on the float4 frame",
            "preprocess_workers": 1 if (args.dropin or args.workers <= 1) else args.workers,
            "mean_clusters_per_frame": people / max(1, len(mine)),
            "frame_pipeline_ms_per_frame_device": vox_ms,
            "frame_pipeline_Mpoints_per_s_device": pts_per_frame / (vox_ms * 1e-3) / 1e6,
            "voxels_per_point": fr.n_voxels / pool[0].shape[0],
            "clusters_last_frame": pd[preprocess.DEVICE_KEY].n_clusters, "matched_total_rank0": matched,
            "timing": "host wall clock around the per-frame drop-in calls (numpy in / numpy out, copies included), max over ranks",
            "synth_s_per_frame_host": gen_s / max(1, pool_n), "data": "synthetic (Appendix C.3)", "scaling": "strong",
        }
        return line
    return None


def run_scan(args, torch, dev, rank, world, dist):
    """cfg 5: one merged scan sharded by points.  The product call is ONE enqueue per rank
    (sharding.ScanDensity, backend "fused": bbox -> bbox exchange -> device-side edges -> histogram -> two-shot
    NVLink all-reduce of the grid -> density), then one exactly-sized read-back.  Alongside: the same call through
    the lidar_nccl_* fallback, and the fused kernel on the same shard WITHOUT the peers (the difference is what the
    exchanges cost)."""
    import ctypes as C
    from lidar_ai_recommendation_software_b200 import _capi, ops, sharding, synth
    n_total = args.points
    parts = max(world, args.host_shards)
    per_rank = parts // world
    chunks = []
    t0 = time.perf_counter()
    for s in range(rank * per_rank, (rank + 1) * per_rank):
        chunks.append(torch.from_numpy(synth.venue_scan_shard(n_total, 7, s, parts)).to(dev))
    shard = torch.cat(chunks)
    del chunks
    gen_s = time.perf_counter() - t0
    n_local = shard.shape[0]
    hbm, _, src = peaks()
    g = 0.5

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def call_times(ctx):
        """median host wall clock of the whole call (enqueue -> owned numpy arrays) and median device time of the
        enqueue alone (CUDA events), max over ranks"""
        ctx(shard, g)
        ctx(shard, g)
        ts, ks = [], []
        for _ in range(args.reps):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            ctx.enqueue(shard, g)
            e1.record()
            out = ctx.result()
            ts.append(time.perf_counter() - t0)
            torch.cuda.synchronize()
            ks.append(e0.elapsed_time(e1))
        t = torch.tensor([float(np.median(ts)), float(np.median(ks))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0].item()), float(t[1].item()), out

    fused = sharding.ScanDensity(dev, backend="auto" if world > 1 else "fused")
    call_s, kern_ms, (gx, gy, dens) = call_times(fused)
    # the same call when only rank 0 wants the arrays on its host (every rank's device copy is complete either way):
    # seven of eight read-backs over the shared PCIe complex disappear
    root_s = None
    if world > 1:
        ts = []
        for _ in range(args.reps):
            barrier()
            t0 = time.perf_counter()
            fused.enqueue(shard, g)
            fused.result(fetch=(rank == 0))
            ts.append(time.perf_counter() - t0)
            torch.cuda.synchronize()
        t = torch.tensor([float(np.median(ts))], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        root_s = float(t.item())
    total_counts = float(dens.sum() * g * g)
    backend = fused.backend
    multicast = bool(getattr(fused, "multicast", False))
    nccl_call_s = nccl_kern_ms = ar_ms = None
    local_ms = kern_ms
    if world > 1:
        if backend == "fused":
            nccl = sharding.ScanDensity(dev, backend="nccl")
            nccl_call_s, nccl_kern_ms, (_, _, d2) = call_times(nccl)
            assert np.array_equal(d2, dens), "fused and NCCL grids differ"
            grid_buf = torch.zeros(len(gx) * len(gy), dtype=torch.int32, device=dev)
            st = torch.cuda.current_stream().cuda_stream
            ar_ms, _ = ev_time(lambda: _capi.check(_capi.lib.lidar_nccl_allreduce(nccl.nccl, grid_buf.data_ptr(), grid_buf.numel(),
                                                                                   _capi.NCCL_SUM_I32, st)), args.reps, 3, torch)
            nccl.close()
        # the same shard through the single-rank kernel: what the rank does without its peers
        solo = sharding.ScanDensity(dev, backend="fused", solo=True)
        for _ in range(3):
            solo(shard, g)
        ks = []
        for _ in range(args.reps):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            solo.enqueue(shard, g)
            e1.record()
            solo.result()
            ks.append(e0.elapsed_time(e1))
        t = torch.tensor([float(np.median(ks))], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        local_ms = float(t.item())
        solo.close()
    fused.close()
    if rank == 0:
        return {
            "config": f"configs[4]: {n_total}-point merged venue scan sharded by points over {world} GPU(s), all-reduce of the int32 grid",
            "n_gpus": world, "points": n_total, "points_per_rank": n_local, "grid": [len(gx), len(gy)],
            "counts_sum": total_counts, "backend": backend, "nvls_multicast": multicast,
            "call_ms": call_s * 1e3, "Mpoints_per_s_call": n_total / call_s / 1e6,
            "call_ms_result_on_rank0_only": None if root_s is None else root_s * 1e3,
            "kernel_ms": kern_ms, "kernel_ms_without_peers": local_ms,
            "collectives_us": (kern_ms - local_ms) * 1e3 if world > 1 else 0.0,
            "Mpoints_per_s_kernel": n_total / (kern_ms * 1e-3) / 1e6,
            "nccl_fallback": None if nccl_call_s is None else {"call_ms": nccl_call_s * 1e3, "enqueue_ms": nccl_kern_ms,
                                                                "allreduce_grid_us": ar_ms * 1e3},
            "roofline": {"bound": "hbm", "kernel": "k_scan_density", "achieved": 32.0 * n_local / (local_ms * 1e-3) / 1e9,
                         "peak": hbm, "unit": "GB/s", "frac": 32.0 * n_local / (local_ms * 1e-3) / 1e9 / hbm, "peak_source": src,
                         "algorithmic_bytes_per_point": 32.0,
                         "note": "two passes over the shard (bbox, then histogram: the edges depend on the global bbox), 16 B/point each"},
            "collective_payload_bytes": 4 * len(gx) * len(gy) + 32,
            "timing": "call = host wall clock from enqueue to owned numpy arrays (median), kernel = CUDA events around the one enqueue; max over ranks",
            "synth_s_host": gen_s, "data": "synthetic (Appendix C.4)", "scaling": "strong",
        }
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True, choices=["sa", "seq", "scan", "surfaces"])
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--frames", type=int, default=300)
    ap.add_argument("--pool", type=int, default=4, help="distinct synthetic frames generated per rank (cycled)")
    ap.add_argument("--rings", type=int, default=128)
    ap.add_argument("--azimuth", type=int, default=20480)
    ap.add_argument("--dropin", action="store_true", help="seq: time the numpy-in / numpy-out drop-in surface instead")
    ap.add_argument("--seq-frame-mode", default="fused", choices=["fused", "multikernel"],
                    help="seq: back end of the voxel / density pipeline that shares the device with the preprocess workers")
    ap.add_argument("--workers", type=int, default=2,
                    help="seq: preprocess worker threads (one CUDA stream each); 1 = the serial loop")
    ap.add_argument("--points", type=int, default=50_000_000)
    ap.add_argument("--host-shards", type=int, default=8)
    args = ap.parse_args()
    _redirect_stdout()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench_configs.py needs a CUDA device; there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.config == "surfaces":
        if rank == 0:
            run_surfaces(args, torch, dev)
    elif args.config == "sa":
        if rank == 0:
            emit(json.dumps(run_sa(args, torch, dev, rank, world)))
    else:
        line = (run_seq if args.config == "seq" else run_scan)(args, torch, dev, rank, world, dist)
        if line is not None:
            emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
