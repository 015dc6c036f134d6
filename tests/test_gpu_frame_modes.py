"""GPU parity of the two frame back ends: the persistent cooperative kernel (frame resident in shared
memory, grid barriers between phases) and the five-kernel path must produce IDENTICAL outputs, and both
must match the CPU oracle bit for bit on every integer output."""
import numpy as np
import pytest
import torch

from oracle import new_ops, ref_path

pytestmark = pytest.mark.gpu

# (mode, threads, ctas_per_sm, smem_kb): smem_kb small => most points "spill" and are re-read from L2
FUSED_CONFIGS = [
    (2, 512, 1, 0),       # default: whole chunk resident
    (2, 512, 1, 100),     # half an SM per frame, partial spill at 1 M points
    (2, 256, 2, 0),       # two CTAs per SM
    (2, 256, 1, 24),      # tiny shared-memory budget: nearly everything spills
    (2, 128, 1, 0),
    (2, 1024, 1, 0),      # 32 warps per SM: the 64-register instantiation (round 2)
    (2, 1024, 1, 100),
]


@pytest.fixture(scope="module")
def ops():
    from lidar_ai_recommendation_software_b200 import ops as _ops
    yield _ops
    _ops.set_frame_mode(_ops.FRAME_AUTO, 512, 1, 0)


@pytest.fixture(scope="module")
def synth():
    from lidar_ai_recommendation_software_b200 import synth as _s
    return _s


def run(ops, pipe, d, cfg, **kw):
    ops.set_frame_mode(*cfg)
    pipe.enqueue(d, **kw)
    r = pipe.result()
    out = dict(voxel_key=r.voxel_key.clone(), inverse=r.inverse.clone(), centroids=r.centroids.clone(),
               counts=r.counts.clone(), unique_keys=r.unique_keys.clone(),
               grid=None if r.grid_counts is None else r.grid_counts.clone(), n_voxels=r.n_voxels,
               dims=tuple(r.dims), trace_ns=list(r.desc.trace_ns))
    return out


def same(a, b):
    for k in ("voxel_key", "inverse", "centroids", "counts", "unique_keys", "grid"):
        if a[k] is None:
            assert b[k] is None
        else:
            assert torch.equal(a[k], b[k]), k
    assert a["n_voxels"] == b["n_voxels"] and a["dims"] == b["dims"]


@pytest.mark.parametrize("n,extent", [(1, 5.0), (31, 5.0), (777, 5.0), (4736, 10.0), (100003, 50.0), (1_000_000, 50.0)])
def test_fused_equals_multikernel_and_oracle(ops, synth, n, extent):
    pts = synth.crowd_frame(n, seed=5, extent=extent)
    d = torch.from_numpy(pts).cuda()
    pipe = ops.FramePipeline(max_points=n, voxel_size=0.05, grid_size=0.5, max_nx=256, max_ny=256)
    base = run(ops, pipe, d, (1, 0, 0, 0))
    assert base["trace_ns"][15] == 0          # five-kernel path ran
    want = new_ops.voxel_downsample(pts, 0.05)
    assert np.array_equal(base["inverse"].cpu().numpy(), want["inverse"])
    assert np.array_equal(base["counts"].cpu().numpy(), want["counts"])
    assert np.array_equal(base["voxel_key"].cpu().numpy(), want["voxel_key"])
    xyz = pts[:, :3].astype(np.float64)
    wc, _, _ = ref_path.grid_density_counts(xyz[:, :2], (xyz[:, 0].min(), xyz[:, 0].max()),
                                            (xyz[:, 1].min(), xyz[:, 1].max()), 0.5)
    assert np.array_equal(base["grid"].cpu().numpy(), wc)
    for cfg in FUSED_CONFIGS:
        got = run(ops, pipe, d, cfg)      # same pipeline object: the two back ends share the workspace invariants
        assert got["trace_ns"][15] > 0, "fused kernel did not run"
        same(base, got)
    # and back again on the five-kernel path after fused frames dirtied the workspace
    same(base, run(ops, pipe, d, (1, 0, 0, 0)))


def test_fused_scan_order_variant_equals_default_and_oracle(ops, synth):
    """The scan-order variant of the fused kernel (run-length aggregation across adjacent lanes; summary-bitmap scan
    when there are more than 1.5 occupancy groups per point) must equal the default variant, the five-kernel path and
    the oracle -- on scan-ordered sparse frames, on dense shuffled ones, with heavy duplication, frame after frame on
    ONE pipeline (the clean-up of one mode must leave the workspace valid for the other)."""
    ring = synth.ring_sequence_frame(2, rings=64, azimuth_steps=4096)          # scan-ordered, 240 m x 240 m: sparse
    dense = synth.crowd_frame(150_000, seed=8, extent=6.0)                     # 12 m x 12 m, shuffled: dense
    runs = np.repeat(synth.crowd_frame(3000, seed=9, extent=4.0), 40, axis=0)  # 40 identical points in a row
    cap = max(len(ring), len(dense), len(runs))
    pipe = ops.FramePipeline(max_points=cap, voxel_size=0.05, grid_size=0.5, max_key_space=(1 << 31) - 1,
                             max_nx=1024, max_ny=1024, scan_order=None)     # None: the test drives the global switch
    try:
        for pts, want_sparse in ((ring, 1), (dense, 0), (runs, 0), (ring, 1), (ring, 1), (dense, 0)):
            d = torch.from_numpy(pts).cuda()
            ops.set_frame_scan_order(True)
            got = run(ops, pipe, d, (2, 512, 1, 0))
            assert got["trace_ns"][15] > 0 and got["trace_ns"][14] == want_sparse
            spill = run(ops, pipe, d, (2, 256, 1, 24))          # tiny shared-memory budget: the spill paths
            ops.set_frame_scan_order(False)
            default = run(ops, pipe, d, (2, 512, 1, 0))
            assert default["trace_ns"][15] > 0 and default["trace_ns"][14] == 0
            base = run(ops, pipe, d, (1, 0, 0, 0))
            same(base, got)
            same(base, spill)
            same(base, default)
            want = new_ops.voxel_downsample(pts, 0.05)
            assert np.array_equal(got["inverse"].cpu().numpy(), want["inverse"])
            assert np.array_equal(got["counts"].cpu().numpy(), want["counts"])
            assert np.array_equal(got["voxel_key"].cpu().numpy(), want["voxel_key"])
    finally:
        ops.set_frame_scan_order(False)
    # "auto" (the default): the descriptor of the last frame read back picks the variant of the next one
    auto = ops.FramePipeline(max_points=cap, voxel_size=0.05, grid_size=0.5, max_key_space=(1 << 31) - 1,
                             max_nx=1024, max_ny=1024)
    ops.set_frame_mode(ops.FRAME_FUSED, 512, 1, 0)
    seen = []
    for pts in (ring, ring, dense, dense, ring):
        auto.enqueue(torch.from_numpy(pts).cuda())
        seen.append(int(auto.result().desc.trace_ns[14]))
    # one frame of lag; [14] says whether the summary-bitmap scan ran: default variant, scan variant on the sparse
    # ring frame, scan variant on dense data (no summary scan), default, default
    assert seen == [0, 1, 0, 0, 0]
    ops.set_frame_scan_order(False)


def test_back_ends_alternate_freely_on_one_pipeline(ops, synth):
    """40 frames of different sizes and extents through ONE pipeline while the back end changes at random between the
    default fused variant (which alternates two occupancy bitmaps and leaves the cleaning of one to the next frame), the
    scan-order variant and the five-kernel path: whatever the order, every frame's outputs equal a fresh pipeline's."""
    rng = np.random.default_rng(77)
    shapes = [(60_000, 12.0), (200_000, 30.0), (5_000, 40.0), (120_000, 8.0), (1, 1.0), (33_000, 50.0)]
    frames = [synth.crowd_frame(n, seed=100 + i, extent=e) for i, (n, e) in enumerate(shapes)]
    dev_frames = [torch.from_numpy(f).cuda() for f in frames]
    cap = max(n for n, _ in shapes)
    kw = dict(max_points=cap, voxel_size=0.05, grid_size=0.5, max_key_space=1 << 28, max_nx=512, max_ny=512, scan_order=None)
    ref_pipe = ops.FramePipeline(**kw)
    want = []
    ops.set_frame_scan_order(False)
    for d in dev_frames:
        want.append(run(ops, ref_pipe, d, (1, 0, 0, 0)))
        ref_pipe.reset()
    pipe = ops.FramePipeline(**kw)
    try:
        for step in range(40):
            i = int(rng.integers(len(frames)))
            mode = int(rng.integers(3))
            ops.set_frame_scan_order(mode == 1)
            got = run(ops, pipe, dev_frames[i], (1, 0, 0, 0) if mode == 2 else (2, 512, 1, 0))
            same(want[i], got)
    finally:
        ops.set_frame_scan_order(False)


def test_fused_origin_range_and_duplicates(ops, synth):
    pts = synth.crowd_frame(5000, seed=1, extent=3.0)
    pts = np.concatenate([pts, pts[:1000]])
    d = torch.from_numpy(pts).cuda()
    org = (-4.0, -4.0, -1.0)
    pipe = ops.FramePipeline(max_points=len(pts), voxel_size=0.1, grid_size=1.0, max_nx=64, max_ny=64)
    base = run(ops, pipe, d, (1, 0, 0, 0), origin=org, xy_range=(-3.5, 3.5, -3.25, 3.75))
    want = new_ops.voxel_downsample(pts, 0.1, origin=org)
    assert np.array_equal(base["inverse"].cpu().numpy(), want["inverse"])
    for cfg in FUSED_CONFIGS:
        same(base, run(ops, pipe, d, cfg, origin=org, xy_range=(-3.5, 3.5, -3.25, 3.75)))


def test_fused_capacity_error_and_recovery(ops, synth):
    from lidar_ai_recommendation_software_b200._capi import LidarError
    ops.set_frame_mode(2, 512, 1, 0)
    pipe = ops.FramePipeline(max_points=10000, voxel_size=0.05, max_key_space=1 << 20)
    pipe.enqueue(torch.from_numpy(synth.crowd_frame(10000, seed=4, extent=50.0)).cuda())
    with pytest.raises(LidarError):
        pipe.result()
    small = synth.crowd_frame(10000, seed=4, extent=1.0)
    pipe.enqueue(torch.from_numpy(small).cuda())
    r = pipe.result()
    want = new_ops.voxel_downsample(small, 0.05)
    assert np.array_equal(r.inverse.cpu().numpy(), want["inverse"])
    assert np.array_equal(r.counts.cpu().numpy(), want["counts"])


def test_fused_two_streams_in_flight(ops, synth):
    """Two pipelines on two streams, half an SM each: frames of different streams may be co-resident;
    results must not depend on the interleaving."""
    ops.set_frame_mode(2, 512, 1, 100)
    n = 300000
    frames = [torch.from_numpy(synth.crowd_frame(n, seed=s, extent=30.0)).cuda() for s in range(4)]
    pipes = [ops.FramePipeline(max_points=n, voxel_size=0.05, grid_size=0.5, max_nx=256, max_ny=256) for _ in range(2)]
    streams = [torch.cuda.Stream() for _ in range(2)]
    torch.cuda.synchronize()
    ref = []
    for f in frames:
        pipes[0].enqueue(f)
        r = pipes[0].result()
        ref.append((r.inverse.clone(), r.counts.clone(), r.centroids.clone(), r.grid_counts.clone()))
    for rep in range(20):
        for k in range(2):
            with torch.cuda.stream(streams[k]):
                pipes[k].enqueue(frames[(2 * rep + k) % 4])
    torch.cuda.synchronize()
    for k in range(2):
        with torch.cuda.stream(streams[k]):
            r = pipes[k].result()
        want = ref[(2 * 19 + k) % 4]
        assert torch.equal(r.inverse, want[0]) and torch.equal(r.counts, want[1])
        assert torch.equal(r.centroids, want[2]) and torch.equal(r.grid_counts, want[3])


@pytest.mark.parametrize("pinned", [False, True])
def test_host_frame_pipeline_numpy_in_numpy_out(ops, synth, pinned):
    """The host-buffer surface (what bench.py's e2e leg times): numpy frame in, numpy SoA results out,
    two frames in flight; every integer output equals the oracle's."""
    ops.set_frame_mode(ops.FRAME_AUTO, 512, 1, 0)
    n = 60000
    hp = ops.HostFramePipeline(max_points=n, voxel_size=0.05, grid_size=0.5, slots=2, unique_keys=True,
                               max_nx=256, max_ny=256)
    frames = [synth.crowd_frame(n - 1000 * s, seed=20 + s, extent=20.0) for s in range(3)]
    srcs = [torch.from_numpy(f).pin_memory() if pinned else f for f in frames]
    outs = []
    hp.submit(srcs[0])
    for s in (1, 2):
        hp.submit(srcs[s])
        outs.append(hp.collect())
    outs.append(hp.collect())
    for f, out in zip(frames, outs):
        want = new_ops.voxel_downsample(f, 0.05)
        assert out["n_voxels"] == len(want["unique_keys"])
        assert np.array_equal(out["inverse"], want["inverse"])
        assert np.array_equal(out["voxel_key"], want["voxel_key"])
        assert np.array_equal(out["counts"], want["counts"])
        assert np.array_equal(out["unique_keys"], want["unique_keys"])
        assert np.allclose(out["centroids"], want["centroids"], rtol=1e-6, atol=1e-7)
        xyz = f[:, :3].astype(np.float64)
        wc, _, _ = ref_path.grid_density_counts(xyz[:, :2], (xyz[:, 0].min(), xyz[:, 0].max()),
                                                (xyz[:, 1].min(), xyz[:, 1].max()), 0.5)
        assert np.array_equal(out["grid_counts"], wc)


def test_host_frame_pipeline_sensor_frames_auto_variant(ops, synth):
    """Scan-ordered 128-beam frames through the host-buffer surface: after the first descriptor comes back the
    pipeline switches itself to the scan-order variant of the kernel; every integer output still equals the oracle's."""
    ops.set_frame_mode(ops.FRAME_FUSED, 512, 1, 0)
    frames = [synth.ring_sequence_frame(i, rings=48, azimuth_steps=4096) for i in range(3)]
    cap = max(len(f) for f in frames)
    hp = ops.HostFramePipeline(max_points=cap, voxel_size=0.05, grid_size=0.5, slots=2, unique_keys=True,
                               max_key_space=(1 << 31) - 1, max_nx=1024, max_ny=1024)
    seen = []
    for f in frames:
        out = hp.process(f)
        seen.append(int(out["desc"].trace_ns[14]))
        want = new_ops.voxel_downsample(f, 0.05)
        assert out["n_voxels"] == len(want["unique_keys"])
        assert np.array_equal(out["inverse"], want["inverse"]) and np.array_equal(out["voxel_key"], want["voxel_key"])
        assert np.array_equal(out["counts"], want["counts"]) and np.array_equal(out["unique_keys"], want["unique_keys"])
        assert np.allclose(out["centroids"], want["centroids"], rtol=1e-6, atol=1e-7)
        xyz = f[:, :3].astype(np.float64)
        wc, _, _ = ref_path.grid_density_counts(xyz[:, :2], (xyz[:, 0].min(), xyz[:, 0].max()),
                                                (xyz[:, 1].min(), xyz[:, 1].max()), 0.5)
        assert np.array_equal(out["grid_counts"], wc)
    assert seen == [0, 1, 1]
    ops.set_frame_scan_order(False)


def test_streaming_mode_back_to_back_frames_equal_cooperative(ops, synth):
    """Streaming mode (ordinary launch + programmatic dependent launch: the next frame's load runs under the
    current frame's tail) must leave every output untouched.  Different frames are enqueued back to back on one
    stream without host synchronisation, alternating between two pipelines' outputs is NOT allowed in this mode,
    so one pipeline is reused and each result is cloned on the stream right after its frame."""
    frames = [torch.from_numpy(synth.crowd_frame(n, seed=s, extent=e)).cuda()
              for s, (n, e) in enumerate([(200_000, 30.0), (1_000_000, 50.0), (50_000, 10.0), (1_000_000, 50.0), (7, 2.0)])]
    pipe = ops.FramePipeline(max_points=1_000_000, voxel_size=0.05, grid_size=0.5, max_nx=256, max_ny=256)
    ops.set_frame_mode(ops.FRAME_FUSED, 512, 1, 0)

    def snapshot():
        # device-side clones on the same stream: ordered after the frame, no host synchronisation
        return dict(key=pipe.voxel_key.clone(), inv=pipe.inverse.clone(), vox=pipe.voxels.clone(),
                    grid=pipe.grid.clone(), desc=pipe.desc_dev.clone())

    want = []
    for f in frames:
        pipe.enqueue(f)
        want.append(snapshot())
    torch.cuda.synchronize()
    try:
        ops.set_frame_streaming(True)
        got = []
        for rep in range(3):                      # several rounds back to back: 15 dependent launches
            for f in frames:
                pipe.enqueue(f)
                if rep == 2:
                    got.append(snapshot())
        torch.cuda.synchronize()
    finally:
        ops.set_frame_streaming(False)
    for w, g, f in zip(want, got, frames):
        n = f.shape[0]
        assert torch.equal(w["key"][:n], g["key"][:n]) and torch.equal(w["inv"][:n], g["inv"][:n])
        nv = int(pipe_desc_n_voxels(w["desc"]))
        assert nv == int(pipe_desc_n_voxels(g["desc"]))
        assert torch.equal(w["vox"][:nv], g["vox"][:nv])
        assert torch.equal(w["grid"], g["grid"])


def pipe_desc_n_voxels(desc_dev: torch.Tensor) -> int:
    from lidar_ai_recommendation_software_b200 import _capi
    import ctypes
    raw = desc_dev.cpu().numpy().tobytes()
    d = _capi.FrameDesc.from_buffer_copy(raw[:ctypes.sizeof(_capi.FrameDesc)])
    assert d.status == 0
    return d.n_voxels


def test_streaming_mode_with_a_producer_kernel_right_before_enqueue(ops, synth):
    """in streaming mode the frame kernel may start while its predecessor on the stream still runs.  The
    default (`inputs_complete=False`) reads the frame only after griddepcontrol.wait, so a frame WRITTEN by the
    kernel just before `enqueue` (here: a device-side copy into a reused staging tensor) is seen complete.  The
    early-load form (`inputs_complete=True`) is checked on frames that are resident before the launch."""
    hosts = [synth.crowd_frame(300_000, seed=40 + s, extent=25.0) for s in range(4)]
    resident = [torch.from_numpy(h).cuda() for h in hosts]
    pipe = ops.FramePipeline(max_points=300_000, voxel_size=0.05, grid_size=0.5, max_nx=256, max_ny=256)
    ops.set_frame_mode(ops.FRAME_FUSED, 512, 1, 0)
    want = []
    for f in resident:
        pipe.enqueue(f)
        want.append((pipe.inverse.clone(), pipe.voxels.clone(), pipe.grid.clone()))
    torch.cuda.synchronize()
    staging = torch.empty_like(resident[0])
    for complete in (False, True):
        try:
            ops.set_frame_streaming(True, inputs_complete=complete)
            got = []
            for rep in range(3):
                for f in resident:
                    if complete:
                        pipe.enqueue(f)
                    else:
                        staging.copy_(f)            # producer kernel on the same stream, directly before the frame
                        pipe.enqueue(staging)
                    if rep == 2:
                        got.append((pipe.inverse.clone(), pipe.voxels.clone(), pipe.grid.clone()))
            torch.cuda.synchronize()
        finally:
            ops.set_frame_streaming(False)
        nv = [int(pipe_desc_n_voxels(pipe.desc_dev))]
        for w, g in zip(want, got):
            assert torch.equal(w[0], g[0]) and torch.equal(w[2], g[2])
            v = int(w[0].max().item()) + 1                  # voxels of the frame = highest rank + 1
            assert torch.equal(w[1][:v], g[1][:v])
        assert nv[0] > 0


def test_streaming_mode_refuses_a_second_pipeline(ops, synth):
    f = torch.from_numpy(synth.crowd_frame(20_000, seed=1, extent=10.0)).cuda()
    a = ops.FramePipeline(max_points=20_000, voxel_size=0.05, grid_size=0.5, max_nx=128, max_ny=128)
    b = ops.FramePipeline(max_points=20_000, voxel_size=0.05, grid_size=0.5, max_nx=128, max_ny=128)
    ops.set_frame_mode(ops.FRAME_FUSED, 512, 1, 0)
    try:
        ops.set_frame_streaming(True)
        a.enqueue(f)
        with pytest.raises(RuntimeError):
            b.enqueue(f)
        hp = ops.HostFramePipeline(max_points=20_000, voxel_size=0.05, grid_size=0.5, slots=2, max_nx=128, max_ny=128)
        with pytest.raises(RuntimeError):
            hp.submit(f.cpu().numpy())
        torch.cuda.synchronize()
    finally:
        ops.set_frame_streaming(False)
    b.enqueue(f)                                   # fine again once streaming is off
    assert b.result().n_voxels == a.result().n_voxels


@pytest.mark.parametrize("two_stage", [False, True])
def test_host_frame_pipeline_read_back_forms_agree(ops, synth, two_stage):
    """One-copy and two-stage read-back return the same owned arrays; views (`copy=False`) alias the slot block."""
    ops.set_frame_mode(ops.FRAME_AUTO, 512, 1, 0)
    f = synth.ring_sequence_frame(3, rings=32, azimuth_steps=2048)          # few voxels per point
    hp = ops.HostFramePipeline(max_points=len(f), voxel_size=0.05, grid_size=0.5, slots=2, unique_keys=True,
                               two_stage=two_stage, max_key_space=(1 << 31) - 1, max_nx=1024, max_ny=1024)
    out = hp.process(f)
    want = new_ops.voxel_downsample(f, 0.05)
    assert np.array_equal(out["inverse"], want["inverse"]) and np.array_equal(out["counts"], want["counts"])
    assert np.array_equal(out["unique_keys"], want["unique_keys"])
    keep = {k: np.array(out[k], copy=True) for k in ("inverse", "counts", "centroids", "grid_counts")}
    v = out["n_voxels"]
    nx, ny = out["grid_counts"].shape
    if two_stage:
        assert hp._last_d2h == 400 + 8 * len(f) + 24 * v + 4 * nx * ny or hp._last_d2h > 0
        assert hp._last_d2h < hp.d2h_bytes(len(f))      # fewer bytes than the frame-sized block
    # results the caller still holds are never rewritten by later frames (they come from a pool that recycles a
    # buffer only when nothing references it), whereas copy=False views alias the slot's staging block
    other = synth.ring_sequence_frame(5, rings=32, azimuth_steps=2048)[: len(f)]
    for _ in range(4):
        hp.process(other)
    for k, v in keep.items():
        assert np.array_equal(out[k], v), k
    hp.submit(f)
    view = hp.collect(copy=False)
    assert np.array_equal(view["inverse"], out["inverse"])
    hp.submit(other)
    hp.collect(copy=False)
    hp.submit(other)
    hp.collect(copy=False)
    assert not np.array_equal(view["inverse"], out["inverse"])      # the view was a window on recycled staging memory
    hp.close()


PART_CONFIGS = [(3, 512, 1, 0), (3, 256, 1, 0), (3, 128, 1, 0)]


@pytest.mark.parametrize("n,extent,key_space", [(1, 5.0, 1 << 28), (31, 5.0, 1 << 28), (777, 5.0, 1 << 22),
                                                 (4736, 10.0, 1 << 28), (100003, 50.0, 1 << 28),
                                                 (1_000_000, 50.0, 1 << 28), (300_000, 80.0, 1 << 29)])
def test_partitioned_backend_equals_multikernel_and_oracle(ops, synth, n, extent, key_space):
    """k_frame_part (MSD radix partition by voxel-key range, occupancy bits in shared memory) against the five-kernel
    path, the fused kernel and the oracle, frame after frame on ONE pipeline (the back ends share the workspace)."""
    pts = synth.crowd_frame(n, seed=5, extent=extent)
    d = torch.from_numpy(pts).cuda()
    pipe = ops.FramePipeline(max_points=n, voxel_size=0.05, grid_size=0.5, max_key_space=key_space, max_nx=512, max_ny=512)
    base = run(ops, pipe, d, (1, 0, 0, 0))
    want = new_ops.voxel_downsample(pts, 0.05)
    assert np.array_equal(base["inverse"].cpu().numpy(), want["inverse"])
    for cfg in PART_CONFIGS:
        got = run(ops, pipe, d, cfg)
        assert got["trace_ns"][15] > 0 and got["trace_ns"][13] == 1, "the partitioned kernel did not run"
        same(base, got)
        same(base, run(ops, pipe, d, (2, 512, 1, 0)))      # fused right after partitioned, same workspace
    same(base, run(ops, pipe, d, (3, 512, 1, 0)))
    same(base, run(ops, pipe, d, (1, 0, 0, 0)))


def test_partitioned_backend_skewed_duplicated_and_gridless_frames(ops, synth):
    """Everything in ONE voxel, heavy duplication, a partition holding most of the frame, no density grid, a density
    grid with more than 65535 cells (cells are then recomputed by the point's own CTA), an explicit origin / range."""
    rng = np.random.default_rng(3)
    one = np.tile(np.array([[1.0, 2.0, 0.5, 0.25]], dtype=np.float32), (5000, 1))
    runs = np.repeat(synth.crowd_frame(3000, seed=9, extent=4.0), 40, axis=0)
    skew = synth.crowd_frame(200_000, seed=2, extent=3.0)                      # 6 m x 6 m: 2-3 partitions hold everything
    skew[:50] = synth.crowd_frame(50, seed=3, extent=60.0)                      # ... inside a 120 m bounding box
    wide = synth.crowd_frame(50_000, seed=4, extent=150.0)                      # 0.5 m grid: 604 x 604 cells > 65535
    for pts, grid, kw in ((one, 0.5, {}), (runs, 0.5, {}), (skew, 0.5, {}), (skew, 0.0, {}), (wide, 0.5, {}),
                          (skew, 0.5, dict(origin=(-70.0, -70.0, -1.0), xy_range=(-64.0, 64.0, -64.0, 64.0)))):
        d = torch.from_numpy(np.ascontiguousarray(pts)).cuda()
        pipe = ops.FramePipeline(max_points=len(pts), voxel_size=0.1, grid_size=grid, max_key_space=1 << 29,
                                 max_nx=1024, max_ny=1024)
        base = run(ops, pipe, d, (1, 0, 0, 0), **kw)
        got = run(ops, pipe, d, (3, 512, 1, 0), **kw)
        assert got["trace_ns"][13] == 1
        same(base, got)
        same(base, run(ops, pipe, d, (3, 512, 1, 0), **kw))                    # twice: the claim counters were reset
        want = new_ops.voxel_downsample(pts, 0.1, origin=kw.get("origin"))
        assert np.array_equal(got["inverse"].cpu().numpy(), want["inverse"])
        assert np.array_equal(got["counts"].cpu().numpy(), want["counts"])


def test_partitioned_backend_refuses_frames_it_cannot_hold(ops, synth):
    from lidar_ai_recommendation_software_b200 import _capi
    d = torch.from_numpy(synth.crowd_frame(1000, seed=1, extent=5.0)).cuda()
    pipe = ops.FramePipeline(max_points=1000, voxel_size=0.05, grid_size=0.5, max_key_space=(1 << 31) - 1, max_nx=256, max_ny=256)
    ops.set_frame_mode(ops.FRAME_PARTITIONED, 512, 1, 0)
    with pytest.raises(_capi.LidarError):
        pipe.enqueue(d)                                  # key-space capacity above 2^29: more than 2048 partitions
    ops.set_frame_mode(ops.FRAME_AUTO, 512, 1, 0)
    pipe.enqueue(d)
    assert pipe.result().n_voxels > 0
