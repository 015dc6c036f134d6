import sys, time, ctypes as C, numpy as np, torch
sys.path.insert(0, '.')
from lidar_ai_recommendation_software_b200 import ops, synth, _capi, sharding
lib = _capi.lib
torch.cuda.set_device(0)
def t(fn, reps=20):
    fn(); fn()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    return (time.perf_counter() - t0) / reps
nb = 28 << 20
pin = ops._PinnedBlock(nb)
page = np.empty(nb, dtype=np.uint8); page[:] = 1
for th in (1, 2, 4, 8, 12):
    lib.lidar_host_copy_threads(th)
    a = t(lambda: lib.lidar_host_memcpy(page.ctypes.data, pin.ptr, nb))
    b = t(lambda: lib.lidar_host_memcpy(pin.ptr, page.ctypes.data, nb))
    print(f"threads {th}: pinned->pageable {nb/a/1e9:.1f} GB/s, pageable->pinned {nb/b/1e9:.1f} GB/s", flush=True)
a = t(lambda: np.copyto(page, pin.u8)); print(f"np.copyto pinned->pageable {nb/a/1e9:.1f} GB/s")
def fresh():
    x = np.empty(nb, dtype=np.uint8); lib.lidar_host_memcpy(x.ctypes.data, pin.ptr, nb)
lib.lidar_host_copy_threads(8)
a = t(fresh); print(f"fresh np.empty + parallel memcpy: {a*1e3:.2f} ms ({nb/a/1e9:.1f} GB/s)")
def fresh1():
    x = np.empty(nb, dtype=np.uint8); np.copyto(x, pin.u8)
a = t(fresh1); print(f"fresh np.empty + np.copyto: {a*1e3:.2f} ms")
# DMA rates
d = torch.empty(nb, dtype=torch.uint8, device='cuda')
st = torch.cuda.current_stream().cuda_stream
def h2d(): lib.lidar_copy_async(d.data_ptr(), pin.ptr, nb, 1, st); torch.cuda.synchronize()
def d2h(): lib.lidar_copy_async(pin.ptr, d.data_ptr(), nb, 0, st); torch.cuda.synchronize()
print(f"H2D {nb/t(h2d)/1e9:.1f} GB/s  D2H {nb/t(d2h)/1e9:.1f} GB/s")
# one frame through the host pipeline, step by step
n = 1_000_000
f = synth.crowd_frame(n, seed=1, extent=50.0)
hp = ops.HostFramePipeline(max_points=n, voxel_size=0.05, grid_size=0.5, slots=2, max_key_space=1 << 28, max_nx=256, max_ny=256)
for _ in range(3): hp.process(f)
ts = {}
for rep in range(10):
    t0 = time.perf_counter(); hp.submit(f); t1 = time.perf_counter(); out = hp.collect(); t2 = time.perf_counter()
    ts.setdefault('submit', []).append(t1 - t0); ts.setdefault('collect', []).append(t2 - t1)
print({k: round(float(np.median(v)) * 1e3, 3) for k, v in ts.items()}, 'ms per frame (sync process)')
# scan density call breakdown
pts = torch.from_numpy(synth.crowd_frame(4_000_000, seed=2, extent=200.0, extent_y=150.0)).cuda()
ctx = sharding.ScanDensity(torch.device('cuda', 0))
for _ in range(3): ctx(pts, 0.5)
tt = {}
for rep in range(10):
    torch.cuda.synchronize(); t0 = time.perf_counter(); ctx.enqueue(pts, 0.5); t1 = time.perf_counter(); d_ = ctx._wait_desc(); t2 = time.perf_counter()
    r = ctx.result(); t3 = time.perf_counter()
    for k, v in (('enqueue', t1 - t0), ('wait_desc', t2 - t1), ('result_after_desc', t3 - t2)): tt.setdefault(k, []).append(v)
print({k: round(float(np.median(v)) * 1e3, 3) for k, v in tt.items()}, 'ms (4 M points, 804x604 grid)')
