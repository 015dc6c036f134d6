"""A/B of the staging copy with non-temporal stores (lidar_host_copy_nontemporal): the copy alone, the copy while both
DMA directions are busy, and the end-to-end api leg of bench.py, alternating the two settings on one box."""
import sys, time, threading, numpy as np, torch
sys.path.insert(0, '.')
import bench
from lidar_ai_recommendation_software_b200 import ops, synth, _capi
lib = _capi.lib
dev = torch.device('cuda', 0)
torch.cuda.set_device(dev)
ops.bind_to_device_numa(dev)
nb = 16 << 20
pin = ops._PinnedBlock(nb)
srcs = [np.full(nb, i, dtype=np.uint8) for i in range(16)]
big = 128 << 20
pa, pb = ops._PinnedBlock(big), ops._PinnedBlock(big)
da, db_ = torch.empty(big, dtype=torch.uint8, device=dev), torch.empty(big, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def copy_rate(reps=64):
    for i in range(4): lib.lidar_host_memcpy(pin.ptr, srcs[i].ctypes.data, nb)
    t0 = time.perf_counter()
    for i in range(reps): lib.lidar_host_memcpy(pin.ptr, srcs[i % 16].ctypes.data, nb)
    return reps * nb / (time.perf_counter() - t0) / 1e9

def copy_rate_under_dma():
    stop = [False]
    def dma():
        while not stop[0]:
            for _ in range(4):
                lib.lidar_copy_async(da.data_ptr(), pa.ptr, big, 1, s1.cuda_stream)
                lib.lidar_copy_async(pb.ptr, db_.data_ptr(), big, 0, s2.cuda_stream)
            s1.synchronize(); s2.synchronize()
    th = threading.Thread(target=dma); th.start()
    time.sleep(0.05)
    r = copy_rate()
    stop[0] = True; th.join()
    return r

for nt in (0, 1, 0, 1):
    lib.lidar_host_copy_nontemporal(nt)
    print(f"nt={nt}: staging copy alone {copy_rate():.1f} GB/s, under H2D+D2H DMA {copy_rate_under_dma():.1f} GB/s", flush=True)

n = 1_000_000
frames = [synth.crowd_frame(n, seed=s, extent=50.0) for s in range(16)]
for nt in (0, 1, 0, 1, 0, 1):
    lib.lidar_host_copy_nontemporal(nt)
    r = bench.e2e_leg(ops, torch, None, 1, dev, n, frames, 240, 3, 3, "api")
    print(f"nt={nt}: e2e api {r['value']:.0f} Mpoints/s", flush=True)
