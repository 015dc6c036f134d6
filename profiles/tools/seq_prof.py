"""One 128-beam frame through the per-frame path of configs[3], serially (for an ncu launch list): frame pipeline,
preprocess (variant B) with DBSCAN, people positions, flow."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from lidar_ai_recommendation_software_b200 import ops, preprocess, synth
from lidar_ai_recommendation_software_b200.models.crowd_flow_model import CrowdFlowModel
dev = torch.device('cuda', 0)
pool = [synth.ring_sequence_frame(i) for i in range(2)]
pool64 = [torch.from_numpy(np.ascontiguousarray(p[:, :3], dtype=np.float64)).pin_memory() for p in pool]
dpool = [torch.from_numpy(p).to(dev) for p in pool]
pipe = ops.FramePipeline(max_points=max(p.shape[0] for p in pool), voxel_size=0.05, grid_size=0.5,
                         max_key_space=(1 << 31) - 1, max_nx=1024, max_ny=1024, device=dev)
model = CrowdFlowModel()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
for i in range(n):
    k = i % 2
    if i == n - 2:
        torch.cuda.synchronize(); print("MARK last two frames", flush=True)
    pipe.enqueue(dpool[k])
    d64 = pool64[k].to(dev, non_blocking=True)
    pd = preprocess.run(d64, variant="B", host_arrays=False)
    res = model.analyze_sequence_frame(pd, dt=0.1)
torch.cuda.synchronize()
print("clusters", pd[preprocess.DEVICE_KEY].n_clusters)
