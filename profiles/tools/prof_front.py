import sys, time, torch, numpy as np
sys.path.insert(0, '.')
from lidar_ai_recommendation_software_b200 import ops, preprocess as pre, synth
from lidar_ai_recommendation_software_b200.models.crowd_flow_model import CrowdFlowModel
dev = torch.device('cuda')
f = synth.ring_sequence_frame(1)
h64 = torch.from_numpy(np.ascontiguousarray(f[:, :3], dtype=np.float64)).pin_memory()
d = h64.to(dev)
model = CrowdFlowModel()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for rep in range(reps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    pd = pre.run(d, variant='B', host_arrays=False)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    r = model.analyze_sequence_frame(pd)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f'pre.run {1e3*(t1-t0):.3f} ms  analyze_sequence_frame {1e3*(t2-t1):.3f} ms', flush=True)
