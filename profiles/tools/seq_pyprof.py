"""Where the HOST time of one sequence frame goes (configs[3] per-frame path, one thread): cProfile over 60 frames."""
import cProfile, pstats, sys, time, numpy as np, torch
sys.path.insert(0, '.')
from lidar_ai_recommendation_software_b200 import ops, preprocess, synth
from lidar_ai_recommendation_software_b200.models.crowd_flow_model import CrowdFlowModel
dev = torch.device('cuda', 0)
pool = [synth.ring_sequence_frame(i) for i in range(2)]
pool64 = [torch.from_numpy(np.ascontiguousarray(p[:, :3], dtype=np.float64)).pin_memory() for p in pool]
model = CrowdFlowModel()
def frame(i):
    d64 = pool64[i % 2].to(dev, non_blocking=True)
    pd = preprocess.run(d64, variant="B", host_arrays=False)
    return model.analyze_sequence_frame(pd, dt=0.1)
for i in range(6): frame(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(60): frame(i)
torch.cuda.synchronize()
print('serial ms/frame', (time.perf_counter() - t0) / 60 * 1e3)
pr = cProfile.Profile(); pr.enable()
for i in range(60): frame(i)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr); st.strip_dirs(); st.sort_stats('cumulative').print_stats(45); st.sort_stats('tottime').print_stats(25)
