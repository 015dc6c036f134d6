import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from lidar_ai_recommendation_software_b200 import ops, synth
f = synth.ring_sequence_frame(1)
d = torch.from_numpy(np.ascontiguousarray(f[:, :3], dtype=np.float64)).cuda()
desc, inl, col, ng, idx, X = ops.preprocess_front(d, want_colors=False)
lo, hi = np.array(desc.bbox_ng[:3]), np.array(desc.bbox_ng[3:])
for _ in range(3): ops.dbscan(ng, 0.3, 5, tol=0.0, bounds=(lo, hi))
torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): labels, info = ops.dbscan(ng, 0.3, 5, tol=0.0, bounds=(lo, hi), defer=True)
e1.record(); torch.cuda.synchronize()
print('dbscan ms', e0.elapsed_time(e1) / 20, 'clusters', int(info[0].item()) & 0xffffffff)
