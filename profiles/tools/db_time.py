import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from lidar_ai_recommendation_software_b200 import ops, synth
from oracle import np_semantics as nps
for fr in (1, 7):
    f = synth.ring_sequence_frame(fr)
    d = torch.from_numpy(np.ascontiguousarray(f[:, :3], dtype=np.float64)).cuda()
    desc, inl, col, ng, idx, X = ops.preprocess_front(d, want_colors=False)
    lo, hi = np.array(desc.bbox_ng[:3]), np.array(desc.bbox_ng[3:])
    for _ in range(3):
        labels, nc, g = ops.dbscan(ng, 0.3, 5, tol=0.0, bounds=(lo, hi))
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); labels, info = ops.dbscan(ng, 0.3, 5, tol=0.0, bounds=(lo, hi), defer=True); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print('frame', fr, 'm', ng.shape[0], 'clusters', nc, 'dbscan ms', round(float(np.median(ts)), 3), 'labels sha', hash(labels.cpu().numpy().tobytes()) & 0xffffffff)
