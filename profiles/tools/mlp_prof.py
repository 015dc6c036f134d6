import sys, torch
sys.path.insert(0, '.')
from lidar_ai_recommendation_software_b200 import pointnet2 as pn, synth
B, N, M, K, R = 16, 16384, 1024, 32, 0.2
xyz = torch.from_numpy(synth.sa_batch(B, N, seed=0)).cuda()
ws, bs = synth.sa_weights(seed=1)
ws = [torch.from_numpy(w).cuda() for w in ws]; bs = [torch.from_numpy(b).cuda() for b in bs]
fps = pn.furthest_point_sample(xyz, M); new_xyz = pn.gather_points(xyz, fps); idx = pn.ball_query(xyz, new_xyz, R, K)
for _ in range(5):
    out = pn.shared_mlp_maxpool(ws, bs, xyz=xyz, idx=idx, new_xyz=new_xyz, impl=pn.MLP_TCGEN05)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    out = pn.shared_mlp_maxpool(ws, bs, xyz=xyz, idx=idx, new_xyz=new_xyz, impl=pn.MLP_TCGEN05)
e1.record(); torch.cuda.synchronize()
print('mlp us', e0.elapsed_time(e1) / 20 * 1e3, float(out.sum()))
