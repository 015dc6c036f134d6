import sys, numpy as np, torch
sys.path.insert(0, '.')
from lidar_ai_recommendation_software_b200 import ops, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100003
pts = synth.crowd_frame(n, seed=5, extent=50.0)
d = torch.from_numpy(pts).cuda()
pipe = ops.FramePipeline(max_points=n, voxel_size=0.05, grid_size=0.5, max_key_space=1 << 28, max_nx=512, max_ny=512)
ops.set_frame_mode(1, 0, 0, 0)
pipe.enqueue(d); r = pipe.result(); base = r.inverse.clone(); nv = r.n_voxels
for cfg in [(3, 512, 1, 0), (2, 512, 1, 0), (3, 256, 1, 0), (2, 512, 1, 0), (3, 128, 1, 0), (2, 512, 1, 0), (3, 512, 1, 0), (1, 0, 0, 0)]:
    ops.set_frame_mode(*cfg)
    print('cfg', cfg, flush=True)
    pipe.enqueue(d); torch.cuda.synchronize(); r = pipe.result()
    print('  ok', r.n_voxels == nv and torch.equal(r.inverse, base), flush=True)
