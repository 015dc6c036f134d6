import sys, torch
sys.path.insert(0, '.')
from lidar_ai_recommendation_software_b200 import ops, synth
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = 1_000_000
frames = [torch.from_numpy(synth.crowd_frame(n, seed=s, extent=50.0)).cuda() for s in range(4)]
pipe = ops.FramePipeline(max_points=n, voxel_size=0.05, grid_size=0.5, max_key_space=1 << 28, max_nx=256, max_ny=256, scan_order=False)
ops.set_frame_mode(mode, 512, 1, 0)
for i in range(12):
    pipe.enqueue(frames[i % 4])
torch.cuda.synchronize()
print(pipe.result().n_voxels)
