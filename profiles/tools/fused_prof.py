import sys, torch
sys.path.insert(0, '.')
from lidar_ai_recommendation_software_b200 import ops, synth, _capi
n = 1_000_000
frames = [torch.from_numpy(synth.crowd_frame(n, seed=s, extent=50.0)).cuda() for s in range(16)]
pipe = ops.FramePipeline(max_points=n, voxel_size=0.05, grid_size=0.5, max_key_space=1 << 28, max_nx=256, max_ny=256, scan_order=False)
ops.set_frame_mode(ops.FRAME_FUSED, 512, 1, 0)
_capi.check(_capi.lib.lidar_frame_set_fused_l2_persist(24 << 20))
torch.cuda.synchronize()
ops.set_frame_streaming(True, inputs_complete=True)      # the launch mode bench.py times
for i in range(40):
    pipe.enqueue(frames[i % 16])
torch.cuda.synchronize()
ops.set_frame_streaming(False)
print(pipe.result().n_voxels)
