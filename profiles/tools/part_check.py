import sys, numpy as np, torch
sys.path.insert(0, '.')
from lidar_ai_recommendation_software_b200 import ops, synth
sizes = [int(a) for a in sys.argv[1:]] or [777, 4736, 100003, 1000000]
for n in sizes:
    pts = synth.crowd_frame(n, seed=5, extent=10.0 if n < 50000 else 50.0)
    d = torch.from_numpy(pts).cuda()
    pipe = ops.FramePipeline(max_points=n, voxel_size=0.05, grid_size=0.5, max_key_space=1 << 28, max_nx=512, max_ny=512)
    ops.set_frame_mode(1, 0, 0, 0)
    pipe.enqueue(d); r = pipe.result()
    base = dict(inv=r.inverse.clone(), key=r.voxel_key.clone(), vox=pipe.voxels[:r.n_voxels].clone(), grid=r.grid_counts.clone(), nv=r.n_voxels)
    ops.set_frame_mode(3, 512, 1, 0)
    for rep in range(2):
        pipe.enqueue(d); torch.cuda.synchronize(); r = pipe.result()
        ok = dict(nv=r.n_voxels == base['nv'], inv=torch.equal(r.inverse, base['inv']), key=torch.equal(r.voxel_key, base['key']),
                  grid=torch.equal(r.grid_counts, base['grid']),
                  vox=r.n_voxels == base['nv'] and torch.equal(pipe.voxels[:r.n_voxels], base['vox']))
        print(n, rep, ok, 'trace', [x for x in r.desc.trace_ns][:16], flush=True)
        if not all(ok.values()):
            if not ok['inv']:
                bad = (r.inverse != base['inv']).nonzero().flatten()
                print(' inverse mismatches', bad.numel(), bad[:5].tolist(), r.inverse[bad[:5]].tolist(), base['inv'][bad[:5]].tolist())
            if not ok['grid']:
                print(' grid sum', int(r.grid_counts.sum()), int(base['grid'].sum()))
            print(' nv', r.n_voxels, base['nv'])
