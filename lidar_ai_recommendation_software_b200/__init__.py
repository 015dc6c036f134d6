"""lidar_ai_recommendation_software_b200 — B200-native (sm_100a) point-cloud hot path of
FortuneMU2025/LIDAR_AI_Recommendation_Software, behind the reference's own Python call surface.

Layout
  csrc/ + include/lidar_b200.h   hand-written CUDA kernels and their C ABI (liblidar_b200.so)
  _capi.py                       ctypes binding of that ABI (no CPU fallback: import fails loudly)
  ops.py                         device-level operators on torch CUDA tensors
  utils/data_processing.py       drop-in for the reference's utils.data_processing   (surface A)
  models/crowd_density_model.py  drop-in for models.crowd_density_model               (surface A)
  models/crowd_flow_model.py     drop-in for models.crowd_flow_model                  (surface A)
  apps.py                        drop-in for the functions inlined in app_simplified.py (surface B)
  pointnet2.py                   set abstraction: FPS, ball query, grouping, shared MLP (NEW ops)
  sharding.py                    frames across GPUs / one scan sharded by points (torch.distributed)
  synth.py                       seeded synthetic inputs (pure numpy)
"""
__version__ = "0.1.0"

# NOTE: `_capi` (and therefore every compute module) raises ImportError if liblidar_b200.so has not
# been built — there is no CPU fallback.  `build` and `synth` stay importable without it so that the
# library can be compiled and inputs generated first.
