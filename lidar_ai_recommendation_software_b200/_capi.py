"""ctypes binding of the C-ABI CUDA core (include/lidar_b200.h -> liblidar_b200.so).

There is no CPU fallback: if the shared library is missing or does not load, importing this module
raises, and every op of the package fails loudly.  PyTorch is used only to own device memory and
streams; the pointers handed across the ABI are raw `data_ptr()`s.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import torch  # noqa: F401  (loads libcudart.so.12 and friends before our library is opened)

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "liblidar_b200.so"

LIDAR_OK = 0
FMT_F32X4 = 0
FMT_F64X3 = 1
HIST_AUTO, HIST_GLOBAL, HIST_SHARED = 0, 1, 2
MLP_AUTO, MLP_SIMT, MLP_TCGEN05 = 0, 1, 2

ERR_NAMES = {-1: "LIDAR_ERR_INVALID", -2: "LIDAR_ERR_CUDA", -3: "LIDAR_ERR_WORKSPACE", -4: "LIDAR_ERR_CAPACITY"}


class LidarError(RuntimeError):
    """Raised when a C-ABI call returns a negative status (apps catch `Exception`, app.py:103-104)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class FrameDesc(C.Structure):
    """Mirror of `lidar_frame_desc` (include/lidar_b200.h)."""

    _fields_ = [
        ("origin", C.c_double * 4),
        ("bbox_min", C.c_double * 4),
        ("bbox_max", C.c_double * 4),
        ("voxel", C.c_double),
        ("fix_scale_xyz", C.c_double),
        ("fix_scale_w", C.c_double),
        ("dims", C.c_int32 * 4),
        ("key_space", C.c_int64),
        ("grid", C.c_double),
        ("ex0", C.c_double), ("ex1", C.c_double), ("exd", C.c_double),
        ("ey0", C.c_double), ("ey1", C.c_double), ("eyd", C.c_double),
        ("nx", C.c_int32), ("ny", C.c_int32),
        ("n_points", C.c_int64),
        ("n_voxels", C.c_int64),
        ("status", C.c_int32),
        ("fast_f32", C.c_int32),
        ("magic_dz", C.c_uint32), ("magic_dy", C.c_uint32),
        ("shift_dz", C.c_int32), ("shift_dy", C.c_int32),
        ("trace_ns", C.c_uint32 * 16),
    ]


class FrameCaps(C.Structure):
    """Mirror of `lidar_frame_caps`."""

    _fields_ = [
        ("max_points", C.c_int64),
        ("max_key_space", C.c_int64),
        ("max_nx", C.c_int32),
        ("max_ny", C.c_int32),
    ]


class FrontDesc(C.Structure):
    """Mirror of `lidar_front_desc`: every scalar the chained preprocess leaves on the device."""

    _fields_ = [
        ("bbox_raw", C.c_double * 8), ("sum1", C.c_double * 6), ("mean", C.c_double * 3), ("sum2", C.c_double * 6),
        ("std", C.c_double * 3), ("thr", C.c_double * 3), ("tol", C.c_double * 3),
        ("zmin", C.c_double), ("zden", C.c_double),
        ("n_in", C.c_int64), ("guard_sigma", C.c_uint64),
        ("kth", C.c_double * 2), ("z_thr", C.c_double),
        ("n_nonground", C.c_int64), ("guard_ground", C.c_uint64),
        ("plane", C.c_double * 10), ("bbox_in", C.c_double * 6), ("bbox_ng", C.c_double * 6),
        ("t1", C.c_double * 6), ("sc_mean", C.c_double * 3), ("t2", C.c_double * 6), ("scale", C.c_double * 3),
        ("u1", C.c_double * 6), ("xm", C.c_double * 3), ("u2", C.c_double * 6), ("xstd", C.c_double * 3),
        ("eps", C.c_double),
        ("key_in", C.c_uint64 * 6), ("key_ng", C.c_uint64 * 6),
    ]


class SequenceFrameOut(C.Structure):
    """Mirror of `lidar_sequence_frame_out`."""

    _fields_ = [("front", FrontDesc), ("n_clusters", C.c_int32), ("centroids_done", C.c_int32), ("guard_dbscan", C.c_uint64),
                ("need_dbscan_ws", C.c_uint64), ("need_centroid_ws", C.c_uint64)]


class SortedDesc(C.Structure):
    """Mirror of `lidar_sorted_desc` (the 64-bit-key sort path of voxel downsample)."""

    _fields_ = [
        ("origin", C.c_double * 3), ("bbox_min", C.c_double * 3), ("bbox_max", C.c_double * 3),
        ("voxel", C.c_double), ("fix_scale_xyz", C.c_double), ("fix_scale_w", C.c_double),
        ("dims", C.c_int64 * 3), ("key_space", C.c_int64), ("n_points", C.c_int64), ("n_kept", C.c_int64),
        ("n_voxels", C.c_int64), ("passes", C.c_int32), ("status", C.c_int32),
    ]


class ScanDesc(C.Structure):
    """Mirror of `lidar_scan_desc`: the point-sharded density grid's device-derived descriptor."""

    _fields_ = [
        ("bbox", C.c_double * 4), ("grid", C.c_double),
        ("ex0", C.c_double), ("ex1", C.c_double), ("exd", C.c_double),
        ("ey0", C.c_double), ("ey1", C.c_double), ("eyd", C.c_double),
        ("n_local", C.c_int64), ("nx", C.c_int32), ("ny", C.c_int32), ("status", C.c_int32), ("pad", C.c_int32),
    ]


class ScanComm(C.Structure):
    """Mirror of `lidar_scan_comm`: where the peers' symmetric buffers are mapped in this process."""

    _fields_ = [
        ("rank", C.c_int32), ("world", C.c_int32),
        ("symm_bytes", C.c_size_t), ("peer_ptrs", C.c_void_p * 16), ("multicast_ptr", C.c_void_p),
    ]


SCAN_EMPTY = 1
NCCL_SUM_I32, NCCL_MAX_F64 = 0, 1

_vp, _i32, _i64, _sz, _dbl = C.c_void_p, C.c_int, C.c_int64, C.c_size_t, C.c_double
HOST_UNIQUE_KEYS, HOST_NO_PER_POINT = 1, 2     # flags of lidar_frame_voxel_density_host
FRONT_COLORS, FRONT_SCALER = 1, 2              # flags of lidar_preprocess_front

# name -> (restype, argtypes).  tests/test_abi.py checks this table against include/lidar_b200.h.
PROTOTYPES: dict[str, tuple] = {
    "lidar_last_error": (C.c_char_p, []),
    "lidar_abi_version": (_i32, []),
    "lidar_device_props": (_i32, [_i32, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_size_t)]),
    "lidar_reduce_workspace_bytes": (_sz, []),
    "lidar_bbox": (_i32, [_vp, _i32, _i64, _vp, _vp, _sz, _vp]),
    "lidar_moments": (_i32, [_vp, _i32, _i64, C.POINTER(C.c_double), _vp, _vp, _sz, _vp]),
    "lidar_centroid_distances": (_i32, [_vp, _i64, _vp, _vp, _sz, _vp]),
    "lidar_hist2d_f64": (_i32, [_vp, _i64, _vp, _i64, _i64, _vp, _i32, _vp, _i32, _vp, _i32, _vp]),
    "lidar_hist2d_points": (_i32, [_vp, _i32, _i64, _vp, _i32, _vp, _i32, _vp, _i32, _vp]),
    "lidar_compact_workspace_bytes": (_sz, [_i64]),
    "lidar_roi_crop": (_i32, [_vp, _i32, _i64, C.POINTER(C.c_double), C.POINTER(C.c_double), _vp, _vp, _vp,
                              _vp, _sz, _vp]),
    "lidar_preprocess_workspace_bytes": (_sz, [_i64]),
    "lidar_preprocess_front_workspace_bytes": (_sz, [_i64]),
    "lidar_preprocess_front": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "lidar_sigma_filter": (_i32, [_vp, _i64, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                  _dbl, _dbl, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "lidar_select_kth": (_i32, [_vp, _i64, _i64, _i64, _vp, _vp, _sz, _vp]),
    "lidar_ground_split": (_i32, [_vp, _i64, _dbl, C.POINTER(C.c_double), _dbl, _vp, _vp, _vp, _vp, _vp, _vp,
                                  _sz, _vp]),
    "lidar_standardize": (_i32, [_vp, _i64, C.POINTER(C.c_double), C.POINTER(C.c_double), _vp, _vp]),
    "lidar_gather_rows": (_i32, [_vp, _i64, _i32, _vp, _i64, _vp, _vp]),
    "lidar_scatter_labels": (_i32, [_vp, _vp, _i64, _vp, _i64, _vp]),
    "lidar_dbscan_set_dense": (_i32, [_i32]),
    "lidar_dbscan_workspace_bytes": (_sz, [_i64, _dbl, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "lidar_dbscan": (_i32, [_vp, _i64, _dbl, _i32, _dbl, C.POINTER(C.c_double), C.POINTER(C.c_double), _vp, _vp,
                            _vp, _vp, _sz, _vp]),
    "lidar_ball_count_workspace_bytes": (_sz, [_i64, _dbl, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "lidar_ball_count": (_i32, [_vp, _i64, _dbl, C.POINTER(C.c_double), C.POINTER(C.c_double), _vp, _vp, _sz, _vp]),
    "lidar_centroid_workspace_bytes": (_sz, [_i32]),
    "lidar_cluster_centroids": (_i32, [_vp, _vp, _i32, _i64, _i32, _vp, _vp, _vp, _sz, _vp]),
    "lidar_flow_field": (_i32, [_vp, _i32, _vp, _i32, _dbl, _dbl, _dbl, _dbl, C.POINTER(C.c_double), _i32, _dbl,
                                _i32, _dbl, _dbl, _vp, _vp, _vp, _vp, _vp]),
    "lidar_flow_bottlenecks": (_i32, [_i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "lidar_flow_box_max": (_i32, [_i32, _i32, _vp, _vp, _dbl, _vp, _vp]),
    "lidar_radius_count": (_i32, [_vp, _i32, _vp, _i32, _vp, _i32, _dbl, _vp, _vp]),
    "lidar_nearest_grid_cell": (_i32, [_vp, _i32, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "lidar_frame_flow_match": (_i32, [_vp, _i32, _vp, _i32, C.c_float, C.c_float, _vp, _vp, _vp]),
    "lidar_frame_flow_field": (_i32, [_vp, _i32, _vp, _vp, _vp, _i32, _dbl, _vp, _vp, _vp]),
    "lidar_sequence_frame_b": (_i32, [_vp, _i64, _dbl, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _sz,
                               _vp, _sz, _vp, _sz, _vp, _sz, _vp, _vp]),
    "lidar_frame_flow": (_i32, [_vp, _i32, _vp, _i32, C.c_float, C.c_float, _vp, _i32, _vp, _i32, _dbl, _vp, _vp, _vp, _vp, _vp, _vp]),
    "lidar_fps_workspace_bytes": (_sz, [_i32, _i32]),
    "lidar_fps": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp, _sz, _vp]),
    "lidar_ball_query": (_i32, [_vp, _vp, _i32, _i32, _i32, C.c_float, _i32, _vp, _vp]),
    "lidar_group_points": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "lidar_gather_points": (_i32, [_vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    "lidar_shared_mlp_maxpool": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32,
                                        _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp]),
    "lidar_frame_set_ctas_per_sm": (_i32, [_i32]),
    "lidar_frame_set_fused": (_i32, [_i32, _i32, _i32, _i32]),
    "lidar_frame_set_partition_auto": (_i32, [_i32]),
    "lidar_frame_set_fused_plain_launch": (_i32, [_i32]),
    "lidar_frame_set_fused_pdl": (_i32, [_i32]),
    "lidar_frame_set_fused_l2_persist": (_i32, [_sz]),
    "lidar_frame_set_fused_scan_order": (_i32, [_i32]),
    "lidar_frame_trace_offset": (_sz, [C.POINTER(FrameCaps)]),
    "lidar_frame_workspace_bytes": (_sz, [C.POINTER(FrameCaps)]),
    "lidar_frame_pack_soa": (_i32, [_vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "lidar_frame_host_block_bytes": (_sz, [_i64, _vp, _i32]),
    "lidar_frame_host_block_layout": (_i32, [_i64, _vp, _i32, C.POINTER(C.c_size_t)]),
    "lidar_frame_voxel_density_host": (_i32, [_vp, _i64, _dbl, _dbl, C.POINTER(C.c_double), C.POINTER(C.c_double), _vp,
                                              _vp, _vp, _vp, _i32, _vp, _vp, _sz, _vp]),
    "lidar_frame_voxel_density_host_begin": (_i32, [_vp, _i64, _dbl, _dbl, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                                    _vp, _vp, _vp, _vp, _i32, _vp, _vp, _sz, _vp]),
    "lidar_frame_host_fetch": (_i32, [_i64, _i64, _i32, _i32, _vp, _vp, _i32, _vp, _vp]),
    "lidar_scan_workspace_bytes": (_sz, []),
    "lidar_scan_workspace_init": (_i32, [_vp, _sz, _vp]),
    "lidar_scan_symm_bytes": (_sz, [_i64]),
    "lidar_scan_symm_grid_offset": (_sz, []),
    "lidar_scan_density": (_i32, [_vp, _i32, _i64, _dbl, _i32, _i32, _i64, _vp, _vp, _vp, _vp, _vp, _vp,
                                  C.POINTER(ScanComm), C.c_uint32, _vp, _sz, _vp]),
    "lidar_scan_bbox_packed": (_i32, [_vp, _i32, _i64, _vp, _vp, _sz, _vp]),
    "lidar_scan_hist": (_i32, [_vp, _i32, _i64, _vp, _dbl, _i32, _i32, _i64, _vp, _vp, _vp]),
    "lidar_scan_finish": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "lidar_nccl_available": (_i32, []),
    "lidar_nccl_unique_id": (_i32, [_vp]),
    "lidar_nccl_comm_init": (_i32, [_vp, _i32, _i32, C.POINTER(C.c_void_p)]),
    "lidar_nccl_comm_destroy": (_i32, [_vp]),
    "lidar_nccl_allreduce": (_i32, [_vp, _vp, _i64, _i32, _vp]),
    "lidar_bind_to_device_numa": (_i32, [_i32, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "lidar_host_alloc": (_i32, [_sz, C.POINTER(C.c_void_p)]),
    "lidar_host_free": (_i32, [_vp]),
    "lidar_host_copy_threads": (_i32, [_i32]),
    "lidar_host_copy_nontemporal": (_i32, [_i32]),
    "lidar_host_memcpy": (_i32, [_vp, _vp, _sz]),
    "lidar_host_memcpy_batch": (_i32, [_i32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "lidar_host_copy_wake": (_i32, []),
    "lidar_copy_async": (_i32, [_vp, _vp, _sz, _i32, _vp]),
    "lidar_stream_synchronize": (_i32, [_vp]),
    "lidar_voxel_sorted_workspace_bytes": (_sz, [_i64]),
    "lidar_voxel_downsample_sorted": (_i32, [_vp, _i64, _dbl, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                             C.POINTER(C.c_double), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "lidar_frame_workspace_init": (_i32, [_vp, _sz, C.POINTER(FrameCaps), _vp]),
    "lidar_frame_voxel_density": (_i32, [_vp, _i64, _dbl, _dbl, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                         _vp, _vp, _vp, _vp, _vp, C.POINTER(FrameCaps), _vp, _sz, _vp]),
    "lidar_frame_voxel_density_timed": (_i32, [_vp, _i64, _dbl, _dbl, C.POINTER(C.c_double),
                                               C.POINTER(C.c_double), _vp, _vp, _vp, _vp, _vp,
                                               C.POINTER(FrameCaps), _vp, _sz, _vp, C.POINTER(C.c_void_p)]),
}


def _load() -> C.CDLL:
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA core has not been built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (or "
            "`python -m lidar_ai_recommendation_software_b200.build`). There is no CPU fallback.")
    try:
        lib = C.CDLL(str(LIB_PATH), mode=getattr(os, "RTLD_NOW", 2))
    except OSError as e:  # pragma: no cover - environment specific
        raise ImportError(f"cannot load {LIB_PATH}: {e}. There is no CPU fallback.") from e
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def last_error() -> str:
    return (lib.lidar_last_error() or b"").decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != LIDAR_OK:
        raise LidarError(rc, last_error())


def abi_version() -> int:
    return int(lib.lidar_abi_version())
