"""CPU tests of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and exports every
symbol include/*.h declares (no compute calls here — there is no GPU in the build container)."""
import ctypes
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "lidar_b200.h"


def declared_functions():
    src = HEADER.read_text()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    names = re.findall(r"\b(lidar_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_declares_functions():
    names = declared_functions()
    assert "lidar_last_error" in names and "lidar_frame_voxel_density" in names
    assert len(names) >= 10


def test_library_exports_every_declared_symbol():
    from lidar_ai_recommendation_software_b200 import _capi
    lib = ctypes.CDLL(str(_capi.LIB_PATH))
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in {HEADER.name} but not exported"


def test_python_prototypes_cover_the_header():
    from lidar_ai_recommendation_software_b200 import _capi
    assert sorted(_capi.PROTOTYPES) == declared_functions()
    assert _capi.abi_version() == 2
    assert isinstance(_capi.last_error(), str)


def test_struct_mirrors_match_c_layout():
    """sizeof/offsets of the ctypes mirrors must equal the C structs (checked with a tiny C program)."""
    from lidar_ai_recommendation_software_b200 import _capi
    prog = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "lidar_b200.h"
    int main(void){
      printf("%zu %zu %zu %zu %zu %zu\n", sizeof(lidar_frame_desc), offsetof(lidar_frame_desc, dims),
             offsetof(lidar_frame_desc, key_space), offsetof(lidar_frame_desc, nx),
             offsetof(lidar_frame_desc, n_voxels), sizeof(lidar_frame_caps));
      printf("%zu %zu %zu %zu %zu %zu\n", sizeof(lidar_front_desc), offsetof(lidar_front_desc, n_in),
             offsetof(lidar_front_desc, z_thr), offsetof(lidar_front_desc, plane),
             offsetof(lidar_front_desc, eps), offsetof(lidar_front_desc, key_ng));
      printf("%zu %zu %zu %zu %zu %zu\n", sizeof(lidar_scan_desc), offsetof(lidar_scan_desc, n_local),
             offsetof(lidar_scan_desc, status), sizeof(lidar_scan_comm), offsetof(lidar_scan_comm, peer_ptrs),
             offsetof(lidar_scan_comm, multicast_ptr));
      printf("%zu %zu %zu %zu\n", sizeof(lidar_sorted_desc), offsetof(lidar_sorted_desc, dims),
             offsetof(lidar_sorted_desc, n_voxels), offsetof(lidar_sorted_desc, status));
      printf("%zu %zu %zu %zu\n", sizeof(lidar_sequence_frame_out), offsetof(lidar_sequence_frame_out, n_clusters),
             offsetof(lidar_sequence_frame_out, guard_dbscan), offsetof(lidar_sequence_frame_out, need_centroid_ws));
      return 0; }'''
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        c = Path(td) / "t.c"
        c.write_text(prog)
        exe = Path(td) / "t"
        subprocess.run(["gcc", "-I", str(ROOT / "include"), str(c), "-o", str(exe)], check=True)
        out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    D = _capi.FrameDesc
    F = _capi.FrontDesc
    want = [ctypes.sizeof(D), D.dims.offset, D.key_space.offset, D.nx.offset, D.n_voxels.offset,
            ctypes.sizeof(_capi.FrameCaps),
            ctypes.sizeof(F), F.n_in.offset, F.z_thr.offset, F.plane.offset, F.eps.offset, F.key_ng.offset,
            ctypes.sizeof(_capi.ScanDesc), _capi.ScanDesc.n_local.offset, _capi.ScanDesc.status.offset,
            ctypes.sizeof(_capi.ScanComm), _capi.ScanComm.peer_ptrs.offset, _capi.ScanComm.multicast_ptr.offset,
            ctypes.sizeof(_capi.SortedDesc), _capi.SortedDesc.dims.offset, _capi.SortedDesc.n_voxels.offset,
            _capi.SortedDesc.status.offset,
            ctypes.sizeof(_capi.SequenceFrameOut), _capi.SequenceFrameOut.n_clusters.offset,
            _capi.SequenceFrameOut.guard_dbscan.offset, _capi.SequenceFrameOut.need_centroid_ws.offset]
    assert [int(x) for x in out] == want


def test_only_sm100a_code_is_embedded():
    from lidar_ai_recommendation_software_b200 import _capi
    res = subprocess.run(["cuobjdump", "-lelf", str(_capi.LIB_PATH)], capture_output=True, text=True)
    if res.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", res.stdout))
    assert archs == {"sm_100a"}, archs


def test_argument_errors_do_not_need_a_gpu():
    """Validation happens before any CUDA call: bad arguments return LIDAR_ERR_INVALID with a message."""
    from lidar_ai_recommendation_software_b200 import _capi
    rc = _capi.lib.lidar_hist2d_f64(None, 1, None, 1, 10, None, 0, None, 0, None, 0, None)
    assert rc == -1 and "lidar_hist2d_f64" in _capi.last_error()
    with pytest.raises(_capi.LidarError):
        _capi.check(rc)


def test_thin_torch_extension_builds_and_registers_its_operators():
    """The north star's "thin PyTorch C++ extension with a C-ABI core": lidar_b200_torch.so is built in-tree next to
    liblidar_b200.so, registers torch.ops.lidar_b200.* and reports the ABI version of the core it links against."""
    import torch
    from lidar_ai_recommendation_software_b200 import _capi, _torch_ext
    assert _torch_ext.EXT_PATH.exists() and _torch_ext.EXT_PATH.parent == _capi.LIB_PATH.parent
    assert int(torch.ops.lidar_b200.abi_version()) == _capi.abi_version()
    for name in ("frame_voxel_density", "hist2d_points", "bbox"):
        assert hasattr(torch.ops.lidar_b200, name)
    with pytest.raises((RuntimeError, NotImplementedError)):      # CPU tensors: no kernel registered for that backend
        torch.ops.lidar_b200.bbox(torch.zeros((4, 4)), torch.zeros(8, dtype=torch.float64), torch.zeros(16, dtype=torch.uint8))
