"""Loader of the thin PyTorch C++ extension (csrc_torch/lidar_torch_ext.cpp -> lidar_b200_torch.so).

`torch.ops.lidar_b200.*` are the per-frame hot calls of the package (frame kernel, histogram, bounding box): the
extension checks the tensors, takes PyTorch's current CUDA stream in C++ and calls the C ABI of liblidar_b200.so.  The
complete binding of the ABI stays the ctypes one (`_capi.py`); both sit on the same extern "C" core.  Missing library =>
ImportError: there is no fallback.
"""
from __future__ import annotations

from pathlib import Path

import torch

from . import _capi  # noqa: F401  (liblidar_b200.so must be loaded first: the extension links against it)

EXT_PATH = Path(__file__).resolve().parent / "lidar_b200_torch.so"

if not EXT_PATH.exists():
    raise ImportError(f"{EXT_PATH} is missing: build it with `python -m lidar_ai_recommendation_software_b200.build` "
                      "(or __graft_entry__.build()). There is no fallback.")
torch.ops.load_library(str(EXT_PATH))
ops = torch.ops.lidar_b200
if int(ops.abi_version()) != _capi.abi_version():
    raise ImportError("lidar_b200_torch.so and liblidar_b200.so were built from different headers")
