"""PointNet++-style single-scale-grouping set abstraction on the device (NEW ops, SURVEY.md Appendix
B.4-B.7): farthest point sampling, ball query, grouping, shared MLP + max-pool.

Function names and argument order follow the published `pointnet2_ops` CUDA ops the north star refers
to (furthest_point_sample, ball_query, group_points / grouping_operation); the reference itself has no
such code.  All tensors are contiguous CUDA tensors; index outputs are bit-exact against the oracle.
"""
from __future__ import annotations

import torch

from ._capi import MLP_AUTO, MLP_SIMT, MLP_TCGEN05, check, lib
from .ops import _ptr, _scratch, _stream_ptr

__all__ = ["furthest_point_sample", "gather_points", "ball_query", "group_points", "shared_mlp_maxpool",
           "SetAbstraction", "MLP_AUTO", "MLP_SIMT", "MLP_TCGEN05"]


def _f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32):
        raise ValueError(f"{name} must be a float32 CUDA tensor")
    return t.contiguous()


def furthest_point_sample(xyz: torch.Tensor, npoint: int) -> torch.Tensor:
    """(B,N,3) float32 -> (B,npoint) int32 (B.4)."""
    xyz = _f32(xyz, "xyz")
    b, n, _ = xyz.shape
    out = torch.empty((b, npoint), dtype=torch.int32, device=xyz.device)
    nb = lib.lidar_fps_workspace_bytes(b, n)
    ws = _scratch.get("fps", nb, xyz.device)
    check(lib.lidar_fps(_ptr(xyz), b, n, int(npoint), _ptr(out), _ptr(ws), ws.numel(), _stream_ptr()))
    return out


def gather_points(xyz: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """new_xyz[b,m] = xyz[b, idx[b,m]] -> (B,M,3)."""
    xyz = _f32(xyz, "xyz")
    idx = idx.to(torch.int32).contiguous()
    b, n, _ = xyz.shape
    m = idx.shape[1]
    out = torch.empty((b, m, 3), dtype=torch.float32, device=xyz.device)
    check(lib.lidar_gather_points(_ptr(xyz), _ptr(idx), b, n, m, _ptr(out), _stream_ptr()))
    return out


def ball_query(xyz: torch.Tensor, new_xyz: torch.Tensor, radius: float, nsample: int) -> torch.Tensor:
    """(B,N,3),(B,M,3) -> (B,M,nsample) int32 (B.5)."""
    xyz, new_xyz = _f32(xyz, "xyz"), _f32(new_xyz, "new_xyz")
    b, n, _ = xyz.shape
    m = new_xyz.shape[1]
    out = torch.empty((b, m, nsample), dtype=torch.int32, device=xyz.device)
    check(lib.lidar_ball_query(_ptr(xyz), _ptr(new_xyz), b, n, m, float(radius), int(nsample), _ptr(out),
                               _stream_ptr()))
    return out


def group_points(xyz: torch.Tensor, features, idx: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
    """(B,3+C,M,k) float32 (B.6): centred coordinates, then the gathered feature channels."""
    xyz, new_xyz = _f32(xyz, "xyz"), _f32(new_xyz, "new_xyz")
    feats = _f32(features, "features") if features is not None else None
    idx = idx.to(torch.int32).contiguous()
    b, n, _ = xyz.shape
    _, m, k = idx.shape
    c = feats.shape[1] if feats is not None else 0
    out = torch.empty((b, 3 + c, m, k), dtype=torch.float32, device=xyz.device)
    check(lib.lidar_group_points(_ptr(xyz), _ptr(feats), _ptr(idx), _ptr(new_xyz), b, n, m, k, c, _ptr(out),
                                 _stream_ptr()))
    return out


def shared_mlp_maxpool(weights, biases, grouped: torch.Tensor | None = None, xyz=None, idx=None, new_xyz=None,
                       features=None, impl: int = MLP_AUTO) -> torch.Tensor:
    """Three 1x1-conv + bias + ReLU layers and a max over the neighbours -> (B,c3,M) float32 (B.7).

    Either pass the materialised `grouped` tensor (B,c_in,M,k), or `xyz`, `idx`, `new_xyz`
    (+ `features`) for the fused gather that never builds it."""
    if len(weights) != 3 or len(biases) != 3:
        raise ValueError("shared_mlp_maxpool takes exactly three layers")
    w = [_f32(x, "weight") for x in weights]
    bs = [_f32(x, "bias") for x in biases]
    c_in, c1, c2, c3 = w[0].shape[1], w[0].shape[0], w[1].shape[0], w[2].shape[0]
    if w[1].shape[1] != c1 or w[2].shape[1] != c2:
        raise ValueError("layer widths do not chain")
    if grouped is not None:
        grouped = _f32(grouped, "grouped")
        b, cg, m, k = grouped.shape
        if cg != c_in:
            raise ValueError("grouped channel count does not match W1")
        n = 0
        dev = grouped.device
        args = (None, None, None, None, _ptr(grouped))
    else:
        xyz, new_xyz = _f32(xyz, "xyz"), _f32(new_xyz, "new_xyz")
        idx = idx.to(torch.int32).contiguous()
        feats = _f32(features, "features") if features is not None else None
        b, n, _ = xyz.shape
        _, m, k = idx.shape
        if c_in != 3 + (feats.shape[1] if feats is not None else 0):
            raise ValueError("input channel count does not match W1")
        dev = xyz.device
        args = (_ptr(xyz), _ptr(feats), _ptr(idx), _ptr(new_xyz), None)
    out = torch.empty((b, c3, m), dtype=torch.float32, device=dev)
    check(lib.lidar_shared_mlp_maxpool(*args, b, n, m, k, c_in, c1, c2, c3, _ptr(w[0]), _ptr(bs[0]), _ptr(w[1]),
                                       _ptr(bs[1]), _ptr(w[2]), _ptr(bs[2]), _ptr(out), int(impl), _stream_ptr()))
    return out


class SetAbstraction:
    """PointNet++ SSG set-abstraction level: FPS -> ball query -> (fused) group + shared MLP + max-pool."""

    def __init__(self, npoint: int, radius: float, nsample: int, weights, biases, impl: int = MLP_AUTO):
        self.npoint, self.radius, self.nsample = int(npoint), float(radius), int(nsample)
        self.weights = [_f32(w, "weight") for w in weights]
        self.biases = [_f32(b, "bias") for b in biases]
        self.impl = impl

    def __call__(self, xyz: torch.Tensor, features=None):
        fps_idx = furthest_point_sample(xyz, self.npoint)
        new_xyz = gather_points(xyz, fps_idx)
        idx = ball_query(xyz, new_xyz, self.radius, self.nsample)
        new_features = shared_mlp_maxpool(self.weights, self.biases, xyz=xyz, idx=idx, new_xyz=new_xyz,
                                          features=features, impl=self.impl)
        return new_xyz, new_features, fps_idx, idx
