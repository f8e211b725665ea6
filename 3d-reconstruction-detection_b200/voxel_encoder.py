"""``HardSimpleVFE`` with the interface of
mmdetection3d/mmdet3d/models/voxel_encoders/voxel_encoder.py:15-47."""
import torch
from torch import nn

from . import _lib


def hard_simple_vfe(features, num_points, num_features):
    """features (M,K,C) fp32, num_points (M) int -> (M,F) = sum over slots / count."""
    _lib.require_cuda(features, "features", torch.float32)
    if num_points.dtype != torch.int32:
        num_points = num_points.to(torch.int32)
    _lib.require_cuda(num_points, "num_points", torch.int32)
    M, K, C = features.shape
    F = min(int(num_features), C)              # python slicing [:F] clamps the same way
    out = torch.empty((M, F), dtype=torch.float32, device=features.device)
    with torch.cuda.device_of(features):
        st = _lib.lib().rd3_hard_simple_vfe(_lib.ptr(features), _lib.ptr(num_points), M, K, C, F,
                                            _lib.ptr(out), _lib.stream_of(features))
        _lib.check(st, "hard_simple_vfe")
    return out


class HardSimpleVFE(nn.Module):
    """Simple voxel feature encoder used in SECOND: the mean of the points of a voxel.

    Forward-only kernel (the reference's module is used under no_grad for the
    pseudo-point branch and has no parameters)."""

    def __init__(self, num_features=4):
        super(HardSimpleVFE, self).__init__()
        self.num_features = num_features
        self.fp16_enabled = False

    def forward(self, features, num_points, coors=None):
        out_half = features.dtype == torch.half
        out = hard_simple_vfe(features.float().contiguous(), num_points.contiguous(), self.num_features)
        return out.half() if out_half else out
