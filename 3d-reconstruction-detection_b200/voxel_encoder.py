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


class _HardSimpleVFEFn(torch.autograd.Function):
    """forward = the sm_100a kernel; backward = d(sum_k x[m,k,f] / n[m]) = grad[m,f] / n[m] for
    every slot k and f < F (what autograd derives for voxel_encoder.py:45-46)."""

    @staticmethod
    def forward(ctx, features, num_points, num_features):
        ctx.save_for_backward(num_points)
        ctx.shape = features.shape
        return hard_simple_vfe(features, num_points, num_features)

    @staticmethod
    def backward(ctx, grad_out):
        (num_points,) = ctx.saved_tensors
        M, K, C = ctx.shape
        F = grad_out.shape[1]
        g = grad_out / num_points.to(grad_out.dtype).view(-1, 1)
        grad = grad_out.new_zeros((M, K, C))
        grad[:, :, :F] = g.unsqueeze(1)
        return grad, None, None


class HardSimpleVFE(nn.Module):
    """Simple voxel feature encoder used in SECOND: the mean of the points of a voxel."""

    def __init__(self, num_features=4):
        super(HardSimpleVFE, self).__init__()
        self.num_features = num_features
        self.fp16_enabled = False

    def forward(self, features, num_points, coors=None):
        out_half = features.dtype == torch.half
        feats = features.float().contiguous()
        if feats.requires_grad and torch.is_grad_enabled():
            out = _HardSimpleVFEFn.apply(feats, num_points.contiguous(), self.num_features)
        else:
            out = hard_simple_vfe(feats, num_points.contiguous(), self.num_features)
        return out.half() if out_half else out


class DynamicSimpleVFE(nn.Module):
    """mmdet3d/models/voxel_encoders/voxel_encoder.py:50-90: the mean of the points of every (dynamic) voxel,
    i.e. ``DynamicScatter(voxel_size, point_cloud_range, True)`` under ``torch.no_grad`` and ``force_fp32``.
    ``forward(features (N, C), coors (N, 3|4)) -> (voxel_features (M, C), voxel_coors (M, 3|4))``."""

    def __init__(self, voxel_size=(0.2, 0.2, 4), point_cloud_range=(0, -40, -3, 70.4, 40, 1)):
        super(DynamicSimpleVFE, self).__init__()
        from .scatter_points import DynamicScatter
        self.scatter = DynamicScatter(voxel_size, point_cloud_range, True)
        self.fp16_enabled = False

    @torch.no_grad()
    def forward(self, features, coors):
        out_half = features.dtype == torch.half
        feats, feats_coors = self.scatter(features.float().contiguous(), coors)
        return (feats.half() if out_half else feats), feats_coors


def voxel_occupancy(features, num_points, coors=None, hard=False, lambda_n=0.3, gamma_var=5.0, eps=1e-6,
                    dense_shape=None, batch_size=None):
    """Occupancy value per voxel (M, 1) and, with ``dense_shape=(Z, Y, X)``, the dense map
    (batch_size, Z, Y, X) of sparse_refinement.py:572-587 in the same launch."""
    _lib.require_cuda(features, "features", torch.float32)
    num_points = num_points.to(torch.int32).contiguous()
    _lib.require_cuda(num_points, "num_points", torch.int32)
    M, K, C = features.shape
    occ = torch.empty((M, 1), dtype=torch.float32, device=features.device)
    dense, cc, cols, B, Z, Y, X = None, None, 4, 1, 1, 1, 1
    if dense_shape is not None:
        cc = coors.to(torch.int32).contiguous()
        _lib.require_cuda(cc, "coors", torch.int32)
        cols = cc.shape[1]
        Z, Y, X = (int(v) for v in dense_shape)
        B = int(batch_size) if batch_size is not None else 1
        dense = torch.empty((B, Z, Y, X), dtype=torch.float32, device=features.device)
    with torch.cuda.device_of(features):
        st = _lib.lib().rd3_voxel_occupancy(_lib.ptr(features), _lib.ptr(num_points), M, K, C, int(bool(hard)),
                                            float(lambda_n), float(gamma_var), float(eps), _lib.ptr(occ),
                                            _lib.ptr(cc), cols, B, Z, Y, X, _lib.ptr(dense),
                                            _lib.stream_of(features))
        _lib.check(st, "voxel_occupancy")
    return occ if dense is None else (occ, dense)


class HardVoxelOccupancyVFE(nn.Module):
    """voxel_occupancy_encoder.py:12-37: 1 for a non-empty voxel."""

    def __init__(self):
        super().__init__()
        self.fp16_enabled = False

    def forward(self, features, num_points, coors):
        return voxel_occupancy(features, num_points, hard=True)


class SoftVoxelOccupancyVFE(nn.Module):
    """voxel_occupancy_encoder.py:40-99: p_occ = 1 - exp(-lambda*n - gamma*var)."""

    def __init__(self, lambda_n=0.3, gamma_var=5.0, eps=1e-6):
        super().__init__()
        self.lambda_n = lambda_n
        self.gamma_var = gamma_var
        self.eps = eps
        self.fp16_enabled = False

    def forward(self, features, num_points, coors):
        return voxel_occupancy(features, num_points, lambda_n=self.lambda_n, gamma_var=self.gamma_var,
                               eps=self.eps)
