"""Gather side of the pillar encoders, same call interface as the reference:

    PillarDecorator        the decorations of PillarFeatureNet.forward in front of its PFN layers
                           (mmdet3d/models/voxel_encoders/pillar_encoder.py:104-146)
    PointPillarsScatter    mmdet3d/models/middle_encoders/pillar_scatter.py:9-102

The learned PFN layers (Linear + BN + max) stay what they are in the reference; they consume
``PillarDecorator``'s output unchanged.
"""
import torch
from torch import nn

from . import _lib


def pillar_decorate(features, num_points, coors, voxel_size=(0.2, 0.2, 4),
                    point_cloud_range=(0, -40, -3, 70.4, 40, 1), with_cluster_center=True,
                    with_voxel_center=True, with_distance=False, legacy=True):
    """features (M, K, C) fp32, num_points (M) int32, coors (M, 4) [b,z,y,x] (or (M,3) zyx) ->
    (M, K, C + 3 + 2 [+ 1]) masked decorated features (pillar_encoder.py:104-143)."""
    _lib.require_cuda(features, "features", torch.float32)
    if features.dim() != 3 or features.shape[2] < 3:
        raise RuntimeError("features must be (M, max_points, C>=3)")
    num_points = num_points.to(torch.int32).contiguous()
    coors = coors.to(torch.int32).contiguous()
    _lib.require_cuda(num_points, "num_points", torch.int32)
    _lib.require_cuda(coors, "coors", torch.int32)
    M, K, C = features.shape
    if coors.dim() != 2 or coors.shape[0] != M or coors.shape[1] not in (3, 4) or num_points.shape[0] != M:
        raise RuntimeError("coors must be (M, 3|4) and num_points (M)")
    vx, vy = float(voxel_size[0]), float(voxel_size[1])
    x_off = vx / 2 + float(point_cloud_range[0])          # :87-88, python floats
    y_off = vy / 2 + float(point_cloud_range[1])
    cout = C + 3 * bool(with_cluster_center) + 2 * bool(with_voxel_center) + bool(with_distance)
    out = torch.empty((M, K, cout), dtype=torch.float32, device=features.device)
    with torch.cuda.device_of(features):
        st = _lib.lib().rd3_pillar_decorate(_lib.ptr(features), _lib.ptr(num_points), _lib.ptr(coors), M, K, C,
                                            coors.shape[1], int(bool(with_cluster_center)),
                                            int(bool(with_voxel_center)), int(bool(with_distance)),
                                            int(bool(legacy)), vx, vy, x_off, y_off, _lib.ptr(out),
                                            _lib.stream_of(features))
        _lib.check(st, "pillar_decorate")
    return out


class PillarDecorator(nn.Module):
    """Constructor arguments of PillarFeatureNet that shape the decorations (pillar_encoder.py:41-91);
    ``out_channels`` is the ``in_channels`` its first PFN layer sees."""

    def __init__(self, in_channels=4, with_distance=False, with_cluster_center=True, with_voxel_center=True,
                 voxel_size=(0.2, 0.2, 4), point_cloud_range=(0, -40, -3, 70.4, 40, 1), legacy=True):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = in_channels + 3 * bool(with_cluster_center) + 2 * bool(with_voxel_center) + \
            bool(with_distance)
        self._with_distance = with_distance
        self._with_cluster_center = with_cluster_center
        self._with_voxel_center = with_voxel_center
        self.voxel_size = voxel_size
        self.point_cloud_range = point_cloud_range
        self.legacy = legacy

    def forward(self, features, num_points, coors):
        return pillar_decorate(features, num_points, coors, self.voxel_size, self.point_cloud_range,
                               self._with_cluster_center, self._with_voxel_center, self._with_distance,
                               self.legacy)


class PointPillarsScatter(nn.Module):
    """pillar_scatter.py:9-102: (M, C) pillar features -> (B, C, ny, nx) pseudo image."""

    def __init__(self, in_channels, output_shape):
        super().__init__()
        self.output_shape = output_shape
        self.ny = output_shape[0]
        self.nx = output_shape[1]
        self.in_channels = in_channels
        self.fp16_enabled = False

    def forward(self, voxel_features, coors, batch_size=None):
        if batch_size is not None:
            return self.forward_batch(voxel_features, coors, batch_size)
        return self.forward_single(voxel_features, coors)

    def _scatter(self, voxel_features, coors, batch_size, single):
        voxel_features = voxel_features.float().contiguous()
        _lib.require_cuda(voxel_features, "voxel_features", torch.float32)
        coors = coors.to(torch.int32).contiguous()
        _lib.require_cuda(coors, "coors", torch.int32)
        M, C = voxel_features.shape
        if C != self.in_channels or coors.shape[0] != M or coors.dim() != 2 or coors.shape[1] not in (3, 4):
            raise RuntimeError("voxel_features must be (M, in_channels) and coors (M, 3|4)")
        if single and coors.shape[1] == 4:
            # forward_single indexes coors[:, 1] * nx + coors[:, 2] (:53): columns (?, y, x)
            coors = coors[:, :3].contiguous()
        canvas = torch.empty((batch_size, C, self.ny, self.nx), dtype=torch.float32, device=voxel_features.device)
        with torch.cuda.device_of(voxel_features):
            st = _lib.lib().rd3_pillars_scatter(_lib.ptr(voxel_features), _lib.ptr(coors), M, C, coors.shape[1],
                                                batch_size, self.ny, self.nx, _lib.ptr(canvas),
                                                _lib.stream_of(voxel_features))
            _lib.check(st, "pillars_scatter")
        return canvas

    def forward_single(self, voxel_features, coors):
        return [self._scatter(voxel_features, coors, 1, True)]

    def forward_batch(self, voxel_features, coors, batch_size):
        return self._scatter(voxel_features, coors, batch_size, False)


def map_voxel_center_to_point(pts_coors, voxel_mean, voxel_coors, return_index=False):
    """DynamicVFE.map_voxel_center_to_point (voxel_encoder.py:179-219; same code in HardVFE's
    dynamic branch and pillar_encoder.py:235-275): the feature row of each point's voxel.

    pts_coors (N,4) and voxel_coors (M,4) [b,z,y,x]; voxel_mean (M,C) -> (N,C).  No dense
    z*y*x*batch canvas is built; a point whose voxel is missing gets row 0 like the reference."""
    pts_coors = pts_coors.to(torch.int32).contiguous()
    voxel_coors = voxel_coors.to(torch.int32).contiguous()
    voxel_mean = voxel_mean.float().contiguous()
    for t, n in ((pts_coors, "pts_coors"), (voxel_coors, "voxel_coors"), (voxel_mean, "voxel_mean")):
        _lib.require_cuda(t, n)
    if pts_coors.dim() != 2 or pts_coors.shape[1] != 4 or voxel_coors.dim() != 2 or voxel_coors.shape[1] != 4:
        raise RuntimeError("pts_coors / voxel_coors must be (N,4) / (M,4) [b,z,y,x]")
    N, M, C = pts_coors.shape[0], voxel_coors.shape[0], voxel_mean.shape[1]
    if voxel_mean.shape[0] != M or (M == 0 and N > 0):
        raise RuntimeError("voxel_mean must be (M,C) with M = voxel_coors rows > 0")
    out = torch.empty((N, C), dtype=torch.float32, device=voxel_mean.device)
    index = torch.empty((N,), dtype=torch.int32, device=voxel_mean.device) if return_index else None
    L = _lib.lib()
    with torch.cuda.device_of(voxel_mean):
        ws = _lib.workspace(voxel_mean.device, L.rd3_map_voxel_to_point_workspace_bytes(M))
        st = L.rd3_map_voxel_to_point(_lib.ptr(pts_coors), N, _lib.ptr(voxel_coors), _lib.ptr(voxel_mean), M, C,
                                      _lib.ptr(out), _lib.ptr(index), _lib.ptr(ws), ws.numel(),
                                      _lib.stream_of(voxel_mean))
        _lib.check(st, "map_voxel_to_point")
    return (out, index) if return_index else out
