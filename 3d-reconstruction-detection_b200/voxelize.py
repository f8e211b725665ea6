"""``Voxelization`` / ``voxelization`` with the call signatures of
mmdetection3d/mmdet3d/ops/voxel/voxelize.py:10-148 (== mmcv.ops.Voxelization),
so the ``pts_voxel_layer`` config dicts of the reference build it unchanged.
"""
import torch
from torch import nn
from torch.autograd import Function
from torch.nn.modules.utils import _pair

from .voxel_layer import dynamic_voxelize, hard_voxelize


class _Voxelization(Function):

    @staticmethod
    def forward(ctx, points, voxel_size, coors_range, max_points=35, max_voxels=20000,
                deterministic=True):
        """points (N, >=3) -> (voxels (M,max_points,C), coors (M,3) int32 zyx,
        num_points_per_voxel (M) int32); or coors (N,3) when max_points == -1 or
        max_voxels == -1 (dynamic voxelization).  voxelize.py:52-70."""
        if max_points == -1 or max_voxels == -1:
            coors = points.new_zeros(size=(points.size(0), 3), dtype=torch.int)
            dynamic_voxelize(points, coors, voxel_size, coors_range, 3)
            return coors
        # rows >= voxel_num are sliced away below, and the kernel writes every
        # element of rows < voxel_num (zeros included): no zero-fill pass needed.
        voxels = points.new_empty(size=(max_voxels, max_points, points.size(1)))
        coors = points.new_empty(size=(max_voxels, 3), dtype=torch.int)
        num_points_per_voxel = points.new_empty(size=(max_voxels,), dtype=torch.int)
        voxel_num = hard_voxelize(points, voxels, coors, num_points_per_voxel, voxel_size,
                                  coors_range, max_points, max_voxels, 3, deterministic)
        return voxels[:voxel_num], coors[:voxel_num], num_points_per_voxel[:voxel_num]


voxelization = _Voxelization.apply


class Voxelization(nn.Module):
    """voxelize.py:76-148.  ``max_voxels`` is (training, testing)."""

    def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels=20000,
                 deterministic=True):
        super(Voxelization, self).__init__()
        self.voxel_size = voxel_size
        self.point_cloud_range = point_cloud_range
        self.max_num_points = max_num_points
        self.max_voxels = max_voxels if isinstance(max_voxels, tuple) else _pair(max_voxels)
        self.deterministic = deterministic

        pcr = torch.tensor(point_cloud_range, dtype=torch.float32)
        vs = torch.tensor(voxel_size, dtype=torch.float32)
        grid_size = torch.round((pcr[3:] - pcr[:3]) / vs).long()      # voxelize.py:113-118
        self.grid_size = grid_size
        self.pcd_shape = [*grid_size[:2], 1][::-1]

    def forward(self, input):
        max_voxels = self.max_voxels[0] if self.training else self.max_voxels[1]
        return voxelization(input, self.voxel_size, self.point_cloud_range, self.max_num_points,
                            max_voxels, self.deterministic)

    def __repr__(self):
        return (self.__class__.__name__ + '(voxel_size=' + str(self.voxel_size) +
                ', point_cloud_range=' + str(self.point_cloud_range) +
                ', max_num_points=' + str(self.max_num_points) +
                ', max_voxels=' + str(self.max_voxels) +
                ', deterministic=' + str(self.deterministic) + ')')
