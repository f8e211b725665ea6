"""``DynamicScatter`` / ``dynamic_scatter`` with the call signatures of
mmdetection3d/mmdet3d/ops/voxel/scatter_points.py:9-107 (== mmcv.ops.DynamicScatter).
"""
import torch
from torch import nn
from torch.autograd import Function

from .voxel_layer import dynamic_point_to_voxel_backward, dynamic_point_to_voxel_forward


class _dynamic_scatter(Function):

    @staticmethod
    def forward(ctx, feats, coors, reduce_type='max', dims=None):
        """feats (N,C), coors (N,3|4) -> (voxel_feats (M,C), voxel_coors (M,3|4)).
        scatter_points.py:12-34."""
        results = dynamic_point_to_voxel_forward(feats, coors, reduce_type, dims)
        voxel_feats, voxel_coors, point2voxel_map, voxel_points_count = results
        ctx.reduce_type = reduce_type
        ctx.save_for_backward(feats, voxel_feats, point2voxel_map, voxel_points_count)
        ctx.mark_non_differentiable(voxel_coors)
        return voxel_feats, voxel_coors

    @staticmethod
    def backward(ctx, grad_voxel_feats, grad_voxel_coors=None):
        feats, voxel_feats, point2voxel_map, voxel_points_count = ctx.saved_tensors
        grad_feats = torch.empty_like(feats)     # fully written by the kernel
        dynamic_point_to_voxel_backward(grad_feats, grad_voxel_feats.contiguous(), feats,
                                        voxel_feats.contiguous(), point2voxel_map,
                                        voxel_points_count.contiguous(), ctx.reduce_type)
        return grad_feats, None, None, None


dynamic_scatter = _dynamic_scatter.apply


class DynamicScatter(nn.Module):
    """scatter_points.py:53-107.  ``average_points``: mean if True else max."""

    def __init__(self, voxel_size, point_cloud_range, average_points: bool):
        super(DynamicScatter, self).__init__()
        self.voxel_size = voxel_size
        self.point_cloud_range = point_cloud_range
        self.average_points = average_points
        # (gz, gy, gx): coordinates produced by dynamic voxelization with the same
        # parameters are below this bound, which lets the kernel skip measuring the extent.
        pcr = torch.tensor(point_cloud_range, dtype=torch.float32)
        vs = torch.tensor(voxel_size, dtype=torch.float32)
        g = torch.round((pcr[3:] - pcr[:3]) / vs).long().tolist()
        self._dims = [max(g[2], 1), max(g[1], 1), max(g[0], 1)]
        self._batch_hint = 1          # samples the bitmap is sized for; grows to the largest batch seen

    def forward_single(self, points, coors):
        reduce = 'mean' if self.average_points else 'max'
        return dynamic_scatter(points.contiguous(), coors.contiguous(), reduce, self._dims)

    def forward(self, points, coors):
        """scatter_points.py:74-99.  Batched ``coors`` (N,4) = (batch,z,y,x): the reference loops over the samples
        with a ``torch.where`` and two host synchronisations each; here the batch column is the slowest dimension of
        the voxel key, so the whole batch is ONE launch sequence whose output is already the reference's
        concatenation in sample order.  The sample count is a remembered hint (checked on the device: a batch
        that outgrows it is measured and redone), so steady state costs the single host read of M."""
        if coors.size(-1) == 3:
            return self.forward_single(points, coors)
        reduce = 'mean' if self.average_points else 'max'
        dims = [self._batch_hint] + self._dims
        voxels, voxel_coors = dynamic_scatter(points.contiguous(), coors.contiguous(), reduce, dims)
        self._batch_hint = max(self._batch_hint, dims[0])        # updated in place when the hint was exceeded
        return voxels, voxel_coors

    def __repr__(self):
        return (self.__class__.__name__ + '(voxel_size=' + str(self.voxel_size) +
                ', point_cloud_range=' + str(self.point_cloud_range) +
                ', average_points=' + str(self.average_points) + ')')
