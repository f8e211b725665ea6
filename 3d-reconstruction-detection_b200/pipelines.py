"""Post-processing pipeline steps of the plugin that sit on the hot path, with the
call interface of projects/mmdet3d_plugin/datasets/pipelines/respoint_post_processing.py:
``__call__(data: dict) -> dict`` with keys 'points', 'colors', 'indices'.
"""
import torch

from .voxel_layer import hard_voxelize


class FilterPointByRange:
    """respoint_post_processing.py:170-205: keep points with min <= p <= max (inclusive)."""

    def __init__(self, point_cloud_range=None):
        self.point_cloud_range = point_cloud_range

    def __call__(self, data):
        if self.point_cloud_range is None:
            return data
        pts = data['points']
        colors = data.get('colors')
        x0, y0, z0, x1, y1, z1 = self.point_cloud_range
        mask = ((pts[:, 0] >= x0) & (pts[:, 0] <= x1) & (pts[:, 1] >= y0) & (pts[:, 1] <= y1) &
                (pts[:, 2] >= z0) & (pts[:, 2] <= z1))
        return {'points': pts[mask], 'colors': colors[mask] if colors is not None else None,
                'indices': torch.nonzero(mask, as_tuple=False).squeeze(1)}


class VoxelDownsample:
    """respoint_post_processing.py:18-98: one centroid per voxel (first 100 points of each of
    at most 200 000 voxels).  The reference's Python loop over voxels (:69-75) is replaced by
    the voxel mean fused into the hard-voxelization kernels; the colour of a centroid is that
    of the nearest input point (:90-94), searched in chunks instead of one (M,N) matrix."""

    MAX_POINTS = 100
    MAX_VOXELS = 200000

    def __init__(self, voxel_size=None, point_cloud_range=None, device=None):
        self.voxel_size = voxel_size
        self.point_cloud_range = point_cloud_range
        self.device = device

    def __call__(self, data):
        if self.voxel_size is None:
            return data
        points = data['points']
        colors = data.get('colors')
        if not torch.is_tensor(points):
            points = torch.as_tensor(points, device=self.device, dtype=torch.float32)
        if not torch.is_floating_point(points):
            points = points.float()
        dev = points.device
        if self.point_cloud_range is None:
            pcr = (points.min(dim=0).values - 1.0).tolist() + (points.max(dim=0).values + 1.0).tolist()
        else:
            pcr = self.point_cloud_range
        vs = self.voxel_size
        if isinstance(vs, (int, float)):
            vs = [vs, vs, vs]
        pts = points.contiguous()
        N, C = pts.shape
        mv, mp = self.MAX_VOXELS, self.MAX_POINTS
        voxels = torch.empty((mv, mp, C), dtype=torch.float32, device=dev)
        coors = torch.empty((mv, 3), dtype=torch.int32, device=dev)
        num = torch.empty((mv,), dtype=torch.int32, device=dev)
        mean = torch.empty((mv, C), dtype=torch.float32, device=dev)
        n = hard_voxelize(pts, voxels, coors, num, vs, pcr, mp, mv, 3, True, voxel_mean=mean)
        if n == 0:
            return {'points': points, 'colors': colors, 'indices': torch.arange(N, device=dev)}
        centers = mean[:n]
        idx = torch.arange(n, device=dev)
        vcol = None
        if colors is not None:
            idx = torch.empty(n, dtype=torch.long, device=dev)
            step = max(1, (1 << 26) // max(N, 1))
            for s in range(0, n, step):
                idx[s:s + step] = torch.cdist(centers[s:s + step], pts,
                                               compute_mode='donot_use_mm_for_euclid_dist').argmin(dim=1)
            vcol = colors[idx]
        return {'points': centers, 'colors': vcol, 'indices': idx}
