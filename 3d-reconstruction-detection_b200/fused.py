"""Fused depth -> hard voxels (+ HardSimpleVFE mean) for a batch of frames.

Replaces, in one launch sequence and without materialising the point cloud,
the chain  _backproject_depth_to_points (reconstruction_backbone.py:285-386)
-> [FilterPointByRange, respoint_post_processing.py:170-205]
-> per-sample Voxelization loop (sparse_refinement.py:382-391)
-> HardSimpleVFE (voxel_encoder.py:45-46).
"""
import torch
from torch import nn

from . import _lib
from .backproject import _prep, make_params


class DepthToVoxels(nn.Module):
    """Batched fused front end.  Output buffers and scratch are allocated once
    per input shape and reused, so steady-state calls allocate nothing and do
    not synchronise with the host."""

    def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels,
                 max_depth=None, range_filter=None, with_mean=True):
        super().__init__()
        self.voxel_size = list(voxel_size)
        self.point_cloud_range = list(point_cloud_range)
        self.max_num_points = int(max_num_points)
        self.max_voxels = max_voxels if isinstance(max_voxels, tuple) else (max_voxels, max_voxels)
        self.max_depth = max_depth
        self.range_filter = range_filter
        self.with_mean = with_mean
        self._out_cache = {}

    def _get_buffers(self, B, dev, max_voxels):
        key = (B, dev, max_voxels)
        if key not in self._out_cache:
            K = self.max_num_points
            self._out_cache[key] = dict(
                voxels=torch.empty((B, max_voxels, K, 3), dtype=torch.float32, device=dev),
                coors=torch.empty((B, max_voxels, 3), dtype=torch.int32, device=dev),
                num=torch.empty((B, max_voxels), dtype=torch.int32, device=dev),
                mean=torch.empty((B, max_voxels, 3), dtype=torch.float32, device=dev) if self.with_mean else None,
                voxel_num=torch.empty((B,), dtype=torch.int32, device=dev))
        return self._out_cache[key]

    def forward(self, depths, intrinsics, cam2lidar_rts, confs=None, conf_thresh=None, sky_masks=None):
        """depths (B,N,H,W) -> dict(voxels (B,MV,K,3), coors (B,MV,3), num_points (B,MV),
        voxel_mean (B,MV,3) | None, voxel_num (B,)): rows [0, voxel_num[b]) of sample b
        are valid and equal to the reference's outputs; the rest is undefined."""
        depths = depths.contiguous()
        K, M, conf, sky = _prep(depths, intrinsics, cam2lidar_rts, confs, sky_masks, conf_thresh)
        B, N, H, W = depths.shape
        max_voxels = self.max_voxels[0] if self.training else self.max_voxels[1]
        p = make_params(B, N, H, W, self.max_depth, conf_thresh if conf is not None else None,
                        self.range_filter)
        out = self._get_buffers(B, depths.device, max_voxels)
        L = _lib.lib()
        with torch.cuda.device_of(depths):
            nbytes = L.rd3_depth_to_voxels_workspace_bytes(p, self.max_num_points, max_voxels)
            ws = _lib.workspace(depths.device, nbytes)
            st = L.rd3_depth_to_voxels(_lib.ptr(depths), _lib.ptr(K), _lib.ptr(M), _lib.ptr(conf),
                                       _lib.ptr(sky), p, _lib.f3(self.voxel_size),
                                       _lib.f6(self.point_cloud_range), self.max_num_points, max_voxels,
                                       _lib.ptr(out["voxels"]), _lib.ptr(out["coors"]),
                                       _lib.ptr(out["num"]), _lib.ptr(out["mean"]),
                                       _lib.ptr(out["voxel_num"]), _lib.ptr(ws), ws.numel(),
                                       _lib.stream_of(depths))
            _lib.check(st, "depth_to_voxels")
        return dict(voxels=out["voxels"], coors=out["coors"], num_points=out["num"],
                    voxel_mean=out["mean"], voxel_num=out["voxel_num"])

    @staticmethod
    def to_sparse_encoder_inputs(result):
        """(voxel_features (sum M, 3), coors (sum M, 4) [b,z,y,x], batch_size): what
        SparseEncoder.forward consumes (sparse_encoder.py:96-128); mirrors the cat + F.pad
        of sparse_refinement.py:393-402.  One D2H read of the per-sample voxel counts."""
        n = result["voxel_num"].tolist()
        feats, coors = [], []
        for b, m in enumerate(n):
            feats.append(result["voxel_mean"][b, :m])
            coors.append(nn.functional.pad(result["coors"][b, :m], (1, 0), mode="constant", value=b))
        return torch.cat(feats, dim=0), torch.cat(coors, dim=0), len(n)
