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
from .backproject import _prep, make_params, squeeze_head


class DepthToVoxels(nn.Module):
    """Batched fused front end.  Like the reference ops every call returns FRESH output tensors (torch's
    caching allocator makes that cheap: no cudaMalloc in steady state, no host synchronisation).
    ``reuse_buffers=True`` is the opt-in for callers that consume a result before the next call with the
    same shape on the same stream (benchmarks, CUDA-graph capture): the outputs then live in buffers
    allocated once per (shape, stream) and are overwritten by the next call."""

    def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels,
                 max_depth=None, range_filter=None, with_mean=True, with_voxels=True, reuse_buffers=False,
                 flat_outputs=False):
        """``with_voxels=False`` skips the padded (B, max_voxels, max_points, 3) tensor -- 85 % of
        the output bytes -- for callers that only feed the sparse encoder (mean + coors + num)."""
        super().__init__()
        if not with_voxels and not with_mean:
            raise ValueError("with_voxels=False needs with_mean=True")
        self.with_voxels = with_voxels
        self.voxel_size = list(voxel_size)
        self.point_cloud_range = list(point_cloud_range)
        self.max_num_points = int(max_num_points)
        self.max_voxels = max_voxels if isinstance(max_voxels, tuple) else (max_voxels, max_voxels)
        self.max_depth = max_depth
        self.range_filter = range_filter
        self.with_mean = with_mean
        self.reuse_buffers = reuse_buffers
        self.flat_outputs = bool(flat_outputs)      # mean / coors / num / voxel_num as views of one buffer (result["flat"])
        if self.flat_outputs and not with_mean:
            raise ValueError("flat_outputs=True needs with_mean=True")
        self._out_cache = {}

    def _alloc(self, B, dev, max_voxels):
        K = self.max_num_points
        if self.flat_outputs:
            # what the sparse encoder consumes in ONE allocation (all 4-byte types): [mean | coors | num | voxel_num].
            # A rank that gathers its shard's results moves them with one collective instead of four.
            n = B * max_voxels
            flat = torch.empty((n * 3 + n * 3 + n + B,), dtype=torch.int32, device=dev)
            return dict(
                voxels=(torch.empty((B, max_voxels, K, 3), dtype=torch.float32, device=dev)
                        if self.with_voxels else None),
                mean=flat[:n * 3].view(torch.float32).view(B, max_voxels, 3),
                coors=flat[n * 3:n * 6].view(B, max_voxels, 3),
                num=flat[n * 6:n * 7].view(B, max_voxels),
                voxel_num=flat[n * 7:n * 7 + B],
                flat=flat)
        return dict(
            voxels=(torch.empty((B, max_voxels, K, 3), dtype=torch.float32, device=dev)
                    if self.with_voxels else None),
            coors=torch.empty((B, max_voxels, 3), dtype=torch.int32, device=dev),
            num=torch.empty((B, max_voxels), dtype=torch.int32, device=dev),
            mean=torch.empty((B, max_voxels, 3), dtype=torch.float32, device=dev) if self.with_mean else None,
            voxel_num=torch.empty((B,), dtype=torch.int32, device=dev))

    def _get_buffers(self, B, dev, max_voxels):
        if not self.reuse_buffers:
            return self._alloc(B, dev, max_voxels)
        key = (B, dev, max_voxels, torch.cuda.current_stream(dev).cuda_stream)
        if key not in self._out_cache:
            self._out_cache[key] = self._alloc(B, dev, max_voxels)
        return self._out_cache[key]

    def forward(self, depths, intrinsics, cam2lidar_rts, confs=None, conf_thresh=None, sky_masks=None):
        """depths (B,N,H,W[,1]); confs / sky_masks of the same shape -- sky_masks either a bool / uint8 mask or DA3's
        raw fp32 sky output (sky iff >= 0.5, thresholded inside the kernels) -> dict(voxels (B,MV,K,3), coors (B,MV,3), num_points (B,MV),
        voxel_mean (B,MV,3) | None, voxel_num (B,)): rows [0, voxel_num[b]) of sample b
        are valid and equal to the reference's outputs; the rest is undefined."""
        depths = squeeze_head(depths).contiguous()
        K, M, conf, sky, sky_prob = _prep(depths, intrinsics, cam2lidar_rts, confs, sky_masks, conf_thresh)
        B, N, H, W = depths.shape
        max_voxels = self.max_voxels[0] if self.training else self.max_voxels[1]
        p = make_params(B, N, H, W, self.max_depth, conf_thresh if conf is not None else None,
                        self.range_filter, sky_prob)
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
        res = dict(voxels=out["voxels"], coors=out["coors"], num_points=out["num"],
                   voxel_mean=out["mean"], voxel_num=out["voxel_num"])
        if "flat" in out:
            res["flat"] = out["flat"]
        return res

    @staticmethod
    def to_sparse_encoder_inputs(result, batch_offset=0, with_num_points=False):
        """(voxel_features (sum M, 3), coors (sum M, 4) [b,z,y,x], batch_size): what
        SparseEncoder.forward consumes (sparse_encoder.py:96-128)."""
        return pack_sparse_inputs(result, batch_offset=batch_offset, with_num_points=with_num_points)


_pack_cache = {}


def pack_sparse_inputs(result, batch_offset=0, with_num_points=False, sync=True, reuse_buffers=False):
    """One launch instead of the per-sample slice + F.pad + torch.cat tail of
    ``_voxelize_and_encode`` (sparse_refinement.py:393-402).

    ``result``: output dict of :class:`DepthToVoxels` (voxel_mean, coors, num_points, voxel_num).
    Returns ``(voxel_features (sum M, F), coors (sum M, 4) [b,z,y,x], batch_size)`` -- plus
    ``num_points (sum M)`` after the coors when ``with_num_points`` -- like the reference's
    ``(voxel_features, num_points, coors)``.  ``sync=True`` reads sum M back (the one D2H read
    the reference's ``hard_voxelize`` return value forces per sample); ``sync=False`` returns the
    worst-case-sized buffers and a device tensor ``offsets (B+1)`` instead of the batch size.
    Fresh tensors per call unless ``reuse_buffers=True`` (then: views of buffers that the next call with the
    same shape on the same stream overwrites)."""
    feats, coors, num, vnum = result["voxel_mean"], result["coors"], result["num_points"], result["voxel_num"]
    if feats is None:
        raise RuntimeError("pack_sparse_inputs needs voxel_mean (DepthToVoxels(with_mean=True))")
    for t, n, d in ((feats, "voxel_mean", torch.float32), (coors, "coors", torch.int32),
                    (num, "num_points", torch.int32), (vnum, "voxel_num", torch.int32)):
        _lib.require_cuda(t, n, d)
    B, MV, F = feats.shape
    dev = feats.device
    def alloc():
        return (torch.empty((B * MV, F), dtype=torch.float32, device=dev),
                torch.empty((B * MV, 4), dtype=torch.int32, device=dev),
                torch.empty((B * MV,), dtype=torch.int32, device=dev) if with_num_points else None,
                torch.empty((B + 1,), dtype=torch.int32, device=dev))
    if reuse_buffers:
        key = (B, MV, F, dev, with_num_points, torch.cuda.current_stream(dev).cuda_stream)
        if key not in _pack_cache:
            _pack_cache[key] = alloc()
        of, oc, on, offs = _pack_cache[key]
    else:
        of, oc, on, offs = alloc()
    with torch.cuda.device_of(feats):
        st = _lib.lib().rd3_pack_sparse_inputs(_lib.ptr(feats), _lib.ptr(coors), _lib.ptr(num), _lib.ptr(vnum),
                                               B, MV, F, int(batch_offset), _lib.ptr(of), _lib.ptr(oc),
                                               _lib.ptr(on), _lib.ptr(offs), _lib.stream_of(feats))
        _lib.check(st, "pack_sparse_inputs")
    if not sync:
        return (of, oc, on, offs) if with_num_points else (of, oc, offs)
    total = int(offs[B].item())
    if with_num_points:
        return of[:total], oc[:total], on[:total], B
    return of[:total], oc[:total], B
