"""Drop-in for the reference's pybind module ``voxel_layer``
(mmdetection3d/mmdet3d/ops/voxel/src/voxelization.cpp:6-11): the same four
names, argument order and error behaviour (RuntimeError), backed by the sm_100a
kernels in librd3_b200.so.  CUDA tensors only -- there is no CPU dispatch.
"""
import torch

from . import _lib

_REDUCE = {"sum": 0, "mean": 1, "max": 2}


def _reduce_code(reduce_type):
    if reduce_type not in _REDUCE:          # voxelization.h:97-106
        raise RuntimeError("do not support reduce type " + str(reduce_type))
    return _REDUCE[reduce_type]


def hard_voxelize(points, voxels, coors, num_points_per_voxel, voxel_size, coors_range,
                  max_points, max_voxels, NDim=3, deterministic=True, voxel_mean=None,
                  point2voxel=None):
    """voxelization.h:58-81.  Fills the caller's zero-initialised ``voxels``
    (max_voxels,max_points,C), ``coors`` (max_voxels,3) and
    ``num_points_per_voxel`` (max_voxels) in place and returns the voxel count as
    a Python int (one device->host read, like the reference).

    ``deterministic`` is accepted for signature compatibility; the result is
    always the deterministic one (bit-identical to hard_voxelize_cpu).
    Extra keyword outputs: ``voxel_mean`` (max_voxels,F) fused HardSimpleVFE,
    ``point2voxel`` (N,) int32.
    """
    if NDim != 3:
        raise RuntimeError("rd3_b200.hard_voxelize supports NDim == 3 only")
    _lib.require_cuda(points, "points", torch.float32)
    _lib.require_cuda(voxels, "voxels", torch.float32)
    _lib.require_cuda(coors, "coors", torch.int32)
    _lib.require_cuda(num_points_per_voxel, "num_points_per_voxel", torch.int32)
    N, C = points.shape
    if tuple(voxels.shape) != (max_voxels, max_points, C) or coors.shape[0] != max_voxels \
            or num_points_per_voxel.shape[0] != max_voxels:
        raise RuntimeError("output buffers do not match (max_voxels, max_points, C)")
    F = 0
    if voxel_mean is not None:
        _lib.require_cuda(voxel_mean, "voxel_mean", torch.float32)
        F = voxel_mean.shape[1]
    if point2voxel is not None:
        _lib.require_cuda(point2voxel, "point2voxel", torch.int32)
    L = _lib.lib()
    with torch.cuda.device_of(points):
        nbytes = L.rd3_hard_voxelize_workspace_bytes(N, max_points, max_voxels)
        ws = _lib.workspace(points.device, nbytes)
        d_num = torch.empty(1, dtype=torch.int32, device=points.device)
        st = L.rd3_hard_voxelize(_lib.ptr(points), N, C, _lib.f3(voxel_size), _lib.f6(coors_range),
                                 int(max_points), int(max_voxels), _lib.ptr(voxels), _lib.ptr(coors),
                                 _lib.ptr(num_points_per_voxel), _lib.ptr(d_num),
                                 _lib.ptr(voxel_mean), F, _lib.ptr(point2voxel), _lib.ptr(ws),
                                 ws.numel(), _lib.stream_of(points))
        _lib.check(st, "hard_voxelize")
        return int(d_num.item())


def dynamic_voxelize(points, coors, voxel_size, coors_range, NDim=3):
    """voxelization.h:83-94.  Writes (z,y,x) or (-1,-1,-1) rows into ``coors`` (N,3) int32."""
    if NDim != 3:
        raise RuntimeError("rd3_b200.dynamic_voxelize supports NDim == 3 only")
    _lib.require_cuda(points, "points", torch.float32)
    _lib.require_cuda(coors, "coors", torch.int32)
    N, C = points.shape
    with torch.cuda.device_of(points):
        st = _lib.lib().rd3_dynamic_voxelize(_lib.ptr(points), N, C, _lib.f3(voxel_size),
                                             _lib.f6(coors_range), _lib.ptr(coors),
                                             _lib.stream_of(points))
        _lib.check(st, "dynamic_voxelize")


def dynamic_point_to_voxel_forward(feats, coors, reduce_type, dims=None):
    """voxelization.h:108-121 -> scatter_points_cuda.cu:183-239.

    Returns [voxel_feats (M,C), voxel_coors (M,ncols), point2voxel_map (N) int32,
    voxel_points_count (M) int32]; voxels in lexicographic coordinate order.
    ``coors`` is (N,3), or (N,4) with the batch index in front: the whole batch is then one launch
    sequence and the result is the reference's per-sample loop + concatenation (scatter_points.py:86-97).
    ``dims`` (optional, ncols ints): exclusive upper bound of valid coordinates (the
    voxel grid, with the sample count in front for (N,4)); when absent or exceeded it is measured on
    the device first; a ``dims`` LIST is updated in place with the bounds actually used, so a caller can keep
    them as the next call's hint.  One device->host read per call (M and the status word).
    """
    dims_arg = dims if isinstance(dims, list) else None
    code = _reduce_code(reduce_type)
    _lib.require_cuda(feats, "feats", torch.float32)
    _lib.require_cuda(coors, "coors", torch.int32)
    N, C = feats.shape
    if N == 0:                                        # scatter_points_cuda.cu:192-196
        return [feats.clone().detach(), coors.clone().detach(),
                coors.new_empty((0,), dtype=torch.int32), coors.new_empty((0,), dtype=torch.int32)]
    ncols = coors.shape[1] if coors.dim() == 2 else 0
    if coors.shape[0] != N or ncols not in (3, 4):
        raise RuntimeError("coors must be (N, 3) or (N, 4)")
    if ncols == 4 and coors.data_ptr() % 16:
        coors = coors.clone()                         # rows are moved as 16-byte words
    if dims is not None and len(dims) != ncols:
        raise RuntimeError("dims must have one entry per coors column")
    L = _lib.lib()
    dev = feats.device
    with torch.cuda.device_of(feats):
        stream = _lib.stream_of(feats)
        voxel_feats = torch.empty((N, C), dtype=torch.float32, device=dev)
        voxel_coors = torch.empty((N, ncols), dtype=torch.int32, device=dev)
        p2v = torch.empty((N,), dtype=torch.int32, device=dev)
        count = torch.empty((N,), dtype=torch.int32, device=dev)
        meta = torch.empty(2, dtype=torch.int32, device=dev)        # [M, status]
        for attempt in range(2):
            if dims is None:
                ext = torch.empty(4, dtype=torch.int32, device=dev)
                _lib.check(L.rd3_coors_extent(_lib.ptr(coors), N, ncols, _lib.ptr(ext), stream), "coors_extent")
                dims = [max(int(v), 1) for v in ext.tolist()[:ncols]]
            cd = _lib.i4(list(dims) if ncols == 4 else [1] + list(dims))
            nbytes = L.rd3_dynamic_scatter_workspace_bytes(N, C, ncols, cd)
            if nbytes == 0:
                raise RuntimeError("rd3_b200.dynamic_point_to_voxel_forward: coordinate extent %s "
                                   "too large for the bitmap path" % (list(dims),))
            ws = _lib.workspace(dev, nbytes)
            st = L.rd3_dynamic_scatter_forward(_lib.ptr(feats), _lib.ptr(coors), N, C, ncols, cd, code,
                                               _lib.ptr(voxel_feats), _lib.ptr(voxel_coors),
                                               _lib.ptr(p2v), _lib.ptr(count), _lib.ptr(meta[0:]),
                                               _lib.ptr(meta[1:]), _lib.ptr(ws), ws.numel(), stream)
            _lib.check(st, "dynamic_point_to_voxel_forward")
            M, status = meta.tolist()
            if status == 0:
                break
            dims = None                                   # hint was too small: measure and retry
        else:
            raise RuntimeError("rd3_b200.dynamic_point_to_voxel_forward: extent retry failed")
    if dims_arg is not None:
        dims_arg[:] = list(dims)
    if voxel_coors.dtype != coors.dtype:
        voxel_coors = voxel_coors.to(coors.dtype)
    return [voxel_feats[:M], voxel_coors[:M], p2v, count[:M]]


def dynamic_point_to_voxel_backward(grad_feats, grad_reduced_feats, feats, reduced_feats,
                                    coors_idx, reduce_count, reduce_type):
    """voxelization.h:123-140 -> scatter_points_cuda.cu:241-308 (writes grad_feats in place)."""
    code = _reduce_code(reduce_type)
    for t, n in ((grad_feats, "grad_feats"), (grad_reduced_feats, "grad_reduced_feats"),
                 (feats, "feats"), (reduced_feats, "reduced_feats")):
        _lib.require_cuda(t, n, torch.float32)
    _lib.require_cuda(coors_idx, "coors_idx", torch.int32)
    _lib.require_cuda(reduce_count, "reduce_count", torch.int32)
    N, C = feats.shape
    M = reduced_feats.shape[0]
    L = _lib.lib()
    with torch.cuda.device_of(feats):
        nbytes = L.rd3_dynamic_scatter_backward_workspace_bytes(M, C)
        ws = _lib.workspace(feats.device, nbytes)
        st = L.rd3_dynamic_scatter_backward(_lib.ptr(grad_feats), _lib.ptr(grad_reduced_feats),
                                            _lib.ptr(feats), _lib.ptr(reduced_feats),
                                            _lib.ptr(coors_idx), _lib.ptr(reduce_count), N, M, C,
                                            code, _lib.ptr(ws), ws.numel(), _lib.stream_of(feats))
        _lib.check(st, "dynamic_point_to_voxel_backward")
