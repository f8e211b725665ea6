"""Depth maps -> ego-frame points with the interface of
``ReconstructionBackbone._backproject_depth_to_points``
(projects/mmdet3d_plugin/models/backbone/reconstruction_backbone.py:285-386).
"""
import math
import struct

import torch

from . import _lib


def f32_ceil(x):
    """Smallest fp32 value >= x.  ``conf >= thr`` with an fp32 conf and a float64
    threshold (numpy percentile, tools/inference_nuscenes.py:360,401) is the same
    predicate as ``conf >= f32_ceil(thr)`` evaluated in fp32."""
    x = float(x)
    if math.isnan(x) or math.isinf(x):
        return x
    f = struct.unpack("f", struct.pack("f", x))[0]
    if f >= x:
        return f
    bits = struct.unpack("I", struct.pack("f", f))[0]
    if f == 0.0:
        bits = 1
    elif f > 0:
        bits += 1
    else:
        bits -= 1
    return struct.unpack("f", struct.pack("I", bits))[0]


def squeeze_head(t):
    """DA3's heads emit (B, N, H, W, 1) (output_processor.py:79-168 squeezes them): accept both, as a view."""
    if t is not None and torch.is_tensor(t) and t.dim() == 5 and t.shape[-1] == 1:
        return t.squeeze(-1)
    return t


def sky_arg(sky_masks, dev):
    """(uint8 mask | None, fp32 probability | None).  A boolean / integer tensor is a mask; a floating-point
    tensor is DA3's raw sky output, thresholded inside the kernels (sky iff >= 0.5, output_processor.py:165-167)
    so that the boolean tensor is never written."""
    if sky_masks is None:
        return None, None
    sky = squeeze_head(sky_masks).to(device=dev)
    if sky.is_floating_point():
        return None, sky.to(torch.float32).contiguous()
    sky = sky.contiguous()
    return (sky.view(torch.uint8) if sky.dtype == torch.bool else sky.to(torch.uint8)), None


SKY_PROB_THRESH = 0.5


def make_params(B, N, H, W, max_depth=None, conf_thresh=None, range_filter=None, sky_prob=None):
    p = _lib.DepthParams()
    p.sky_prob = sky_prob.data_ptr() if sky_prob is not None else None
    p.sky_prob_thresh = SKY_PROB_THRESH
    p._keepalive_sky = sky_prob
    p.B, p.ncam, p.H, p.W = int(B), int(N), int(H), int(W)
    p.use_max_depth = int(max_depth is not None)
    p.max_depth = float(max_depth) if max_depth is not None else 0.0
    if torch.is_tensor(conf_thresh):
        # per-sample thresholds that already live on the device (conf_threshold() below): no host read
        if not conf_thresh.is_cuda or conf_thresh.dtype != torch.float32 or conf_thresh.numel() != int(B) \
                or not conf_thresh.is_contiguous():
            raise RuntimeError("a tensor conf_thresh must be a contiguous CUDA fp32 tensor with one value per sample")
        p.conf_thresh = 0.0
        p.conf_thresh_dev = conf_thresh.data_ptr()
        p._keepalive = conf_thresh
    else:
        p.conf_thresh = f32_ceil(conf_thresh) if conf_thresh is not None else 0.0
        p.conf_thresh_dev = None
    p.use_range = int(range_filter is not None)
    for i in range(6):
        p.range[i] = float(range_filter[i]) if range_filter is not None else 0.0
    return p


def _prep(depths, intrinsics, cam2lidar, confs, sky_masks, conf_thresh):
    _lib.require_cuda(depths, "multi_batch_depths", torch.float32)
    if depths.dim() != 4:
        raise RuntimeError("multi_batch_depths must be (B, N, H, W)")
    dev = depths.device
    K = intrinsics.to(device=dev, dtype=torch.float32).contiguous()
    M = cam2lidar.to(device=dev, dtype=torch.float32).contiguous()
    B, N = depths.shape[:2]
    if tuple(K.shape) != (B, N, 3, 3) or tuple(M.shape) != (B, N, 4, 4):
        raise RuntimeError("intrinsics must be (B,N,3,3) and cam2lidar_rts (B,N,4,4)")
    conf = None
    if confs is not None and conf_thresh is not None:
        conf = squeeze_head(confs).to(device=dev, dtype=torch.float32).contiguous()
        if conf.numel() != depths.numel():
            raise RuntimeError("confs must have the shape of the depths")
    sky, sky_prob = sky_arg(sky_masks, dev)
    for t in (sky, sky_prob):
        if t is not None and t.numel() != depths.numel():
            raise RuntimeError("sky_masks must have the shape of the depths")
    return K, M, conf, sky, sky_prob


def unproject_padded(depths, intrinsics, cam2lidar_rts, max_depth=None, confs=None,
                     conf_thresh=None, sky_masks=None, range_filter=None, return_pix=False):
    """One launch sequence for all samples and cameras.

    Returns (points (B, N*H*W, 3) fp32, counts (B,) int32 on device[, pix (B, N*H*W) int32]):
    the first counts[b] rows of sample b are its points, cameras in index order
    and pixels row-major -- exactly the order of the reference's concatenation.
    No host synchronisation.
    """
    depths = squeeze_head(depths)
    K, M, conf, sky, sky_prob = _prep(depths, intrinsics, cam2lidar_rts, confs, sky_masks, conf_thresh)
    B, N, H, W = depths.shape
    dev = depths.device
    p = make_params(B, N, H, W, max_depth, conf_thresh if conf is not None else None, range_filter, sky_prob)
    L = _lib.lib()
    with torch.cuda.device_of(depths):
        pts = torch.empty((B, N * H * W, 3), dtype=torch.float32, device=dev)
        pix = torch.empty((B, N * H * W), dtype=torch.int32, device=dev) if return_pix else None
        counts = torch.empty((B,), dtype=torch.int32, device=dev)
        ws = _lib.workspace(dev, L.rd3_unproject_workspace_bytes(p))
        st = L.rd3_unproject(_lib.ptr(depths), _lib.ptr(K), _lib.ptr(M), _lib.ptr(conf), _lib.ptr(sky),
                             p, _lib.ptr(pts), _lib.ptr(pix), _lib.ptr(counts), _lib.ptr(ws),
                             ws.numel(), _lib.stream_of(depths))
        _lib.check(st, "unproject")
    return (pts, counts, pix) if return_pix else (pts, counts)


def backproject_depth_to_points(multi_batch_depths, multi_batch_intrinsics,
                                multi_batch_ori_imgs=None, multi_batch_cam2lidar_rts=None,
                                max_depth=None, multi_batch_confs=None, conf_thresh=None,
                                multi_batch_sky_masks=None, range_filter=None):
    """reconstruction_backbone.py:285-386 as a function.

    Returns (list_B[(P_b,3) fp32], list_B[(P_b,3) fp32 | None]).  The colour
    branch (:350-365) gathers the per-pixel colours of the valid points; images
    of another resolution are bilinearly resized first (torch, on the GPU) and
    values are divided by 255 per camera when that camera's colours exceed 1.5.
    """
    if multi_batch_cam2lidar_rts is None:
        raise RuntimeError("multi_batch_cam2lidar_rts is required (the reference indexes it unconditionally)")
    depths = squeeze_head(multi_batch_depths).contiguous()
    want_cols = multi_batch_ori_imgs is not None
    out = unproject_padded(depths, multi_batch_intrinsics, multi_batch_cam2lidar_rts, max_depth,
                           multi_batch_confs, conf_thresh, multi_batch_sky_masks, range_filter,
                           return_pix=want_cols)
    pts, counts = out[0], out[1]
    B, N, H, W = depths.shape
    n = counts.tolist()                               # the one unavoidable D2H (ragged outputs)
    points = [pts[b, :n[b]] for b in range(B)]
    if not want_cols:
        return points, [None] * B
    pix = out[2]
    colors = []
    for b in range(B):
        if n[b] == 0:
            colors.append(None)
            continue
        img = multi_batch_ori_imgs[b].to(depths.device)
        if img.dtype != torch.float:
            img = img.float()
        if img.shape[-2] != H or img.shape[-1] != W:
            img = torch.nn.functional.interpolate(img, size=(H, W), mode='bilinear', align_corners=False)
        flat = img.permute(0, 2, 3, 1).reshape(-1, 3)          # (N*H*W, 3), pixel-major
        idx = pix[b, :n[b]].long()
        cols = flat[idx]
        cam = idx // (H * W)
        cam_max = torch.full((N,), -float("inf"), device=cols.device).scatter_reduce_(
            0, cam, cols.max(dim=1).values, reduce="amax", include_self=True)
        scale = torch.where(cam_max > 1.5, 255.0, 1.0)
        colors.append(cols / scale[cam].unsqueeze(1))
    return points, colors


class DepthToPointsMixin:
    """Drop-in method for ``ReconstructionBackbone``: same name, arguments and
    return value as reconstruction_backbone.py:285-304; reads ``self.max_depth``."""

    max_depth = None

    def _backproject_depth_to_points(self, multi_batch_depths, multi_batch_intrinsics,
                                     multi_batch_ori_imgs=None, multi_batch_cam2lidar_rts=None):
        return backproject_depth_to_points(multi_batch_depths, multi_batch_intrinsics,
                                           multi_batch_ori_imgs, multi_batch_cam2lidar_rts,
                                           max_depth=getattr(self, "max_depth", None))


def conf_threshold(confs, sky_masks=None, percentile=30.0, numpy2=True, return_float64=False):
    """Per-sample ``np.percentile(conf[~sky] if (~sky).sum() > 10 else conf.flatten(), percentile)``
    (tools/inference_nuscenes.py:351-361) computed on the device by an exact radix select.

    confs (B, N, H, W[, 1]) fp32 CUDA, sky_masks (B, N, H, W[, 1]) bool/uint8, DA3's raw fp32 sky output
    (sky iff >= 0.5) or None.  Returns a (B,) fp32
    CUDA tensor that can be passed as ``conf_thresh`` to ``unproject_padded`` / ``DepthToVoxels`` /
    ``backproject_depth_to_points`` without a host round trip.  ``numpy2``: follow NumPy >= 2's
    fp32 index arithmetic for fp32 data (what ``np.percentile`` of the installed numpy returns);
    False: the fp64 arithmetic of NumPy < 2 (the reference's pin), rounded to fp32 like its
    ``conf >= thresh`` comparison does.  ``return_float64``: also return the unrounded values."""
    confs = squeeze_head(confs)
    _lib.require_cuda(confs, "confs", torch.float32)
    B = confs.shape[0]
    npix = confs[0].numel()
    sky, sky_prob = sky_arg(sky_masks, confs.device)
    for t in (sky, sky_prob):
        if t is not None and t.numel() != confs.numel():
            raise RuntimeError("sky_masks must have the shape of confs")
    t64 = torch.empty((B,), dtype=torch.float64, device=confs.device)
    t32 = torch.empty((B,), dtype=torch.float32, device=confs.device)
    L = _lib.lib()
    with torch.cuda.device_of(confs):
        ws = _lib.workspace(confs.device, L.rd3_conf_percentile_workspace_bytes(B))
        if sky_prob is not None:
            st = L.rd3_conf_percentile_skyprob(_lib.ptr(confs), _lib.ptr(sky_prob), SKY_PROB_THRESH, B, npix,
                                               float(percentile), int(bool(numpy2)), _lib.ptr(t64), _lib.ptr(t32),
                                               None, _lib.ptr(ws), ws.numel(), _lib.stream_of(confs))
        else:
            st = L.rd3_conf_percentile(_lib.ptr(confs), _lib.ptr(sky), B, npix, float(percentile),
                                       int(bool(numpy2)), _lib.ptr(t64), _lib.ptr(t32), None, _lib.ptr(ws),
                                       ws.numel(), _lib.stream_of(confs))
        _lib.check(st, "conf_percentile")
    return (t32, t64) if return_float64 else t32
