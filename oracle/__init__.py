"""CPU oracle for the depth->voxel hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package, and there
only as the checker.  The product package (``rd3_b200``) never imports it and
fails loudly when its CUDA library is missing.

Three layers, strongest first:

* ``oracle/_ref``  -- the reference's own C++ CPU op compiled unmodified from
  /root/reference (``build_ref.py``); loaded with :func:`ref_voxel_layer`.
* ``rd3_oracle.c`` -- plain-C restatement (numpy in / numpy out wrappers below),
  each function citing the reference file:line it follows.
* ``torch_restatement.py`` -- torch-CPU transliteration of the Python-level
  reference code that cannot be imported here (mmcv/mmdet are not installed).

Parity status: pinned (see the header of rd3_oracle.c and tests/test_oracle.py).
"""
import ctypes
import os

import numpy as np

from . import build as _build
from .build_ref import load_ref as ref_voxel_layer  # noqa: F401

_LIB = None

_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)
_i32p = ctypes.POINTER(ctypes.c_int32)
_u8p = ctypes.POINTER(ctypes.c_uint8)


def lib():
    global _LIB
    if _LIB is None:
        path = _build.LIB
        if not os.path.exists(path) or (
                os.path.exists(_build.SRC)
                and os.path.getmtime(path) < os.path.getmtime(_build.SRC)):
            path = _build.build()
        L = ctypes.CDLL(path)
        L.orc_grid_size.argtypes = [_f32p, _f32p, _i32p]
        L.orc_grid_size.restype = None
        L.orc_dynamic_voxelize.argtypes = [_f32p, ctypes.c_int64, ctypes.c_int,
                                           _f32p, _f32p, _i32p]
        L.orc_dynamic_voxelize.restype = None
        L.orc_hard_voxelize.argtypes = [_f32p, ctypes.c_int64, ctypes.c_int,
                                        _f32p, _f32p, ctypes.c_int,
                                        ctypes.c_int, _f32p, _i32p, _i32p,
                                        _i32p]
        L.orc_hard_voxelize.restype = ctypes.c_int
        L.orc_hard_simple_vfe.argtypes = [_f32p, _i32p, ctypes.c_int64,
                                          ctypes.c_int, ctypes.c_int,
                                          ctypes.c_int, _f32p]
        L.orc_hard_simple_vfe.restype = None
        L.orc_hard_simple_vfe_f64.argtypes = [_f32p, _i32p, ctypes.c_int64,
                                              ctypes.c_int, ctypes.c_int,
                                              ctypes.c_int, _f64p]
        L.orc_hard_simple_vfe_f64.restype = None
        L.orc_unproject.argtypes = [_f32p, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_int, _f32p, _f32p, ctypes.c_int,
                                    ctypes.c_float, _f32p, ctypes.c_float,
                                    _u8p, _f32p, _f32p, _i32p, ctypes.c_int64]
        L.orc_unproject.restype = ctypes.c_int64
        L.orc_dynamic_scatter.argtypes = [_f32p, _i32p, ctypes.c_int64,
                                          ctypes.c_int, ctypes.c_int, _f32p,
                                          _i32p, _i32p, _i32p]
        L.orc_dynamic_scatter.restype = ctypes.c_int64
        _LIB = L
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def grid_size(voxel_size, coors_range):
    """(gx, gy, gz) = round((max-min)/vs) in fp32 (voxelization_cpu.cpp:121-124)."""
    vs, cr = _f32(voxel_size), _f32(coors_range)
    g = np.zeros(3, np.int32)
    lib().orc_grid_size(_ptr(vs, _f32p), _ptr(cr, _f32p), _ptr(g, _i32p))
    return g


def dynamic_voxelize(points, voxel_size, coors_range):
    """points (N,C) -> coors (N,3) int32 zyx or -1 (voxelization_cpu.cpp:7-43)."""
    pts = _f32(points)
    N, C = pts.shape
    vs, cr = _f32(voxel_size), _f32(coors_range)
    coors = np.zeros((N, 3), np.int32)
    lib().orc_dynamic_voxelize(_ptr(pts, _f32p), N, C, _ptr(vs, _f32p),
                               _ptr(cr, _f32p), _ptr(coors, _i32p))
    return coors


def hard_voxelize(points, voxel_size, coors_range, max_points, max_voxels,
                  return_point2voxel=False):
    """Sequential hard voxelization (voxelization_cpu.cpp:45-101).

    Returns (voxels (M,max_points,C), coors (M,3) zyx, num (M)) sliced to the
    voxel count like voxelize.py:67-70, plus point2voxel (N) on request.
    """
    pts = _f32(points)
    N, C = pts.shape
    vs, cr = _f32(voxel_size), _f32(coors_range)
    voxels = np.zeros((max_voxels, max_points, C), np.float32)
    coors = np.zeros((max_voxels, 3), np.int32)
    num = np.zeros((max_voxels,), np.int32)
    p2v = np.zeros((N,), np.int32)
    m = lib().orc_hard_voxelize(_ptr(pts, _f32p), N, C, _ptr(vs, _f32p),
                                _ptr(cr, _f32p), max_points, max_voxels,
                                _ptr(voxels, _f32p), _ptr(coors, _i32p),
                                _ptr(num, _i32p), _ptr(p2v, _i32p))
    if m < 0:
        raise MemoryError("oracle could not allocate the dense grid")
    out = (voxels[:m], coors[:m], num[:m])
    return out + (p2v,) if return_point2voxel else out


def hard_simple_vfe(voxels, num, num_features=None, f64=False):
    """voxel mean, voxel_encoder.py:45-46 (fp32 sequential; f64=True anchor)."""
    v = _f32(voxels)
    M, K, C = v.shape
    F = C if num_features is None else num_features
    n = np.ascontiguousarray(num, dtype=np.int32)
    if f64:
        out = np.zeros((M, F), np.float64)
        lib().orc_hard_simple_vfe_f64(_ptr(v, _f32p), _ptr(n, _i32p), M, K, C,
                                      F, _ptr(out, _f64p))
    else:
        out = np.zeros((M, F), np.float32)
        lib().orc_hard_simple_vfe(_ptr(v, _f32p), _ptr(n, _i32p), M, K, C, F,
                                  _ptr(out, _f32p))
    return out


def _f32_ceil(x):
    """smallest fp32 >= x: `conf_f32 >= thr_f64` (numpy, tools/inference_nuscenes.py:401)
    evaluated exactly, expressed as an fp32 threshold."""
    f = np.float32(x)
    if float(f) < float(x):
        f = np.nextafter(f, np.float32(np.inf))
    return float(f)


def unproject(depth, intrinsics, cam2lidar, max_depth=None, conf=None,
              conf_thresh=0.0, sky=None, range_filter=None, return_pix=False):
    """One sample: depth (ncam,H,W) -> ego points (P,3) fp32.

    reconstruction_backbone.py:305-386 (+ the live conf/sky masks of
    tools/inference_nuscenes.py:399-414 and the inclusive range filter of
    respoint_post_processing.py:190-195 when given).
    """
    d = _f32(depth)
    ncam, H, W = d.shape
    K = _f32(intrinsics).reshape(ncam, 9)
    M = _f32(cam2lidar).reshape(ncam, 16)
    cap = ncam * H * W
    pts = np.zeros((cap, 3), np.float32)
    pix = np.zeros((cap,), np.int32) if return_pix else None
    cf = _f32(conf) if conf is not None else None
    sk = np.ascontiguousarray(sky, dtype=np.uint8) if sky is not None else None
    rf = _f32(range_filter) if range_filter is not None else None
    P = lib().orc_unproject(_ptr(d, _f32p), ncam, H, W, _ptr(K, _f32p),
                            _ptr(M, _f32p), int(max_depth is not None),
                            float(max_depth if max_depth is not None else 0.0),
                            _ptr(cf, _f32p), _f32_ceil(conf_thresh),
                            _ptr(sk, _u8p), _ptr(rf, _f32p), _ptr(pts, _f32p),
                            _ptr(pix, _i32p), cap)
    if return_pix:
        return pts[:P].copy(), pix[:P].copy()
    return pts[:P].copy()


_REDUCE = {"sum": 0, "mean": 1, "max": 2}


def dynamic_scatter(feats, coors, reduce_type):
    """scatter_points_cuda.cu:183-239 on the CPU.

    Returns (voxel_feats (M,C), voxel_coors (M,3), point2voxel_map (N) int32,
    voxel_points_count (M) int32).
    """
    if reduce_type not in _REDUCE:
        raise RuntimeError("do not support reduce type " + reduce_type)
    f = _f32(feats)
    c = np.ascontiguousarray(coors, dtype=np.int32)
    N, C = f.shape
    of = np.zeros((max(N, 1), C), np.float32)
    oc = np.zeros((max(N, 1), 3), np.int32)
    mp = np.zeros((max(N, 1),), np.int32)
    ct = np.zeros((max(N, 1),), np.int32)
    M = lib().orc_dynamic_scatter(_ptr(f, _f32p), _ptr(c, _i32p), N, C,
                                  _REDUCE[reduce_type], _ptr(of, _f32p),
                                  _ptr(oc, _i32p), _ptr(mp, _i32p),
                                  _ptr(ct, _i32p))
    return of[:M].copy(), oc[:M].copy(), mp[:N].copy(), ct[:M].copy()


def soft_voxel_occupancy_f64(voxels, num, lambda_n=0.3, gamma_var=5.0, eps=1e-6):
    """SoftVoxelOccupancyVFE.forward (voxel_occupancy_encoder.py:60-99) evaluated in fp64, plus a
    per-voxel bound on what ANY fp32 evaluation of the same formula (whatever its summation order) may
    differ from it.  Returns (p_occ (M,1) float64, tol (M,1) float64).

    Derivation, u = 2^-24 (fp32 unit roundoff), X = max |xyz| of the voxel, D = max |xyz - mean|:
      mean   n-term sum, any order, then one division:  dm   <= u X (n + 2)
      diff   one subtraction on top of the mean's error: dd  <= dm + u D
      var    sum of n squares / denom, mean of 3 axes:   dvar <= 2 D dd + u D^2 (n + 4)
             (the cancellation |xyz| ~ 50 m against |diff| ~ 0.2 m sits in the 2 D dm term: ~1e-4 relative)
      a = -lambda n - gamma var:                          da   <= gamma dvar + 3 u |a|
      p = 1 - exp(a), exp within 2 ulp:                   dp   <= exp(a) (da + 3 u) + u
    The returned tol is twice that bound.
    """
    v = np.asarray(voxels, dtype=np.float64)[:, :, :3]
    n = np.asarray(num, dtype=np.float64)
    M, K, _ = v.shape
    mask = (np.arange(K)[None, :] < n[:, None])[:, :, None].astype(np.float64)
    denom = n[:, None] + np.float64(np.float32(eps))
    mean = (v * mask).sum(axis=1) / denom
    diff = (v - mean[:, None, :]) * mask
    var = ((diff ** 2).sum(axis=1) / denom).mean(axis=1)
    a = -np.float64(np.float32(lambda_n)) * n - np.float64(np.float32(gamma_var)) * var
    p = 1.0 - np.exp(a)
    u = 2.0 ** -24
    X = np.abs(v * mask).max(axis=(1, 2))
    D = np.abs(diff).max(axis=(1, 2))
    dm = u * X * (n + 2)
    dd = dm + u * D
    dvar = 2 * D * dd + u * D * D * (n + 4)
    da = gamma_var * dvar + 3 * u * np.abs(a)
    dp = np.exp(a) * (da + 3 * u) + u
    return p.reshape(-1, 1), (2 * dp).reshape(-1, 1)


def masked_mean_f64(voxels, num, num_features=None):
    """Per-voxel mean of the first num[m] slots in fp64 and the bound 2 (n + 2) 2^-24 max|x| on the
    difference of any fp32 evaluation (n-term sum in any order + one division).  Returns (mean, tol)."""
    v = np.asarray(voxels, dtype=np.float64)
    n = np.asarray(num, dtype=np.float64)
    M, K, C = v.shape
    F = C if num_features is None else num_features
    mask = (np.arange(K)[None, :] < n[:, None])[:, :, None]
    vm = np.where(mask, v[:, :, :F], 0.0)
    mean = vm.sum(axis=1) / np.maximum(n, 1.0)[:, None]
    X = np.abs(vm).max(axis=1)
    tol = 2 * (n[:, None] + 2) * 2.0 ** -24 * X
    return mean, tol
