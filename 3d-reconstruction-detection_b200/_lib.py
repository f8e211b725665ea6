"""ctypes binding of librd3_b200.so (the C ABI in include/rd3_b200.h).

There is no fallback: if the library is missing or a call fails, a RuntimeError
is raised.  Nothing in this package computes on the CPU.
"""
import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# RD3_LIB_PATH lets A/B experiments load an alternative build of the same C ABI
LIB_PATH = os.environ.get("RD3_LIB_PATH") or os.path.join(_HERE, "librd3_b200.so")

_c = ctypes
_vp, _i32, _i64, _sz, _f32 = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_size_t, _c.c_float
_F3, _F6, _I3, _I4 = _c.c_float * 3, _c.c_float * 6, _c.c_int32 * 3, _c.c_int32 * 4


class DepthParams(ctypes.Structure):
    """struct rd3_depth_params"""
    _fields_ = [("B", _c.c_int32), ("ncam", _c.c_int32), ("H", _c.c_int32), ("W", _c.c_int32),
                ("use_max_depth", _c.c_int32), ("max_depth", _c.c_float),
                ("conf_thresh", _c.c_float), ("use_range", _c.c_int32),
                ("range", _c.c_float * 6), ("conf_thresh_dev", _c.c_void_p),
                ("sky_prob", _c.c_void_p), ("sky_prob_thresh", _c.c_float)]


# name -> (restype, argtypes); must list every symbol include/rd3_b200.h declares
SIGNATURES = {
    "rd3_version": (_i32, []),
    "rd3_status_string": (_c.c_char_p, [_i32]),
    "rd3_last_cuda_error": (_c.c_char_p, []),
    "rd3_profile_enable": (_i32, [_i32]),
    "rd3_profile_read": (_i32, [_c.POINTER(_c.c_double * 7), _c.POINTER(_c.c_int)]),
    "rd3_grid_size": (_i32, [_F3, _F6, _I3]),
    "rd3_dynamic_voxelize": (_i32, [_vp, _i64, _i32, _F3, _F6, _vp, _vp]),
    "rd3_hard_voxelize_workspace_bytes": (_sz, [_i64, _i32, _i32]),
    "rd3_hard_voxel_rounds": (_i32, [_i64, _i32]),
    "rd3_hard_voxelize": (_i32, [_vp, _i64, _i32, _F3, _F6, _i32, _i32, _vp, _vp, _vp, _vp, _vp,
                                 _i32, _vp, _vp, _sz, _vp]),
    "rd3_hard_simple_vfe": (_i32, [_vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp]),
    "rd3_pack_sparse_inputs": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp,
                                      _vp]),
    "rd3_unproject_workspace_bytes": (_sz, [_c.POINTER(DepthParams)]),
    "rd3_unproject": (_i32, [_vp, _vp, _vp, _vp, _vp, _c.POINTER(DepthParams), _vp, _vp, _vp, _vp,
                             _sz, _vp]),
    "rd3_depth_to_voxels_workspace_bytes": (_sz, [_c.POINTER(DepthParams), _i32, _i32]),
    "rd3_depth_to_voxels": (_i32, [_vp, _vp, _vp, _vp, _vp, _c.POINTER(DepthParams), _F3, _F6, _i32,
                                   _i32, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "rd3_pillar_decorate": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _f32,
                                   _f32, _f32, _vp, _vp]),
    "rd3_pillars_scatter": (_i32, [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "rd3_map_voxel_to_point_workspace_bytes": (_sz, [_i64]),
    "rd3_map_voxel_to_point": (_i32, [_vp, _i64, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _sz, _vp]),
    "rd3_conf_percentile_workspace_bytes": (_sz, [_i32]),
    "rd3_conf_percentile": (_i32, [_vp, _vp, _i32, _i64, _c.c_double, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "rd3_conf_percentile_skyprob": (_i32, [_vp, _vp, _f32, _i32, _i64, _c.c_double, _i32, _vp, _vp, _vp, _vp, _sz,
                                           _vp]),
    "rd3_voxel_occupancy": (_i32, [_vp, _vp, _i64, _i32, _i32, _i32, _f32, _f32, _f32, _vp, _vp, _i32, _i32, _i32,
                                   _i32, _i32, _vp, _vp]),
    "rd3_coors_extent": (_i32, [_vp, _i64, _i32, _vp, _vp]),
    "rd3_dynamic_scatter_workspace_bytes": (_sz, [_i64, _i32, _i32, _I4]),
    "rd3_dynamic_scatter_forward": (_i32, [_vp, _vp, _i64, _i32, _i32, _I4, _i32, _vp, _vp, _vp, _vp, _vp,
                                           _vp, _vp, _sz, _vp]),
    "rd3_dynamic_scatter_backward_workspace_bytes": (_sz, [_i64, _i32]),
    "rd3_dynamic_scatter_backward": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i32, _i32,
                                            _vp, _sz, _vp]),
}

_lib = None
_lock = threading.Lock()


def lib():
    """Load librd3_b200.so once.  Raises if it was not built (no fallback)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        "rd3_b200: %s not found. Build it with "
                        "`python 3d-reconstruction-detection_b200/build.py` "
                        "(there is no CPU/PyTorch fallback)." % LIB_PATH)
                L = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(L, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = L
    return _lib


def check(status, what):
    if status != 0:
        L = lib()
        msg = L.rd3_status_string(status).decode()
        if status == 4:
            msg += ": " + L.rd3_last_cuda_error().decode()
        raise RuntimeError("rd3_b200.%s failed: %s" % (what, msg))


def ptr(t):
    return _vp(t.data_ptr()) if t is not None else _vp(0)


def stream_of(t):
    return _vp(torch.cuda.current_stream(t.device).cuda_stream)


def f3(v):
    return _F3(*[float(x) for x in v])


def f6(v):
    return _F6(*[float(x) for x in v])


def i3(v):
    return _I3(*[int(x) for x in v])


def i4(v):
    return _I4(*[int(x) for x in v])


def require_cuda(t, name, dtype=None):
    if not torch.is_tensor(t) or not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor (rd3_b200 has no CPU path)" % name)
    if not t.is_contiguous():
        raise RuntimeError("%s must be contiguous" % name)
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError("%s must be %s, got %s" % (name, dtype, t.dtype))


# grow-only scratch buffer per (device, stream).  Work on one stream is ordered, so calls on the same stream may
# share scratch; different streams get their own.  Streams come and go (stream pools, graph captures): the
# table keeps the most recently used _MAX_WORKSPACES entries.
_workspaces = {}
_MAX_WORKSPACES = 16


def workspace(device, nbytes):
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.pop(key, None)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
    _workspaces[key] = ws                       # re-inserted last: dicts keep insertion order
    while len(_workspaces) > _MAX_WORKSPACES:
        _workspaces.pop(next(iter(_workspaces)))
    return ws


PROFILE_STAGES = ("memset", "insert", "flags", "cull", "lookup", "emit", "meta")


def profile_enable(on=True):
    check(lib().rd3_profile_enable(int(on)), "profile_enable")


def profile_read():
    """-> (dict stage -> summed ms, calls)"""
    ms = (_c.c_double * 7)()
    calls = _c.c_int(0)
    check(lib().rd3_profile_read(_c.byref(ms), _c.byref(calls)), "profile_read")
    return dict(zip(PROFILE_STAGES, list(ms))), calls.value


def release_workspaces():
    _workspaces.clear()
