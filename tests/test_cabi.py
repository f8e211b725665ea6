"""The C-ABI library loads and exports every symbol include/rd3_b200.h declares.
CPU only: no kernel is launched; only host-side helpers (grid size, workspace
sizes, status strings, argument validation) are called."""
import ctypes
import os
import re

import pytest

import rd3_b200
from rd3_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rd3_b200.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"RD3_API\s+[\w\s\*]+?\b(rd3_\w+)\s*\(", src)))


def test_library_is_built_in_tree():
    assert os.path.exists(_lib.LIB_PATH), "run python 3d-reconstruction-detection_b200/build.py"
    assert os.path.dirname(_lib.LIB_PATH).startswith(ROOT)


def test_every_declared_symbol_is_exported_and_bound():
    names = header_symbols()
    assert len(names) >= 19
    L = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), "missing export: " + n
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header disagree"
    assert _lib.lib().rd3_version() >= 100


def test_status_strings_and_errors():
    L = _lib.lib()
    assert L.rd3_status_string(0) == b"ok"
    assert b"workspace" in L.rd3_status_string(2)
    with pytest.raises(RuntimeError):
        _lib.check(3, "x")


def test_grid_size_matches_reference_formula():
    """round((max-min)/vs) in fp32: voxelization_cpu.cpp:121-124 / voxelize.py:113-121"""
    import torch
    L = _lib.lib()
    for vs, pcr in [([0.075, 0.075, 0.2], [-54, -54, -5, 54, 54, 3]),
                    ([0.2, 0.2, 8.0], [-51.2, -51.2, -5, 51.2, 51.2, 3]),
                    ([0.5, 0.5, 0.5], [0, -40, -3, 70.4, 40, 1]),
                    ([0.32, 0.32, 6.0], [-74.88, -74.88, -2, 74.88, 74.88, 4])]:
        g = _lib.i3([0, 0, 0])
        assert L.rd3_grid_size(_lib.f3(vs), _lib.f6(pcr), g) == 0
        ref = torch.round((torch.tensor(pcr[3:], dtype=torch.float32) - torch.tensor(pcr[:3], dtype=torch.float32))
                          / torch.tensor(vs, dtype=torch.float32)).long().tolist()
        assert list(g) == ref
        assert rd3_b200.Voxelization(vs, pcr, 10).grid_size.tolist() == ref
    assert L.rd3_grid_size(_lib.f3([0, 1, 1]), _lib.f6([0, 0, 0, 1, 1, 1]), g) == 1   # invalid voxel size


def test_workspace_sizes_and_argument_validation_without_gpu():
    L = _lib.lib()
    n = 6 * 504 * 896
    a = L.rd3_hard_voxelize_workspace_bytes(n, 10, 120000)
    assert 10 << 20 < a < 200 << 20
    assert L.rd3_hard_voxelize_workspace_bytes(-1, 10, 10) == 0
    # the table is bounded by max_voxels + round length, not by N
    assert L.rd3_hard_voxelize_workspace_bytes(8 * n, 10, 120000) < 8 * a
    p = rd3_b200.backproject.make_params(2, 6, 504, 896, max_depth=100.0)
    assert L.rd3_unproject_workspace_bytes(ctypes.byref(p)) > 0
    assert L.rd3_depth_to_voxels_workspace_bytes(ctypes.byref(p), 10, 120000) > 2 * (a // 2)
    assert L.rd3_dynamic_scatter_workspace_bytes(1000, 3, 3, _lib.i4([1, 40, 1440, 1440])) > 10 << 20
    assert L.rd3_dynamic_scatter_workspace_bytes(1000, 3, 4, _lib.i4([4, 40, 1440, 1440])) > 40 << 20
    assert L.rd3_dynamic_scatter_workspace_bytes(1000, 3, 3, _lib.i4([1, 70000, 70000, 70000])) == 0  # too large
    assert L.rd3_dynamic_scatter_workspace_bytes(1000, 3, 4, _lib.i4([64, 40, 1440, 1440])) == 0       # > 2^32 cells
    # null pointers / bad sizes are rejected before any CUDA call
    null = ctypes.c_void_p(0)
    st = L.rd3_hard_voxelize(null, 10, 3, _lib.f3([1, 1, 1]), _lib.f6([0, 0, 0, 1, 1, 1]), 5, 5, null, null, null,
                             null, null, 0, null, null, 0, null)
    assert st == 1
    st = L.rd3_dynamic_voxelize(null, -5, 3, _lib.f3([1, 1, 1]), _lib.f6([0, 0, 0, 1, 1, 1]), null, null)
    assert st == 1
    # a workspace that is not 256-byte aligned is an invalid argument (256-bit table loads, TMA bulk copies);
    # the pointers are never dereferenced on the host and the check comes before the first CUDA call
    fake = lambda a: ctypes.c_void_p(0x10000 + a)
    vs, pcr = _lib.f3([0.075, 0.075, 0.2]), _lib.f6([-54, -54, -5, 54, 54, 3])
    st = L.rd3_depth_to_voxels(fake(0), fake(0), fake(0), null, null, ctypes.byref(p), vs, pcr, 10, 1000, fake(0),
                               fake(0), fake(0), null, fake(0), fake(64), 1 << 40, null)
    assert st == 1
    st = L.rd3_hard_voxelize(fake(0), 10, 3, vs, pcr, 5, 5, fake(0), fake(0), fake(0), fake(0), null, 0, null,
                             fake(128), 1 << 40, null)
    assert st == 1


def test_no_fallback_without_cuda_tensors():
    import torch
    with pytest.raises(RuntimeError, match="CUDA"):
        rd3_b200.voxel_layer.dynamic_voxelize(torch.zeros(4, 3), torch.zeros(4, 3, dtype=torch.int32),
                                              [1, 1, 1], [0, 0, 0, 1, 1, 1])
    with pytest.raises(RuntimeError, match="CUDA"):
        rd3_b200.HardSimpleVFE(3)(torch.zeros(2, 5, 3), torch.ones(2, dtype=torch.int32), None)
    with pytest.raises(RuntimeError, match="CUDA"):
        rd3_b200.backproject_depth_to_points(torch.zeros(1, 6, 8, 8), torch.zeros(1, 6, 3, 3), None,
                                             torch.zeros(1, 6, 4, 4))
    with pytest.raises(RuntimeError, match="reduce type"):
        rd3_b200.voxel_layer.dynamic_point_to_voxel_forward(torch.zeros(1, 3), torch.zeros(1, 3, dtype=torch.int32), "min")
