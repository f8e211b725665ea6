"""Rebind the reference's operator modules to this implementation.

``patch_mmdet3d()`` replaces, when ``mmdet3d`` is importable, the symbols every
call site of the reference goes through:

    mmdet3d.ops.voxel.voxelize:        hard_voxelize, dynamic_voxelize, Voxelization, voxelization
    mmdet3d.ops.voxel.scatter_points:  dynamic_point_to_voxel_forward/backward, DynamicScatter, dynamic_scatter
    mmdet3d.ops (and mmdet3d.ops.voxel): Voxelization, voxelization, DynamicScatter, dynamic_scatter
    mmdet3d.models.voxel_encoders.voxel_encoder.HardSimpleVFE.forward
    every already imported mmdet3d.* / projects.* module that holds the old Voxelization / DynamicScatter through a
    ``from mmdet3d.ops import ...`` (DynamicSimpleVFE, DynamicVFE, the pillar encoders, SparseRefinement, VoxelDownsample)
    ReconstructionBackbone._backproject_depth_to_points   (the plugin, when it is importable)

so existing configs (``pts_voxel_layer=dict(max_num_points=..., voxel_size=..., ...)``) run
unchanged.  mmdet3d / mmcv are not installed in the build container; the patch
is exercised against stand-in modules in tests/test_host_logic.py.
"""
import importlib
import sys

from . import backproject, scatter_points, voxel_encoder, voxel_layer, voxelize


def _try_import(name):
    try:
        return importlib.import_module(name)
    except Exception:
        return sys.modules.get(name)


def patch_mmdet3d(verbose=False):
    """Returns the list of 'module.attribute' names that were rebound."""
    done = []

    def rebind(mod, attr, value):
        if mod is not None and hasattr(mod, attr):
            setattr(mod, attr, value)
            done.append("%s.%s" % (mod.__name__, attr))

    # the objects call sites may already hold through ``from mmdet3d.ops import Voxelization, DynamicScatter``
    # (voxel_encoder.py, pillar_encoder.py, sparse_refinement.py:15, respoint_post_processing.py:15, mvx_two_stage.py)
    stale = {}
    for name, attrs, new_mod in (("mmdet3d.ops.voxel.voxelize", ("Voxelization", "voxelization"), voxelize),
                                 ("mmdet3d.ops.voxel.scatter_points", ("DynamicScatter", "dynamic_scatter"),
                                  scatter_points)):
        m = _try_import(name)
        for a in attrs:
            old = getattr(m, a, None) if m is not None else None
            if old is not None and old is not getattr(new_mod, a):
                stale[a] = (old, getattr(new_mod, a))

    m = _try_import("mmdet3d.ops.voxel.voxelize")
    for a in ("hard_voxelize", "dynamic_voxelize"):
        rebind(m, a, getattr(voxel_layer, a))
    for a in ("Voxelization", "voxelization", "_Voxelization"):
        rebind(m, a, getattr(voxelize, a))
    m = _try_import("mmdet3d.ops.voxel.scatter_points")
    for a in ("dynamic_point_to_voxel_forward", "dynamic_point_to_voxel_backward"):
        rebind(m, a, getattr(voxel_layer, a))
    for a in ("DynamicScatter", "dynamic_scatter", "_dynamic_scatter"):
        rebind(m, a, getattr(scatter_points, a))
    for name in ("mmdet3d.ops.voxel", "mmdet3d.ops"):
        m = _try_import(name)
        for a in ("Voxelization", "voxelization"):
            rebind(m, a, getattr(voxelize, a))
        for a in ("DynamicScatter", "dynamic_scatter"):
            rebind(m, a, getattr(scatter_points, a))
    for mname, m in list(sys.modules.items()):
        if m is None or not (mname.startswith("mmdet3d") or mname.startswith("projects.")):
            continue
        for a, (old, new) in stale.items():
            if getattr(m, a, None) is old:
                setattr(m, a, new)
                done.append("%s.%s" % (mname, a))
    m = _try_import("mmdet3d.models.voxel_encoders.voxel_encoder")
    if m is not None and hasattr(m, "HardSimpleVFE"):
        def forward(self, features, num_points, coors=None):
            return voxel_encoder.HardSimpleVFE.forward(self, features, num_points, coors)
        m.HardSimpleVFE.forward = forward
        done.append(m.__name__ + ".HardSimpleVFE.forward")
    m = _try_import("projects.mmdet3d_plugin.models.backbone.reconstruction_backbone")
    if m is not None and hasattr(m, "ReconstructionBackbone"):
        m.ReconstructionBackbone._backproject_depth_to_points = \
            backproject.DepthToPointsMixin._backproject_depth_to_points
        done.append(m.__name__ + ".ReconstructionBackbone._backproject_depth_to_points")
    if verbose:
        print("rd3_b200.patch_mmdet3d:", ", ".join(done) if done else "nothing to patch")
    return done
