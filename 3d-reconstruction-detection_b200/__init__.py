"""rd3_b200 -- B200-native depth->voxel path (unproject, voxelize, scatter, VFE).

Importable as ``rd3_b200`` (see /rd3_b200.py at the repo root).  The public
names mirror the reference's operator API for this path:

    Voxelization, voxelization            mmdet3d/ops/voxel/voxelize.py
    DynamicScatter, dynamic_scatter       mmdet3d/ops/voxel/scatter_points.py
    voxel_layer (4 functions)             mmdet3d/ops/voxel/src/voxelization.cpp
    HardSimpleVFE, DynamicSimpleVFE       mmdet3d/models/voxel_encoders/voxel_encoder.py
    backproject_depth_to_points           plugin ReconstructionBackbone._backproject_depth_to_points
    DepthToVoxels                         the fused, batched path (no reference equivalent)
    pack_sparse_inputs                    batched SparseEncoder inputs (sparse_refinement.py:393-402)
    PillarDecorator, PointPillarsScatter  mmdet3d/models/voxel_encoders/pillar_encoder.py, middle_encoders/pillar_scatter.py

All compute runs in librd3_b200.so (hand-written sm_100a CUDA, C ABI in
include/rd3_b200.h).  There is no CPU or PyTorch fallback.
"""
__version__ = "0.1.0"

from . import voxel_layer  # noqa: F401
from .voxelize import Voxelization, voxelization  # noqa: F401
from .scatter_points import DynamicScatter, dynamic_scatter  # noqa: F401
from .voxel_encoder import (DynamicSimpleVFE, HardSimpleVFE, HardVoxelOccupancyVFE, SoftVoxelOccupancyVFE,  # noqa: F401
                            hard_simple_vfe, voxel_occupancy)
from .backproject import (DepthToPointsMixin, backproject_depth_to_points,  # noqa: F401
                          conf_threshold, unproject_padded)
from .fused import DepthToVoxels, pack_sparse_inputs  # noqa: F401
from .parallel import gather_flat_outputs, gather_voxel_outputs, shard_range, shard_sizes  # noqa: F401
from .patch import patch_mmdet3d  # noqa: F401
from .pipelines import FilterPointByRange, VoxelDownsample  # noqa: F401
from .pillar import (PillarDecorator, PointPillarsScatter, map_voxel_center_to_point,  # noqa: F401
                     pillar_decorate)
