"""rd3_b200 -- B200-native depth->voxel path (unproject, voxelize, scatter, VFE).

Importable as ``rd3_b200`` (see /rd3_b200.py at the repo root).
"""
__version__ = "0.1.0"
