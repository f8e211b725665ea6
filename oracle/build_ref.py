"""Build recipe for oracle/_ref: the REFERENCE's own CPU voxel op, compiled unmodified.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this.

Compiles, from the sources where they lie under /root/reference (never copied
into this repo):

    mmdetection3d/mmdet3d/ops/voxel/src/voxelization.cpp      (pybind: 4 symbols, :6-11)
    mmdetection3d/mmdet3d/ops/voxel/src/voxelization_cpu.cpp  (hard/dynamic voxelize, :107-171)
    mmdetection3d/mmdet3d/ops/voxel/src/scatter_points_cpu.cpp

into oracle/_ref/ref_voxel_layer.so (CPU only, no -DWITH_CUDA).  The output
directory is git-ignored but NOT gpurun-ignored, so the prebuilt module travels
to the GPU box, where /root/reference does not exist.

The build uses torch.utils.cpp_extension (ninja + g++) directly on those three
files; the reference's own build system (setup.py) is not run.
"""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/mmdetection3d/mmdet3d/ops/voxel/src"
OUT_DIR = os.path.join(HERE, "_ref")
MOD_NAME = "ref_voxel_layer"
SOURCES = ["voxelization.cpp", "voxelization_cpu.cpp", "scatter_points_cpu.cpp"]


def so_path():
    return os.path.join(OUT_DIR, MOD_NAME + ".so")


def build(verbose=False):
    """Compile oracle/_ref/ref_voxel_layer.so if the reference tree is present.

    Returns the path of the .so, or None when /root/reference is absent and no
    prebuilt module exists.
    """
    if os.path.exists(so_path()):
        return so_path()
    if not os.path.isdir(REF_SRC):
        return None
    os.makedirs(OUT_DIR, exist_ok=True)
    from torch.utils.cpp_extension import load
    load(name=MOD_NAME,
         sources=[os.path.join(REF_SRC, s) for s in SOURCES],
         extra_cflags=["-O3"],
         build_directory=OUT_DIR,
         is_python_module=True,
         verbose=verbose)
    return so_path() if os.path.exists(so_path()) else None


def load_ref():
    """Import the prebuilt reference module (or None if it was never built)."""
    p = so_path()
    if not os.path.exists(p):
        return None
    import torch  # noqa: F401  (the .so links against libtorch)
    if MOD_NAME in sys.modules:
        return sys.modules[MOD_NAME]
    spec = importlib.util.spec_from_file_location(MOD_NAME, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[MOD_NAME] = mod
    return mod


if __name__ == "__main__":
    print(build(verbose=True))
