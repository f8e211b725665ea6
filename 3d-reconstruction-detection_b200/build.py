"""Build librd3_b200.so (sm_100a only) in-tree with nvcc.

    python 3d-reconstruction-detection_b200/build.py [--force]

The library has no torch / Python dependency: it is a plain C-ABI shared object
(include/rd3_b200.h) that the Python host side drives through ctypes.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librd3_b200.so")
SOURCES = ["voxelize.cu", "depth.cu", "scatter.cu", "pillar.cu", "select.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--fmad=true", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
              "-Xptxas", "-v"]


def _newest_source():
    t = os.path.getmtime(os.path.abspath(__file__))
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            t = max(t, os.path.getmtime(os.path.join(root, f)))
    return t


def build(force=False, verbose=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest_source():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("RD3_NVCC_EXTRA", "").split()      # e.g. -DRD3_INS_MINB=6 for A/B builds
    objs = []
    logs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + extra + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        logs.append(r.stderr)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed on " + src)
        objs.append(obj)
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.check_call(cmd)
    with open(os.path.join(HERE, "csrc", "ptxas.log"), "w") as f:
        f.write("\n".join(logs))
    if verbose:
        sys.stderr.write("\n".join(logs))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
