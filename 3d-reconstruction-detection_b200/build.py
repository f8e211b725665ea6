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


SASS_HASH = os.path.join(HERE, "librd3_b200.sass_md5")


def _write_sass_hash(objs):
    """md5 of the machine code (cuobjdump -sass without addresses / line info): the stamp that ties numbers measured
    offline (ncu traffic in profiles/) to a build.  Comment-only edits of the sources leave it unchanged."""
    import hashlib
    import re
    cuobjdump = os.path.join(os.path.dirname(os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")), "cuobjdump")
    h = hashlib.md5()
    try:
        for o in objs:
            out = subprocess.run([cuobjdump, "-sass", o], capture_output=True, text=True, check=True).stdout
            for line in out.splitlines():
                m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?);", line)
                if m:
                    h.update(m.group(1).encode())
                elif "Function :" in line:
                    h.update(line.strip().encode())
        with open(SASS_HASH, "w") as f:
            f.write(h.hexdigest()[:16] + "\n")
    except Exception:                                  # noqa: BLE001  (no cuobjdump: bench.py falls back to a source hash)
        if os.path.exists(SASS_HASH):
            os.remove(SASS_HASH)


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
    _write_sass_hash(objs)
    if verbose:
        sys.stderr.write("\n".join(logs))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
