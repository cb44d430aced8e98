"""In-tree build of the CUDA extension (libdrs_b200.so) for sm_100a.

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the
GPU box with the gpurun snapshot.  `python -m deeprecsys_b200.build` rebuilds.
"""
from __future__ import absolute_import, division, print_function

import hashlib
import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libdrs_b200.so")
STAMP = LIB_PATH + ".stamp"
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), "include")

# host-only translation units (g++: GCC vector extensions + per-ISA clones, which nvcc's front
# end does not take); linked into the same .so
HOST_SOURCES = ["hostpack.cpp", "wire.cpp"]
HOST_FLAGS = ["-O3", "-std=c++17", "-fPIC", "-Wall", "-pthread"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
    "-Xcompiler", "-Wall",
]


def _sources():
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)
                   if f.endswith((".cu", ".cuh", ".h", ".cpp")))
    files.append(os.path.join(INCLUDE, "drs_b200.h"))
    return files


def _digest():
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS + HOST_FLAGS).encode())
    for path in _sources():
        with open(path, "rb") as f:
            h.update(os.path.basename(path).encode())   # location-independent: the tree is copied
            h.update(f.read())                          # to a scratch path on the GPU box
    return h.hexdigest()


def needs_build():
    if not os.path.exists(LIB_PATH) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != _digest()


def build(force=False, verbose=False):
    """Compile csrc/engine.cu -> lib/libdrs_b200.so if sources changed."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(LIB_DIR, exist_ok=True)
    objs = []
    for src in HOST_SOURCES:
        obj = os.path.join(LIB_DIR, os.path.splitext(src)[0] + ".o")
        hcmd = [os.environ.get("CXX", "g++")] + HOST_FLAGS + ["-I", INCLUDE, "-c",
                                                              os.path.join(CSRC, src), "-o", obj]
        hp = subprocess.run(hcmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if verbose or hp.returncode != 0:
            sys.stderr.write(hp.stdout)
        if hp.returncode != 0:
            raise RuntimeError("g++ failed (%d): %s" % (hp.returncode, " ".join(hcmd)))
        objs.append(obj)
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [
        "-I", INCLUDE, "-o", LIB_PATH, os.path.join(CSRC, "engine.cu")] + objs
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed (%d): %s" % (proc.returncode, " ".join(cmd)))
    with open(STAMP, "w") as f:
        f.write(_digest())
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
