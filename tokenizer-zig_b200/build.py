"""Builds libtokzig_b200.so / .a IN-TREE for sm_100a (nvcc cross-compiles without a GPU).

The shared library travels to the GPU box with the repo snapshot; nothing is JIT-compiled at run time.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(HERE, "_lib")
SO = os.path.join(LIB_DIR, "libtokzig_b200.so")
AR = os.path.join(LIB_DIR, "libtokzig_b200.a")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]

CU_SRCS = [os.path.join(HERE, "csrc", "tkz_api.cu")]
CXX_SRCS = [os.path.join(HERE, "host", "tokzig_host.cpp"), os.path.join(HERE, "host", "tokzig_multi.cpp")]


def _deps():
    out = []
    for d in ("csrc", "host"):
        for f in os.listdir(os.path.join(HERE, d)):
            out.append(os.path.join(HERE, d, f))
    out.append(os.path.join(os.path.dirname(HERE), "include", "tokzig_b200.h"))
    return out


def needs_build():
    if not (os.path.exists(SO) and os.path.exists(AR)):
        return True
    t = min(os.path.getmtime(SO), os.path.getmtime(AR))
    return any(os.path.getmtime(f) > t for f in _deps())


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    os.makedirs(LIB_DIR, exist_ok=True)
    objs = []
    for src in CU_SRCS:
        obj = os.path.join(LIB_DIR, os.path.basename(src) + ".o")
        cmd = [NVCC, *ARCH, "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", *os.environ.get("TKZ_NVCC_FLAGS", "").split(), "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.check_call(cmd)
        objs.append(obj)
    for src in CXX_SRCS:
        obj = os.path.join(LIB_DIR, os.path.basename(src) + ".o")
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-Wall", "-pthread", "-c", src, "-o", obj])
        objs.append(obj)
    # the CUDA runtime is linked dynamically (the image and torch both ship libcudart.so.12): nothing of the runtime is embedded in the artefact
    subprocess.check_call([NVCC, *ARCH, "-shared", "-o", SO, *objs, "-cudart", "shared", "-Xlinker", "-rpath=/usr/local/cuda/lib64"])
    if os.path.exists(AR):
        os.remove(AR)
    subprocess.check_call(["ar", "rcs", AR, *objs])
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
