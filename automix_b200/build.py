"""In-tree build of libautomix.so (CUDA kernels + C-ABI + the drop-in C API) for sm_100a.

nvcc cross-compiles without a GPU.  Objects go to automix_b200/_obj, the library to
automix_b200/lib/libautomix.so (git-ignored, but shipped to the GPU box by gpurun).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INC = os.path.join(ROOT, "include")
OBJ = os.path.join(HERE, "_obj")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libautomix.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CC = os.environ.get("CC", "gcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVFLAGS = [f"-D{d}" for d in os.environ.get("AMX_NVCC_DEFS", "").split()] + ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-I", INC, "-I", CSRC] + ARCH
CFLAGS = ["-O2", "-fPIC", "-Wall", "-I", INC, "-I", CSRC]


def _sources():
    cu = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu") and f != "amx_plugin_tu.cu")  # (the plug-in TU is built per plug-in)
    c = sorted(f for f in os.listdir(CSRC) if f.endswith(".c"))
    return cu, c


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    return h.hexdigest()


def _compile(src):
    out = os.path.join(OBJ, os.path.basename(src) + ".o")
    log = out + ".log"
    if src.endswith(".cu"):
        cmd = [NVCC] + NVFLAGS + ["-c", src, "-o", out]
    else:
        cmd = [CC] + CFLAGS + ["-c", src, "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"compile failed: {src}\n{r.stdout}\n{r.stderr}")
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    cu, c = _sources()
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(INC, f) for f in os.listdir(INC)]
    deps.append(os.path.abspath(__file__))
    stamp = os.path.join(OBJ, "stamp")
    dig = _digest(deps)
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    srcs = [os.path.join(CSRC, f) for f in cu + c]
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(_compile, srcs))
    cmd = [NVCC, "-shared", "-o", LIB] + ARCH + objs + ["-Xlinker", "-soname=libautomix.so", "-lcudart_static", "-lpthread", "-ldl", "-lrt", "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(dig)
    if verbose:
        for o in objs:
            print(open(o + ".log").read())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
