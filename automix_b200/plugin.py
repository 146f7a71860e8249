"""Build a user-supplied __device__ log-posterior into a plug-in shared object.

    python -m automix_b200.plugin build my_target.cuh [-o libamx_plugin_my_target.so]

The source defines ``struct AmxUserTarget`` (see automix_b200/csrc/amx_plugin_tu.cu for the contract and
tests/plugins/toy1_user.cuh for an example); nvcc instantiates the library's own kernel templates for it
(sm_100a).  Load the result with ``amx_target_plugin`` (include/amx.h) / ``Target({"kind": "plugin", ...})``.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def build(source: str, out: str | None = None, verbose: bool = False) -> str:
    from . import build as libbuild

    libbuild.build()  # the plug-in links against libautomix.so (runtime state: stream, launch counter, error channel)
    source = os.path.abspath(source)
    name = os.path.splitext(os.path.basename(source))[0]
    out = os.path.abspath(out or os.path.join(os.path.dirname(source), f"libamx_plugin_{name}.so"))
    stamp = out + ".stamp"
    deps = [source, os.path.join(HERE, "csrc", "amx_plugin_tu.cu")] + [os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc")) if f.endswith(".cuh")]
    dig = libbuild._digest(deps)
    if os.path.exists(out) and os.path.exists(stamp) and open(stamp).read() == dig:
        return out
    wrapper = out + ".tu.cu"  # the user's path reaches the preprocessor through a generated wrapper, not a -D string
    with open(wrapper, "w") as f:
        f.write(f'#define AMX_PLUGIN_SOURCE "{source}"\n#include "amx_plugin_tu.cu"\n')
    cmd = [libbuild.NVCC, "-O3", "-std=c++17", "-lineinfo", "-shared", "-Xcompiler", "-fPIC"] + libbuild.ARCH + [
        "-I", libbuild.INC, "-I", libbuild.CSRC, wrapper, "-o", out, "-L", libbuild.LIBDIR, "-l:libautomix.so",
        "-Xlinker", "-rpath=" + libbuild.LIBDIR, "-lcudart_static", "-lpthread", "-ldl", "-lrt"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(" ".join(cmd), r.stdout, r.stderr, sep="\n")
    if r.returncode != 0:
        raise RuntimeError(f"plug-in build failed: {source}")
    with open(stamp, "w") as f:
        f.write(dig)
    return out


if __name__ == "__main__":
    if len(sys.argv) < 3 or sys.argv[1] != "build":
        sys.exit(__doc__)
    o = sys.argv[sys.argv.index("-o") + 1] if "-o" in sys.argv else None
    print(build(sys.argv[2], o, verbose="-v" in sys.argv))
