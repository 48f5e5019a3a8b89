"""Builds libb200seed.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels
with the repository snapshot to the GPU box)."""
from __future__ import annotations

import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "csrc", "b200seed_api.cu")
DEPS = [SRC, os.path.join(_HERE, "csrc", "seed_kernels.cuh"), os.path.join(_HERE, "csrc", "seed_math.cuh"), os.path.join(_HERE, "csrc", "seed_lanes.cuh"),
        os.path.join(_HERE, "csrc", "seed_tile.cuh"),
        os.path.join(_HERE, "..", "include", "b200seed.h")]
OUT = os.path.join(_HERE, "libb200seed.so")

# -fmad=false + IEEE div/sqrt (nvcc defaults) + no flush-to-zero: the cut arithmetic must
# round like the reference's CPU build (see csrc/seed_math.cuh).
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
              "-Xcompiler", "-fPIC,-ffp-contract=off,-O2", "-shared", "-cudart", "static"]


def build(force: bool = False, verbose: bool = False, out: str = OUT, defines=()) -> str:
    """out / defines: experiment builds (e.g. out=build/variant.so, defines=["B200_X=1"])."""
    if not force and os.path.exists(out) and all(
            os.path.getmtime(out) >= os.path.getmtime(d) for d in DEPS):
        return out
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    tmp = out + ".tmp"   # built aside and moved into place: a snapshot never sees a partial file
    cmd = ([nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) +
           [f"-D{d}" for d in defines] + ["-o", tmp, SRC])
    subprocess.check_call(cmd)
    os.replace(tmp, out)
    return out


if __name__ == "__main__":
    import sys
    if len(sys.argv) > 1:   # python build.py out.so DEFINE[=v] ...
        print(build(force=True, verbose=True, out=os.path.abspath(sys.argv[1]), defines=sys.argv[2:]))
    else:
        print(build(force=True, verbose=True))
