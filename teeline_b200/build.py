"""Builds libteeline_cuda.so in-tree with nvcc for sm_100a (no torch involved).

Every .cu is compiled to its own object (in parallel, only when it or a header changed) and the
objects are linked into the shared library; kernels never call across translation units, so no
relocatable device code is needed."""
from __future__ import annotations

import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libteeline_cuda.so")
OBJ = os.path.join(HERE, "..", "build", "obj")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=false",              # belt and braces: every rounding is also spelled with _rn intrinsics
    "-prec-sqrt=true", "-prec-div=true", "-ftz=false",
    "-Xcompiler", "-fPIC",
]
LINK_FLAGS = ["-shared", "-cudart", "static"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def headers():
    return glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.hpp")) + \
        [os.path.join(HERE, "..", "include", "teeline_cuda.h")]


def _obj(src: str) -> str:
    return os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in sources() + headers())


def build(force: bool = False, verbose: bool = False, extra: list[str] | None = None, lib: str | None = None,
          obj_dir: str | None = None) -> str:
    """extra/lib/obj_dir: tuning builds (scripts/build_variant.sh) compile with extra -D flags into
    another library and object directory."""
    out = lib or LIB
    if not force and lib is None and not is_stale():
        return out
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    odir = obj_dir or OBJ
    os.makedirs(odir, exist_ok=True)
    hdr_t = max(os.path.getmtime(h) for h in headers())
    flags = NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + (extra or [])

    def compile_one(src):
        o = os.path.join(odir, os.path.basename(src)[:-3] + ".o")
        if not force and os.path.exists(o) and os.path.getmtime(o) > max(os.path.getmtime(src), hdr_t):
            return o
        cmd = [nvcc] + flags + ["-c", src, "-o", o]
        print("[teeline_b200] " + " ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
        return o

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, sources()))
    cmd = [nvcc] + NVCC_FLAGS[:2] + LINK_FLAGS + ["-o", out] + objs + ["-ldl"]
    print("[teeline_b200] " + " ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
