"""Builds the C++ host mirror (libteeline_host.so) and the `teeline` CLI stand-in with g++,
linked against the in-tree libteeline_cuda.so (rpath $ORIGIN so the pair is relocatable)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
HOST = os.path.join(HERE, "host")
LIB = os.path.join(HERE, "libteeline_host.so")
CLI = os.path.join(HERE, "teeline")
SELFTEST = os.path.join(HERE, "host_selftest")
CXXFLAGS = ["-O2", "-std=c++17", "-fPIC", "-Wall", "-Wextra", "-ffp-contract=off"]


def _stale(target: str, deps) -> bool:
    return not os.path.exists(target) or any(os.path.getmtime(d) > os.path.getmtime(target) for d in deps)


def build(force: bool = False) -> str:
    hdrs = [os.path.join(HOST, "teeline_host.hpp"), os.path.join(HERE, "..", "include", "teeline_cuda.h")]
    cuda_lib = os.path.join(HERE, "libteeline_cuda.so")
    cxx = os.environ.get("CXX", "g++")
    src = os.path.join(HOST, "teeline_host.cpp")
    if force or _stale(LIB, [src, cuda_lib] + hdrs):
        cmd = [cxx] + CXXFLAGS + ["-shared", "-o", LIB, src, "-L" + HERE, "-lteeline_cuda", "-Wl,-rpath,$ORIGIN"]
        print("[teeline_b200] " + " ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
    cli_src = os.path.join(HOST, "teeline_cli.cpp")
    if force or _stale(CLI, [cli_src, LIB] + hdrs):
        cmd = [cxx] + CXXFLAGS + ["-o", CLI, cli_src, "-L" + HERE, "-lteeline_host", "-lteeline_cuda", "-Wl,-rpath,$ORIGIN"]
        print("[teeline_b200] " + " ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
    st_src = os.path.join(HOST, "host_selftest.cpp")
    if force or _stale(SELFTEST, [st_src, LIB] + hdrs):
        cmd = [cxx] + CXXFLAGS + ["-o", SELFTEST, st_src, "-L" + HERE, "-lteeline_host", "-lteeline_cuda", "-Wl,-rpath,$ORIGIN"]
        print("[teeline_b200] " + " ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
    return CLI


if __name__ == "__main__":
    build(force=True)
