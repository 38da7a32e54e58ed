"""Builds mizuroute_b200/libmizuroute_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmizuroute_b200.so")
SOURCES = ["mr_lib.cu"]
DEPS = ["mr_lib.cu", "mr_kernels.cuh", "mr_kwt.cuh", "mr_kwt_scalar.cuh", "mr_irf.cuh", "mr_euler.cuh", "mr_lake.cuh", "mr_dev.h", "mr_lanes.h", "mr_topo.h", "mr_uh.h",
        "mr_calendar.h", "mr_lakeparams.h", "mr_ingest.h", os.path.join("..", "..", "include", "mizuroute_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",                      # a*b+c stays two roundings: bit-parity with the scalar CPU evaluation
    "-Xcompiler", "-fPIC,-fopenmp,-ffp-contract=off,-fno-fast-math",
    "-shared", "-lgomp",
]


HOST = os.path.join(HERE, "route_runoff")          # stand-alone host (control file + NetCDF-3 I/O), links the library
HOST_DEPS = ["route_runoff.cpp", "nc3.h", os.path.join("..", "..", "include", "mizuroute_b200.h")]


def host_stale() -> bool:
    if not os.path.exists(HOST):
        return True
    t = os.path.getmtime(HOST)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in HOST_DEPS) or os.path.getmtime(LIB) > t


def build_host(force: bool = False) -> str:
    build()
    if not force and not host_stale():
        return HOST
    cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-o", HOST, os.path.join(CSRC, "route_runoff.cpp"), "-L" + HERE, "-lmizuroute_b200",
           "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed building route_runoff")
    return HOST


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("MR_NVCC_EXTRA", "").split()          # development: tuning macros (-DKWS_WPB_N=2 ...)
    cmd = [nvcc, *NVCC_FLAGS, *extra, *(["-Xptxas", "-v"] if verbose else []), "-o", LIB, *[os.path.join(CSRC, s) for s in SOURCES]]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libmizuroute_b200.so")
    if verbose:
        sys.stderr.write(r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_host(force="--force" in sys.argv))
