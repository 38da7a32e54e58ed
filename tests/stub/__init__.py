"""Test-only stand-in for libmizuroute_b200.so built on the CPU oracle (mr_stub.c), and a copy of the stand-alone host
linked against it.  Lets the CPU suite run the host program end to end (forcing ingest, history and restart files);
the product host `mizuroute_b200/route_runoff` never sees it.  Outputs go to tests/stub/_build/ (git-ignored)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libmr_hoststub.so")
HOST = os.path.join(OUT, "route_runoff_stub")
_DEPS = [os.path.join(HERE, "mr_stub.c"), os.path.join(ROOT, "oracle", "mr_oracle.c"), os.path.join(ROOT, "include", "mizuroute_b200.h"),
         os.path.join(ROOT, "mizuroute_b200", "csrc", "route_runoff.cpp"), os.path.join(ROOT, "mizuroute_b200", "csrc", "nc3.h")]


def build() -> str:
    if os.path.exists(HOST) and os.path.exists(LIB) and all(os.path.getmtime(d) <= min(os.path.getmtime(HOST), os.path.getmtime(LIB)) for d in _DEPS):
        return HOST
    os.makedirs(OUT, exist_ok=True)
    for cmd in (["gcc", "-O3", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fPIC", "-w", "-shared", "-o", LIB, _DEPS[0], "-lm"],
                ["g++", "-O2", "-std=c++17", "-o", HOST, _DEPS[3], "-L" + OUT, "-lmr_hoststub", "-Wl,-rpath,$ORIGIN"]):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("building the host stub failed:\n" + r.stdout + r.stderr)
    return HOST
