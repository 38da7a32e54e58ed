"""Pins the oracle (through its golden fixtures) to the reference's own Fortran where that can be compiled:
oracle/ref_build builds oracle/_ref/ref_replay from /root/reference with gfortran and replays the fixtures' inputs through
kwt_rch / irf_rch / accum_inst_runoff / IRF_route_basin.  Skipped without a Fortran compiler or the reference checkout --
which is the case in the build image and on the GPU boxes, so parity stays "unpinned" there."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/route/build/src"
CASES = ["tree60_hourly_012", "tree60_daily_012", "tree80_zero_area_hourly_12", "binary127_hourly_1_hwtop"]
HAVE = shutil.which("gfortran") is not None and os.path.isdir(REF)


@pytest.fixture(scope="module")
def replay():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle", "ref_build")])
    return os.path.join(ROOT, "oracle", "_ref", "ref_replay")


@pytest.mark.skipif(not HAVE, reason="needs gfortran and /root/reference (oracle/ref_build/README.md)")
@pytest.mark.parametrize("name", CASES)
def test_reference_fortran_reproduces_the_golden_fixture(replay, name, tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_build"))
    from export_fixture import export
    fx, out = str(tmp_path / "fx.txt"), str(tmp_path / "q.txt")
    q, methods = export(name, fx)
    subprocess.check_call([replay, fx, out])
    got = np.loadtxt(out).reshape(q.shape[1], len(methods), q.shape[2]).transpose(1, 0, 2)
    for i, m in enumerate(methods):
        tol = 1e-4 if m == 2 else 1e-6
        err = np.max(np.abs(got[i] - q[i]) / np.maximum(np.abs(q[i]), 1e-300))
        assert err <= tol, (name, m, err)


@pytest.mark.parametrize("name", CASES)
def test_fixture_export_is_a_consistent_network(name, tmp_path):
    """The exporter itself (runs everywhere): processing order puts upstream reaches first, indices are 1-based and in range,
    the per-step rows have one value per reach."""
    sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_build"))
    from export_fixture import export
    fx = str(tmp_path / "fx.txt")
    q, methods = export(name, fx)
    lines = open(fx).read().splitlines()
    n, _, steps, nr = (int(v) for v in lines[0].split())
    assert [int(v) for v in lines[2].split()] == methods and len(methods) == nr
    k, seen = 3, 0
    for i in range(1, n + 1):
        f = lines[k].split(); k += 1
        down, nups = int(f[1]), int(f[8])
        assert down == 0 or i < down <= n
        for _ in range(nups):
            u = int(lines[k].split()[0]); k += 1
            assert 1 <= u < i
            seen += 1
    assert len(lines) == k + steps and all(len(lines[k + t].split()) == n for t in range(steps))
    assert q.shape == (nr, steps, n)
