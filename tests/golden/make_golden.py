"""Generates the golden fixtures in this directory from the CPU oracle.

    python tests/golden/make_golden.py

The reference (Fortran) cannot be built or run in this environment and ships no golden vectors for the routing
path, so these fixtures are produced by oracle/mr_oracle.c (cross-checked against the
independent Python twin).  They pin REGRESSIONS of the oracle and of the CUDA path -- not the Fortran: parity
stays "unpinned" in the sense of DESIGN.md section 2.  Inputs are stored with the outputs so a fixture is
self-contained."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.oracle import Oracle  # noqa: E402
from tests.util import case  # noqa: E402

CASES = {
    # name: kwargs of tests.util.case
    "tree60_hourly_012": dict(kind="random", n=60, seed=5, dt=3600.0, route_opt="012", steps=36),
    "tree60_daily_012": dict(kind="random", n=60, seed=6, dt=86400.0, route_opt="012", steps=20),
    "conus400_lakes_daily_12": dict(kind="conus", n=400, seed=4, dt=86400.0, route_opt="12", steps=20, lakes=6),
    "tree80_zero_area_hourly_12": dict(kind="random", n=80, seed=9, dt=3600.0, route_opt="12", steps=24, zero_area_frac=0.15),
    "binary127_hourly_1_hwtop": dict(kind="binary", n=127, seed=2, dt=3600.0, route_opt="1", steps=30, hw_drain_point=1),
}

NET_FIELDS = ["segId", "downSegId", "length", "slope", "hruId", "hruSegId", "area", "islake", "lakeModelType",
              "D03_MaxStorage", "D03_Coefficient", "D03_Power", "D03_S0"]


def build(name):
    net, params, opts, ro = case(**CASES[name])
    o = Oracle(net, params, opts)
    q = o.run(ro)
    st = o.get_state()
    out = {"runoff": ro, "q": q, "frac_future": o.frac_future()}
    for f in NET_FIELDS:
        v = getattr(net, f)
        if v is not None:
            out["net_" + f] = v
    for k in ("qfuture", "basin_qr1", "irf_qfuture", "irf_vol", "kwt_n"):
        if k in st:
            out["state_" + k] = st[k]
    return out


if __name__ == "__main__":
    for name in CASES:
        d = build(name)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, {k: v.shape for k, v in d.items() if k in ("runoff", "q")})
