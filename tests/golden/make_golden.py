"""Generates the golden fixtures in this directory from the CPU oracle.

    python tests/golden/make_golden.py [--options-only]      (--options-only: rewrite only the OPTION_CASES fixtures)

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

# Fixtures of the options that hang off the per-reach loop: the Euler schemes, water management (is_flux_wm: flux rows stored with
# the fixture) and data assimilation (qmodOption 1: gauge rows + record flags stored with the fixture)
OPTION_CASES = {
    "tree60_hourly_345": dict(case=dict(kind="random", n=60, seed=5, dt=3600.0, route_opt="345", steps=30)),
    "conus300_hourly_135_wm": dict(case=dict(kind="conus", n=300, seed=4, dt=3600.0, route_opt="135", steps=16), wm=True),
    "conus300_hourly_0145_da": dict(case=dict(kind="conus", n=300, seed=4, dt=3600.0, route_opt="0145", steps=20), da=(4, 2)),   # qBlendPeriod, QerrTrend
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


def option_inputs(name):
    """(net, params, opts, runoff, flux_wm or None, (obs, has_record) or None) of an OPTION_CASES fixture, seeded."""
    from tests.util import gauge_series
    spec = OPTION_CASES[name]
    net, params, opts, ro = case(**spec["case"])
    K = ro.shape[0]
    flux = obs = None
    if spec.get("wm"):
        rng = np.random.default_rng(17)
        flux = np.full((K, net.nRch), -9999.0)
        pick = rng.random((K, net.nRch)) < 0.3
        flux[pick] = rng.choice([-1.0, 1.0], pick.sum()) * rng.lognormal(np.log(0.02), 1.5, pick.sum())
    if spec.get("da"):
        base = Oracle(net, params, opts).run(ro)[1]
        o, has, _ = gauge_series(net, K, seed=23, base=base)
        obs = (o, has)
    return net, params, opts, ro, flux, obs


def run_option_case(name, net, params, opts, ro, flux, obs):
    o = Oracle(net, params, opts)
    if obs is not None:
        o.set_da(1, *OPTION_CASES[name]["da"])
    q = np.empty((len(opts.route_opt), ro.shape[0], net.nRch))
    for t in range(ro.shape[0]):
        if flux is not None:
            o.set_wm(flux[t])
        if obs is not None:
            o.set_obs(obs[0][t] if obs[1][t] else None)
        o.step(ro[t])
        for i, c in enumerate(opts.route_opt):
            q[i, t] = o.get(0, int(c))
    return q


def build_option(name):
    net, params, opts, ro, flux, obs = option_inputs(name)
    out = {"runoff": ro, "q": run_option_case(name, net, params, opts, ro, flux, obs)}
    if flux is not None:
        out["flux_wm"] = flux
    if obs is not None:
        out["obs"], out["has_record"] = obs
    for f in NET_FIELDS:
        v = getattr(net, f)
        if v is not None:
            out["net_" + f] = v
    return out


if __name__ == "__main__":
    only_options = "--options-only" in sys.argv          # leave the fixtures of CASES as they are
    for name in ([] if only_options else CASES):
        d = build(name)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, {k: v.shape for k, v in d.items() if k in ("runoff", "q")})
    for name in OPTION_CASES:
        d = build_option(name)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, d["q"].shape)
