"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np

from mizuroute_b200 import synth
from mizuroute_b200.network import RouteOptions, RouteParams

IRF_RTOL = 1.0e-6     # north_star tolerance, IRF (double precision)
KWT_RTOL = 1.0e-4     # north_star tolerance, KWT


def rel_err(a, b, floor=1e-300):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0


def case(kind="random", n=60, seed=5, dt=3600.0, route_opt="012", steps=24, zero_area_frac=0.0, lakes=0, **kw):
    if kind == "random":
        net = synth.random_tree(n, seed=seed, zero_area_frac=zero_area_frac)
    elif kind == "binary":
        net = synth.binary_tree(n, seed=seed)
    else:
        net = synth.conus_like(n, seed=seed, n_lakes=0)
    opts = RouteOptions(dt=dt, route_opt=route_opt, runoffMin=1e-15, **kw)
    if lakes:
        synth.add_lakes(net, lakes, np.random.default_rng(seed + 100))
        opts.is_lake_sim = True
        opts.LakeInputOption = 1
    ro = synth.runoff_series(net, steps, seed=seed + 1, dt=dt)
    return net, RouteParams(), opts, ro
