"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the CPU oracle).

CPU: the oracle still reproduces every fixture bit for bit (regression pin of the checker itself).
GPU: the CUDA path, through the C ABI, reproduces them: SUM/IRF exactly, KWT within 1e-4 relative.
OPTION_CASES: the Euler schemes, water management and data assimilation, with their per-step inputs stored in the fixture."""
import os

import numpy as np
import pytest

from mizuroute_b200.network import RiverNetwork, RouteOptions, RouteParams
from tests.golden.make_golden import CASES, NET_FIELDS, OPTION_CASES, run_option_case
from tests.util import IRF_RTOL, KWT_RTOL, rel_err

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    z = np.load(os.path.join(HERE, name + ".npz"))
    kw = CASES[name]
    net = RiverNetwork(**{f: (z["net_" + f] if "net_" + f in z.files else None) for f in NET_FIELDS})
    extra = {k: v for k, v in kw.items() if k not in ("kind", "n", "seed", "dt", "route_opt", "steps", "zero_area_frac", "lakes")}
    opts = RouteOptions(dt=kw["dt"], route_opt=kw["route_opt"], runoffMin=1e-15, **extra)
    if kw.get("lakes"):
        opts.is_lake_sim = True
        opts.LakeInputOption = 1
    return z, net, RouteParams(), opts


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden(name):
    from oracle.oracle import Oracle
    z, net, params, opts = load(name)
    o = Oracle(net, params, opts)
    q = o.run(z["runoff"])
    assert np.array_equal(q, z["q"])
    assert np.array_equal(o.frac_future(), z["frac_future"])
    st = o.get_state()
    for k in ("qfuture", "irf_qfuture", "kwt_n"):
        if "state_" + k in z.files:
            assert np.array_equal(st[k], z["state_" + k]), k


@pytest.mark.gpu
@pytest.mark.parametrize("batch", [1, 7])
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_reproduces_golden(name, batch):
    from mizuroute_b200 import capi
    from mizuroute_b200.route import Router
    z, net, params, opts = load(name)
    ro = z["runoff"]
    r = Router(net, params, opts, max_batch=batch)
    q = np.concatenate([r.route_batch(np.ascontiguousarray(ro[s:s + batch])) for s in range(0, ro.shape[0], batch)], axis=1)
    for i, c in enumerate(opts.route_opt):
        if c == "2":
            assert rel_err(q[i], z["q"][i]) <= KWT_RTOL
        elif opts.is_lake_sim and c != "0":        # the Doll-2003 release calls pow(): device and libm differ in the last ulp
            assert rel_err(q[i], z["q"][i]) <= IRF_RTOL
        else:
            assert np.array_equal(q[i], z["q"][i]), f"method {c} expected bit-identical"
    assert np.array_equal(r.basin_uh(), z["frac_future"])
    if "2" in opts.route_opt:
        assert np.array_equal(r.get_state(capi.ST_KWT_NWAVE), z["state_kwt_n"])
    assert rel_err(r.get_state(capi.ST_BASIN_QFUTURE), z["state_qfuture"], 1e-30) <= IRF_RTOL


def load_option(name):
    z = np.load(os.path.join(HERE, name + ".npz"))
    kw = OPTION_CASES[name]["case"]
    net = RiverNetwork(**{f: (z["net_" + f] if "net_" + f in z.files else None) for f in NET_FIELDS})
    opts = RouteOptions(dt=kw["dt"], route_opt=kw["route_opt"], runoffMin=1e-15)
    flux = z["flux_wm"] if "flux_wm" in z.files else None
    obs = (z["obs"], z["has_record"]) if "obs" in z.files else None
    return z, net, RouteParams(), opts, flux, obs


@pytest.mark.parametrize("name", sorted(OPTION_CASES))
def test_oracle_reproduces_option_golden(name):
    """Euler schemes, water management, data assimilation: the oracle reproduces the committed discharge bit for bit."""
    z, net, params, opts, flux, obs = load_option(name)
    assert np.array_equal(run_option_case(name, net, params, opts, z["runoff"], flux, obs), z["q"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(OPTION_CASES))
def test_cuda_reproduces_option_golden(name):
    from mizuroute_b200.route import Router
    z, net, params, opts, flux, obs = load_option(name)
    ro, batch = z["runoff"], 6
    r = Router(net, params, opts, max_batch=batch)
    if obs is not None:
        r.set_da(1, *OPTION_CASES[name]["da"])
    parts = []
    for s in range(0, ro.shape[0], batch):
        if flux is not None:
            r.upload_wm(flux[s:s + batch])
        if obs is not None:
            r.upload_obs(obs[0][s:s + batch], obs[1][s:s + batch])
        parts.append(r.route_batch(np.ascontiguousarray(ro[s:s + batch])))
    q = np.concatenate(parts, axis=1)
    for i, c in enumerate(opts.route_opt):
        if c == "0":
            assert np.array_equal(q[i], z["q"][i])
        else:                                      # IRF 1e-6 (bit-identical without options), Euler schemes 1e-4 (pow / Newton stop, test_schemes_gpu.py)
            assert rel_err(q[i], z["q"][i], floor=1e-9) <= (IRF_RTOL if c == "1" else 1e-4), c
