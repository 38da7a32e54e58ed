"""Parity at the sizes BASELINE.json names (GPU): the exact C2 configuration against the oracle from a cold start, and the
3 M-reach CONUS-like network of C3 / C4 / C5 (KWT, KWT + IRF, lakes) for a few steps continuing from the GPU's own spun-up
state, which is copied into the oracle (restart schema, mr_get_state) -- the way bench.py samples its CPU baseline.
Tolerances: north_star's 1e-6 (IRF) and 1e-4 (KWT), relative."""
import numpy as np
import pytest

from tests.util import IRF_RTOL, KWT_RTOL, rel_err

pytestmark = pytest.mark.gpu


def test_c2_binary_tree_100k_irf_hourly_240_steps():
    """BASELINE.json configs[1]: 100 000-reach binary tree, IRF only, hourly, 240 steps in one batch, every reach and step."""
    from mizuroute_b200 import synth
    from mizuroute_b200.network import RouteOptions, RouteParams
    from mizuroute_b200.route import Router
    from oracle.oracle import Oracle
    net = synth.binary_tree(100_000, seed=2)
    opts = RouteOptions(dt=3600.0, route_opt="1", runoffMin=1e-15)
    ro = synth.runoff_series(net, 240, seed=11, dt=3600.0)
    qo = Oracle(net, RouteParams(), opts, n_threads=8).run(ro)
    qg = Router(net, RouteParams(), opts, max_batch=240).route_batch(ro)
    assert np.isfinite(qg).all()
    assert rel_err(qg[0], qo[0]) <= IRF_RTOL
    assert np.array_equal(qg[0], qo[0])              # (the kernels keep the oracle's operation order: equal to the last bit)


@pytest.mark.parametrize("workload,route_opt,dt,lakes,spin", [("C3", "2", 86400.0, 0, 64), ("C4", "12", 3600.0, 0, 96), ("C5", "2", 86400.0, 10_000, 64)])
def test_three_million_reaches_continue_like_the_oracle(workload, route_opt, dt, lakes, spin):
    """BASELINE.json configs[2..4] on one GPU: `spin` steps on the device, state handed to the oracle, the next two steps of all
    3 M reaches compared (wide confluences, thinning, wave breaking and, for C5, lakes and their outlet reaches included)."""
    from mizuroute_b200 import synth
    from mizuroute_b200.network import RouteOptions, RouteParams
    from mizuroute_b200.route import Router
    from oracle.oracle import Oracle, seed_oracle_from_router
    net = synth.conus_like(3_000_000, seed=3)
    opts = RouteOptions(dt=dt, route_opt=route_opt, runoffMin=1e-15)
    if lakes:
        synth.add_lakes(net, lakes, np.random.default_rng(103))
        opts.is_lake_sim = True
        opts.LakeInputOption = 1
    ro = synth.runoff_series(net, spin + 2, seed=11, dt=dt)
    r = Router(net, RouteParams(), opts, max_batch=spin)
    r.route_batch(np.ascontiguousarray(ro[:spin]), want_q=False)
    o = Oracle(net, RouteParams(), opts, n_threads=8)
    seed_oracle_from_router(o, r)
    qo = o.run(ro[spin:])
    qg = r.route_batch(np.ascontiguousarray(ro[spin:]))
    for i, c in enumerate(route_opt):
        assert np.isfinite(qg[i]).all()
        assert rel_err(qg[i], qo[i]) <= (KWT_RTOL if c == "2" else IRF_RTOL), (workload, c)
