"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

SUM / IRF / hillslope UH follow the oracle's operation order without FMA contraction, so they are expected to
agree to the last bit; the asserted bound is the north_star tolerance (1e-6 relative).  KWT calls pow(), whose
device implementation differs from libm in the last ulp, so it is held to 1e-4 relative."""
import numpy as np
import pytest

from tests.util import IRF_RTOL, KWT_RTOL, case, rel_err

pytestmark = pytest.mark.gpu


def _run_both(net, params, opts, ro, batch):
    from mizuroute_b200.route import Router
    from oracle.oracle import Oracle
    K = ro.shape[0]
    o = Oracle(net, params, opts)
    qo = o.run(ro)
    r = Router(net, params, opts, max_batch=max(batch, 8))
    parts = []
    for s in range(0, K, batch):
        parts.append(r.route_batch(np.ascontiguousarray(ro[s:s + batch])))
    qg = np.concatenate(parts, axis=1)
    return o, r, qo, qg


def _assert_q(opts, qo, qg):
    for i, c in enumerate(opts.route_opt):
        tol = KWT_RTOL if c == "2" else IRF_RTOL
        e = rel_err(qg[i], qo[i])
        assert e <= tol, f"method {c}: rel err {e:.3e} > {tol}"


@pytest.mark.parametrize("batch", [1, 5, 24])
@pytest.mark.parametrize("dt", [3600.0, 86400.0])
def test_small_tree_all_methods(batch, dt):
    net, params, opts, ro = case("random", n=80, seed=7, dt=dt, route_opt="012", steps=24)
    o, r, qo, qg = _run_both(net, params, opts, ro, batch)
    _assert_q(opts, qo, qg)


def test_irf_bit_exact_and_fluxes():
    from mizuroute_b200 import capi
    from oracle import oracle as orc
    net, params, opts, ro = case("random", n=300, seed=11, dt=3600.0, route_opt="01", steps=30, zero_area_frac=0.1)
    o, r, qo, qg = _run_both(net, params, opts, ro, 7)
    assert np.array_equal(qg, qo), "SUM/IRF expected bit-identical to the oracle"
    for f_g, f_o in [(capi.REACH_VOL1, orc.F_REACH_VOL1), (capi.REACH_INFLOW, orc.F_REACH_INFLOW), (capi.WB, orc.F_WB),
                     (capi.BASIN_QR1, orc.F_BASIN_QR1), (capi.BASIN_QI, orc.F_BASIN_QI), (capi.REACH_VOL0, orc.F_REACH_VOL0)]:
        assert rel_err(r.flux(f_g, 1), o.get(f_o, 1), floor=1e-30) <= IRF_RTOL
    for f_g, f_o in [(capi.R_WIDTH, orc.F_WIDTH), (capi.TOTAREA, orc.F_TOTAREA), (capi.BASAREA, orc.F_BASAREA)]:
        assert np.array_equal(r.flux(f_g, 1), o.get(f_o, 1))


def test_unit_hydrographs_match_oracle():
    net, params, opts, ro = case("random", n=120, seed=3, dt=3600.0, route_opt="1", steps=1)
    from mizuroute_b200.route import Router
    from oracle.oracle import Oracle
    o = Oracle(net, params, opts)
    r = Router(net, params, opts)
    assert np.array_equal(r.basin_uh(), o.frac_future())
    ptr, val = o.reach_uh()
    ntdh, uh = r.reach_uh()
    assert np.array_equal(ntdh, np.diff(ptr))
    for i in range(net.nRch):
        assert np.array_equal(uh[i, :ntdh[i]], val[ptr[i]:ptr[i + 1]])


@pytest.mark.parametrize("kind,n", [("binary", 4095), ("conus", 20000)])
def test_medium_networks(kind, n):
    net, params, opts, ro = case(kind, n=n, seed=2, dt=86400.0, route_opt="12", steps=40)
    o, r, qo, qg = _run_both(net, params, opts, ro, 16)
    _assert_q(opts, qo, qg)


def test_large_hourly_network_all_methods():
    """50 k reaches, hourly, 36 steps in batches of 12: deep wavefronts, thinning on every large river, shocks, and
    tasks that spill to the full-capacity scratch -- all against the oracle."""
    net, params, opts, ro = case("conus", n=50000, seed=3, dt=3600.0, route_opt="012", steps=36)
    o, r, qo, qg = _run_both(net, params, opts, ro, 12)
    _assert_q(opts, qo, qg)
    assert np.array_equal(qg[0], qo[0]) and np.array_equal(qg[1], qo[1])


def test_kwt_thinning_and_shocks_exercised():
    """Hourly steps on a tree with wide confluences build >20-particle merges (remove_rch) and shocks."""
    from oracle import oracle as orc
    net, params, opts, ro = case("random", n=200, seed=21, dt=3600.0, route_opt="2", steps=60)
    orc.lib().mro_reset_counters()
    o, r, qo, qg = _run_both(net, params, opts, ro, 20)
    assert orc.lib().mro_counter(0) > 0 and orc.lib().mro_counter(1) > 0
    _assert_q(opts, qo, qg)


def test_wide_confluence_uses_full_scratch():
    """12 interior arms into one reach: more merged particles than the shared-memory scratch holds, so the task is
    re-run with the arena scratch (k_route_kwt); also a 12-way confluence for SUM and IRF."""
    from tests.util import star_network
    net, params, opts, ro = star_network(route_opt="012")
    o, r, qo, qg = _run_both(net, params, opts, ro, 8)
    _assert_q(opts, qo, qg)


def test_lakes():
    net, params, opts, ro = case("conus", n=5000, seed=4, dt=86400.0, route_opt="12", steps=30, lakes=40)
    assert net.islake.sum() > 0
    o, r, qo, qg = _run_both(net, params, opts, ro, 10)
    _assert_q(opts, qo, qg)


def test_state_roundtrip_and_restart():
    """get_state after n steps == oracle state; a fresh handle restarted from it continues identically (ERS idea)."""
    from mizuroute_b200 import capi
    from mizuroute_b200.route import Router
    net, params, opts, ro = case("random", n=150, seed=9, dt=3600.0, route_opt="12", steps=20)
    o, r, qo, qg = _run_both(net, params, opts, ro[:12], 6)
    so = o.get_state()
    assert rel_err(r.get_state(capi.ST_BASIN_QFUTURE), so["qfuture"], 1e-30) <= IRF_RTOL
    ptr, _ = o.reach_uh()
    irf = r.get_state(capi.ST_IRF_QFUTURE)
    for i in range(net.nRch):
        assert rel_err(irf[i, :ptr[i + 1] - ptr[i]], so["irf_qfuture"][ptr[i]:ptr[i + 1]], 1e-30) <= IRF_RTOL
    assert np.array_equal(r.get_state(capi.ST_KWT_NWAVE), so["kwt_n"])
    # restart a second handle from the first one's state
    r2 = Router(net, params, opts, max_batch=8)
    r2.set_steps_done(12)
    for v in (capi.ST_BASIN_QFUTURE, capi.ST_BASIN_QR, capi.ST_IRF_QFUTURE, capi.ST_IRF_VOL, capi.ST_KWT_NWAVE,
              capi.ST_KWT_ROUTED, capi.ST_KWT_QWAVE, capi.ST_KWT_TENTRY, capi.ST_KWT_TEXIT):
        r2.set_state(v, r.get_state(v))
    qa = r.route_batch(np.ascontiguousarray(ro[12:20]))
    qb = r2.route_batch(np.ascontiguousarray(ro[12:20]))
    assert np.array_equal(qa, qb)


def test_step_api_equals_batch_api():
    from mizuroute_b200 import capi
    from mizuroute_b200.route import Router
    net, params, opts, ro = case("random", n=100, seed=13, dt=86400.0, route_opt="012", steps=10)
    a = Router(net, params, opts, max_batch=1)
    b = Router(net, params, opts, max_batch=10)
    qb = b.route_batch(ro)
    for k in range(10):
        a.main_route(ro[k])
        for i, m in enumerate(a.methods):
            assert np.array_equal(a.flux(capi.REACH_Q, m), qb[i, k])
    assert a.TSEC == b.TSEC


def test_async_pipeline_equals_blocking_batches():
    """mr_step_batch_async (upload / route / download overlapped over consecutive calls) returns what mr_step_batch does."""
    import torch
    from mizuroute_b200.route import Router
    net, params, opts, ro = case("conus", n=3000, seed=6, dt=3600.0, route_opt="012", steps=30)
    a = Router(net, params, opts, max_batch=10)
    b = Router(net, params, opts, max_batch=10)
    want = np.concatenate([a.route_batch(np.ascontiguousarray(ro[s:s + 10])) for s in (0, 10, 20)], axis=1)
    ins = [torch.from_numpy(np.ascontiguousarray(ro[s:s + 10])).pin_memory() for s in (0, 10, 20)]
    outs = [torch.empty((3, 10, net.nRch), dtype=torch.float64).pin_memory() for _ in range(3)]
    for i in range(3):
        b.route_batch_async(ins[i], outs[i])
    b.wait()
    got = np.concatenate([o.numpy() for o in outs], axis=1)
    assert np.array_equal(got, want)
    assert a.TSEC == b.TSEC


def test_errors_surface_as_ierr_message():
    from mizuroute_b200.route import Router, RoutingError
    net, params, opts, ro = case("random", n=50, seed=1, dt=86400.0, route_opt="1", steps=2)
    r = Router(net, params, opts, max_batch=2)
    bad = ro.copy()
    bad[1, 3] = -1.0                       # below negRunoffTol -> basin2reach error 20 (process_remap.f90:393-399)
    with pytest.raises(RoutingError) as ei:
        r.route_batch(bad)
    assert ei.value.ierr == 20 and "basin2reach" in ei.value.message
    with pytest.raises(RoutingError):
        r.route_batch(np.zeros((3, net.nHRU)))      # more steps than max_batch
    opts2 = type(opts)(dt=86400.0, route_opt="2", runoffMin=0.0)
    r2 = Router(net, params, opts2, max_batch=1)
    with pytest.raises(RoutingError) as ei:          # zero flow aborts KWT (kwt_route.f90:1365-1368)
        r2.route_batch(np.zeros((1, net.nHRU)))
    assert ei.value.ierr == 20


@pytest.mark.parametrize("n_agg,batch", [(24, 25), (5, 7), (1, 4)])
def test_history_means_on_the_device_equal_host_aggregation(n_agg, batch):
    """mr_history_means (k_history): period means of REACH_Q per method and of BASIN_QR(1) formed on the device -- sums in step
    order in double precision, mean, float32 -- equal the same aggregation of the downloaded series bit for bit, for periods
    that span batches and a last period closed by the flush (histVars_data.f90:154-246)."""
    from mizuroute_b200.route import Router
    net, params, opts, ro = case("conus", n=500, seed=9, dt=3600.0, route_opt="012", steps=61)
    K = ro.shape[0]
    r = Router(net, params, opts, max_batch=batch)
    got, qs, qrs = [], [], []
    for s in range(0, K, batch):
        nb = min(batch, K - s)
        qs.append(r.route_batch(np.ascontiguousarray(ro[s:s + nb])))
        qrs.append(r.download_basin_q(nb))
        got.append(r.history_means(nb, n_agg, want_dlay=True, flush=s + nb == K))
    got = np.concatenate(got, axis=0)
    series = np.concatenate([np.concatenate(qs, axis=1), np.concatenate(qrs, axis=0)[None]], axis=0)      # [4, K, nRch]
    want = []
    for lo in range(0, K, n_agg):
        acc = np.zeros((series.shape[0], net.nRch))
        for t in range(lo, min(lo + n_agg, K)):
            acc = acc + series[:, t]
        want.append((acc / float(min(lo + n_agg, K) - lo)).astype(np.float32))
    want = np.stack(want)
    assert got.shape == want.shape and got.dtype == np.float32
    assert np.array_equal(got, want)
