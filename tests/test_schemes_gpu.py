"""GPU parity of the Euler routing schemes -- <route_opt> 3 kinematic wave, 4 Muskingum-Cunge, 5 diffusive wave
(kwe_route.f90, mc_route.f90, dfw_route.f90) -- through the C ABI against the CPU oracle.

Tolerance: 1e-4 relative, as for KWT.  The arithmetic is the oracle's (tests/test_euler_emul.py holds the same source,
compiled for the host, to bit equality), but flow_depth / celerity call pow(), whose device implementation differs from
libm in the last ulp, and flow_depth's Newton iteration stops at a 0.5 % change: a last-ulp flip of that test moves the
depth by ~1e-5.  Typical agreement is ~1e-14 (tests/test_euler_emul.py measures the conditioning)."""
import numpy as np
import pytest

from tests.util import case, rel_err

pytestmark = pytest.mark.gpu
EULER_RTOL = 1e-4


def _both(net, params, opts, ro, batch):
    from tests.test_gpu_parity import _run_both
    return _run_both(net, params, opts, ro, batch)


@pytest.mark.parametrize("kw,batch", [
    (dict(kind="random", n=80, seed=5, dt=3600.0, steps=40, zero_area_frac=0.08), 1),
    (dict(kind="random", n=60, seed=6, dt=86400.0, steps=20), 7),                      # Muskingum-Cunge sub-steps
    (dict(kind="conus", n=3000, seed=4, dt=3600.0, steps=24), 24),
    (dict(kind="random", n=50, seed=7, dt=900.0, steps=30, hw_drain_point=1), 8),
    (dict(kind="binary", n=1023, seed=2, dt=10800.0, steps=20, floodplain=True), 5),    # finite bankfull depth
    (dict(kind="random", n=40, seed=8, dt=3600.0, steps=12, min_length_route=1500.0), 4),
    (dict(kind="conus", n=1500, seed=4, dt=86400.0, steps=12, lakes=12), 6),           # lake_route under the Euler schemes
], ids=["hourly", "daily-substeps", "conus3000", "hw-top", "floodplain", "pass-through", "lakes"])
def test_kw_mc_dw_match_oracle(kw, batch):
    from mizuroute_b200 import capi
    from oracle import oracle as orc
    net, params, opts, ro = case(route_opt="345", **kw)
    o, r, qo, qg = _both(net, params, opts, ro, batch)
    for i, m in enumerate((3, 4, 5)):
        assert rel_err(qg[i], qo[i]) <= EULER_RTOL, m
        assert rel_err(r.flux(capi.REACH_VOL1, m), o.get(orc.F_REACH_VOL1, m), floor=1e-6) <= EULER_RTOL
        assert rel_err(r.flux(capi.REACH_INFLOW, m), o.get(orc.F_REACH_INFLOW, m), floor=1e-30) <= EULER_RTOL
        mol = r.get_state(capi.ST_MOLECULE_KW + (m - 3))
        assert mol.shape == (net.nRch, capi.N_MOLECULE[m])
        assert rel_err(mol, o.molecule(m), floor=1e-12) <= EULER_RTOL


def test_all_six_methods_in_one_run():
    net, params, opts, ro = case("conus", n=6000, seed=7, dt=3600.0, route_opt="012345", steps=20)   # KWT well-conditioned here (test_kwt_conditioning)
    o, r, qo, qg = _both(net, params, opts, ro, 10)
    assert np.array_equal(qg[:2], qo[:2])                                   # SUM / IRF stay bit-identical
    for i in range(2, 6):
        assert rel_err(qg[i], qo[i]) <= EULER_RTOL, i


def test_molecule_state_round_trip_continues_exactly():
    """mr_get_state / mr_set_state of q_sub_kw|mc|dw + volumes: a second handle seeded with the state of the first one
    reproduces the rest of the run bit for bit (the restart path of the stand-alone host)."""
    from mizuroute_b200 import capi
    from mizuroute_b200.route import Router
    net, params, opts, ro = case("conus", n=800, seed=3, dt=3600.0, route_opt="345", steps=24)
    a = Router(net, params, opts, max_batch=12)
    a.route_batch(np.ascontiguousarray(ro[:12]))
    b = Router(net, params, opts, max_batch=12)
    b.set_steps_done(12)
    for v in (capi.ST_BASIN_QFUTURE, capi.ST_BASIN_QR, capi.ST_MOLECULE_KW, capi.ST_MOLECULE_MC, capi.ST_MOLECULE_DW, capi.ST_LAKE_VOL):
        b.set_state(v, a.get_state(v))
    b.TSEC = list(a.TSEC)
    qa = a.route_batch(np.ascontiguousarray(ro[12:]))
    qb = b.route_batch(np.ascontiguousarray(ro[12:]))
    assert np.array_equal(qa, qb)


@pytest.mark.parametrize("option,route", [(0, "14"), (2, "14"), (1, "14"), (2, "1")])
def test_lake_evaporation_and_precipitation_forcing(option, route):
    """mr_upload_lake_forcing: evaporation / precipitation through basin2reach into lake_route (LakeInputOption 0 / 2) and the
    lake water balance; two lakes run dry and their cut evaporation is seen by the method routed second (methods then run
    on one stream, in route_opt order inside every wavefront).  IRF stays within 1e-6 (Doll's outflow calls pow), MC / KWT within 1e-4."""
    from mizuroute_b200 import capi
    from mizuroute_b200.route import Router
    from oracle import oracle as orc
    from oracle.oracle import Oracle
    net, params, opts, ro = case("conus", n=900, seed=4, dt=86400.0, route_opt=route, steps=12, lakes=9)
    opts.LakeInputOption = option
    rng = np.random.default_rng(3)
    ev = np.abs(rng.lognormal(np.log(3e-5), 0.5, size=ro.shape)); pr = np.abs(rng.lognormal(np.log(2e-5), 0.8, size=ro.shape))
    lakes = np.flatnonzero(net.islake == 1)
    if option != 1:
        ev[:, np.isin(net.hruSegId, net.segId[lakes[:2]])] *= 3.0e4
    o = Oracle(net, params, opts)
    qo = o.run(ro, evapo=ev, precip=pr)
    r = Router(net, params, opts, max_batch=8)
    parts = []
    for s in range(0, ro.shape[0], 5):
        r.upload_lake_forcing(ev[s:s + 5], pr[s:s + 5])
        parts.append(r.route_batch(np.ascontiguousarray(ro[s:s + 5])))
    qg = np.concatenate(parts, axis=1)
    for i, c in enumerate(route):
        tol = 1e-6 if c == "1" else EULER_RTOL
        assert rel_err(qg[i], qo[i], floor=1e-12) <= tol, c
        assert rel_err(r.flux(capi.REACH_VOL1, int(c)), o.get(orc.F_REACH_VOL1, int(c)), floor=1e-3) <= tol
        assert rel_err(r.flux(capi.WB, int(c))[lakes], o.get(orc.F_WB, int(c))[lakes], floor=1.0) <= 1e-6
    if option != 1:
        assert (o.get(orc.F_REACH_VOL1, int(route[0]))[lakes] == 0.0).any()
    # a batch without a preceding upload uses exact zeros again, and a wrong step count is refused
    r.upload_lake_forcing(ev[:3], pr[:3])
    with pytest.raises(Exception):
        r.route_batch(np.ascontiguousarray(ro[:5]))


@pytest.mark.parametrize("calendar,start,route", [("standard", (2000, 2, 25, 0.0), "13"), ("noleap", (2001, 12, 28, 43200.0), "1")])
def test_hype_reservoirs(calendar, start, route):
    """lakeModelType 3 (HYPE): mr_set_lake_param + mr_set_sim_start; seasonal primary spillway by day of year (the runs cross
    a leap day / a year end), emergency spillway, both combination modes.  sin() and pow() on the device differ from libm in
    the last ulp: IRF is held to 1e-6 here, the kinematic wave to 1e-4."""
    from mizuroute_b200 import capi, synth
    from oracle import oracle as orc
    net, params, opts, ro = case("conus", n=900, seed=4, dt=86400.0, route_opt=route, steps=14, lakes=10)
    assert synth.make_hype_lakes(net, np.random.default_rng(5), frac=0.7) >= 2
    opts.sim_start, opts.calendar = start, calendar
    ro = ro * 30.0
    o, r, qo, qg = _both(net, params, opts, ro, 5)
    for i, c in enumerate(route):
        tol = 1e-6 if c == "1" else EULER_RTOL
        assert rel_err(qg[i], qo[i], floor=1e-6) <= tol, c
        assert rel_err(r.flux(capi.REACH_VOL1, int(c)), o.get(orc.F_REACH_VOL1, int(c)), floor=1.0) <= tol
    hy = (net.islake == 1) & (net.lakeModelType == 3)
    assert (qo[0][:, hy] > 0.0).any()


def test_hype_without_calendar_or_parameters_is_an_error():
    from mizuroute_b200 import synth
    from mizuroute_b200.route import Router, RoutingError
    net, params, opts, ro = case("conus", n=300, seed=4, dt=86400.0, route_opt="1", steps=2, lakes=6)
    synth.make_hype_lakes(net, np.random.default_rng(5), frac=0.7)
    r = Router(net, params, opts, max_batch=2)                  # no sim_start
    with pytest.raises(RoutingError, match="simulation start"):
        r.route_batch(ro)
    net.lake_params.pop("HYP_A_avg")
    with pytest.raises(RoutingError, match="HYP_A_avg"):
        Router(net, params, opts, max_batch=2)


@pytest.mark.parametrize("memory,calendar,start,dt,steps,route", [(False, "standard", (2000, 5, 20, 0.0), 86400.0, 30, "1"),
                                                                  (True, "standard", (2000, 2, 20, 0.0), 86400.0, 24, "13"),
                                                                  (True, "noleap", (2001, 12, 25, 0.0), 43200.0, 30, "35")])
def test_hanasaki_reservoirs(memory, calendar, start, dt, steps, route):
    """lakeModelType 2 (Hanasaki 2006): per-lake parameters, ring-buffer inflow memory, release coefficient reset at the start
    of the operational year; with two methods the reservoirs' state is shared, so the methods run on one stream, wavefront
    by wavefront (tests/test_lake_emul.py holds that order to bit equality on the host)."""
    from mizuroute_b200 import capi, synth
    from oracle import oracle as orc
    net, params, opts, ro = case("conus", n=900, seed=4, dt=dt, route_opt=route, steps=steps, lakes=10)
    assert synth.make_h06_lakes(net, np.random.default_rng(6), frac=0.7, memory=memory) >= 2
    opts.sim_start, opts.calendar = start, calendar
    ro = ro * 20.0
    o, r, qo, qg = _both(net, params, opts, ro, 7)
    for i, c in enumerate(route):
        tol = 1e-6 if c == "1" else EULER_RTOL
        assert rel_err(qg[i], qo[i], floor=1e-6) <= tol, c
        assert rel_err(r.flux(capi.REACH_VOL1, int(c)), o.get(orc.F_REACH_VOL1, int(c)), floor=1.0) <= tol


@pytest.mark.parametrize("route,lakes", [("1", 0), ("134", 9), ("5", 0)])
def test_water_management(route, lakes):
    """mr_upload_wm: abstraction / injection fluxes through the storage -> inflow -> lateral-flow cascade of IRF and the Euler
    schemes, lakes losing / gaining the flux, lakes flagged LakeTargVol following their target volume (jump-started)."""
    from mizuroute_b200 import capi
    from mizuroute_b200.route import Router
    from oracle import oracle as orc
    from oracle.oracle import Oracle
    net, params, opts, ro = case("conus", n=900, seed=4, dt=86400.0 if lakes else 3600.0, route_opt=route, steps=12, lakes=lakes)
    K = ro.shape[0]
    rng = np.random.default_rng(11)
    vol = None
    if lakes:
        lk = np.flatnonzero(net.islake == 1)
        net.lake_params = {"LakeTargVol": np.isin(np.arange(net.nRch), lk[:3]).astype(np.float64)}
        vol = np.where(net.islake == 1, rng.uniform(1e6, 5e7, (K, net.nRch)), 0.0)
    flux = np.full((K, net.nRch), -9999.0)
    pick = rng.random((K, net.nRch)) < 0.4
    flux[pick] = rng.choice([-1.0, 1.0], pick.sum()) * rng.lognormal(np.log(0.05), 1.5, pick.sum())
    o = Oracle(net, params, opts)
    qo = np.empty((len(route), K, net.nRch))
    for t in range(K):
        o.set_wm(flux[t], None if vol is None else vol[t], vol_jumpstart=True)
        o.step(ro[t])
        for i, c in enumerate(route):
            qo[i, t] = o.get(orc.F_REACH_Q, int(c))
    r = Router(net, params, opts, max_batch=8)
    parts = []
    for s in range(0, K, 5):
        r.upload_wm(flux[s:s + 5], None if vol is None else vol[s:s + 5], vol_jumpstart=True)
        parts.append(r.route_batch(np.ascontiguousarray(ro[s:s + 5])))
    qg = np.concatenate(parts, axis=1)
    for i, c in enumerate(route):
        tol = 1e-6 if c == "1" else EULER_RTOL
        assert rel_err(qg[i], qo[i], floor=1e-9) <= tol, c
        assert rel_err(r.flux(capi.REACH_VOL1, int(c)), o.get(orc.F_REACH_VOL1, int(c)), floor=1e-3) <= tol
        assert rel_err(r.flux(capi.WB, int(c)), o.get(orc.F_WB, int(c)), floor=1.0) <= 1e-5


def test_water_management_in_kwt():
    """extract_from_rch inside the team KWT kernel (k_route_kwt<true>): one reach at a time, as in tests/test_kwt_emul.py --
    reaches on which the reference's routine sees "no water" (and then stops in kinwav_rch) are found by trial on the oracle."""
    from mizuroute_b200.route import Router
    from oracle import oracle as orc
    from oracle.oracle import Oracle
    net, params, opts, ro = case("random", n=200, seed=21, dt=3600.0, route_opt="2", steps=30)
    K = ro.shape[0]
    ob = Oracle(net, params, opts)
    inflow = []
    for k in range(K):
        ob.step(ro[k]); inflow.append(ob.get(orc.F_REACH_INFLOW, orc.M_KWT))
    low = np.min(np.array(inflow)[4:], axis=0)
    want = np.where(low > 0.0, np.random.default_rng(2).uniform(-0.2, 0.2, net.nRch) * low, -9999.0)
    done = 0
    for j in np.flatnonzero(want != -9999.0)[:60]:
        flux = np.full((K, net.nRch), -9999.0); flux[4:, j] = want[j]
        o = Oracle(net, params, opts)
        qo = np.empty((K, net.nRch))
        try:
            for t in range(K):
                o.set_wm(flux[t]); o.step(ro[t]); qo[t] = o.get(orc.F_REACH_Q, orc.M_KWT)
        except orc.OracleError:
            continue
        if not np.isfinite(qo).all():
            continue
        r = Router(net, params, opts, max_batch=10)
        parts = []
        for s in range(0, K, 10):
            r.upload_wm(flux[s:s + 10])
            parts.append(r.route_batch(np.ascontiguousarray(ro[s:s + 10])))
        assert rel_err(np.concatenate(parts, axis=1)[0], qo) <= 1e-4, j
        done += 1
        if done == 3:
            break
    assert done == 3


@pytest.mark.parametrize("trend", [1, 3])
def test_direct_insertion(trend):
    """mr_set_da / mr_upload_obs: k_da_rows + direct_insertion in the EXT instantiations of IRF and the Euler schemes over
    several batches (one of them without an upload = a stretch without gauge records), SUM and KWT untouched; Qerror and the
    running Qobs / Qelapsed survive a state round trip into a second handle."""
    from mizuroute_b200 import capi
    from mizuroute_b200.route import Router
    from oracle import oracle as orc
    from oracle.oracle import Oracle
    from tests.util import gauge_series
    route = "012345"
    net, params, opts, ro = case("conus", n=900, seed=4, dt=3600.0, route_opt=route, steps=20)
    K = ro.shape[0]
    base = Oracle(net, params, opts).run(ro)[1]
    obs, has, gauges = gauge_series(net, K, seed=trend, base=base)
    has[10:15] = 0                                           # the batch that gets no upload at all
    blend = 4
    o = Oracle(net, params, opts); o.set_da(1, blend, trend)
    qo = np.empty((len(route), K, net.nRch))
    for t in range(K):
        o.set_obs(obs[t] if has[t] else None)
        o.step(ro[t])
        for i, c in enumerate(route):
            qo[i, t] = o.get(orc.F_REACH_Q, int(c))
    r = Router(net, params, opts, max_batch=8)
    r.set_da(1, blend, trend)
    parts = []
    for s in range(0, K, 5):
        if s == 15:                                          # continue in a fresh handle from the saved state
            r2 = Router(net, params, opts, max_batch=8); r2.set_da(1, blend, trend)
            for var in range(16):
                try:
                    r2.set_state(var, r.get_state(var))
                except RuntimeError:
                    pass
            r2.set_steps_done(s); r = r2
        if s != 10:
            r.upload_obs(obs[s:s + 5], has[s:s + 5])
        parts.append(r.route_batch(np.ascontiguousarray(ro[s:s + 5])))
    qg = np.concatenate(parts, axis=1)
    assert np.array_equal(qg[0], qo[0])                      # SUM: bit-identical and uncorrected
    for i, c in enumerate(route):
        tol = 1e-4 if c == "2" else (1e-6 if c == "1" else EULER_RTOL)
        assert rel_err(qg[i], qo[i], floor=1e-9) <= tol, c
        if c in "1345":
            assert rel_err(r.flux(capi.QERROR, int(c)), o.get(orc.F_QERROR, int(c)), floor=1e-6) <= 1e-5, c
    assert np.array_equal(r.get_state(capi.ST_DA_QELAPSED), o.qelapsed())
    assert np.array_equal(r.get_state(capi.ST_DA_QOBS), o.get(orc.F_QOBS))


@pytest.mark.parametrize("batch", [1, 7])
@pytest.mark.parametrize("name", ["one_reach", "isolated_reaches", "chain_of_two", "middle_reach_without_hru", "star_of_five"])
def test_degenerate_networks_all_six_methods(name, batch):
    """One reach (a single stage that holds only a headwater: no wavefront kernel is launched at all), isolated reaches, a chain
    of two, a reach without HRU, a star -- the sizes at which grids, stage ranges and batch tails degenerate."""
    net, params, opts, ro = case("tiny:" + name, dt=3600.0, route_opt="012345", steps=20)
    o, r, qo, qg = _both(net, params, opts, ro, batch)
    assert np.array_equal(qg[0], qo[0]) and np.array_equal(qg[1], qo[1])          # SUM, IRF: bit-identical
    for i in range(2, 6):
        assert rel_err(qg[i], qo[i], floor=1e-12) <= EULER_RTOL, opts.route_opt[i]
