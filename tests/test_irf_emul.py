"""Runoff accumulation and IRF routing as the GPU runs them (mizuroute_b200/csrc/mr_irf.cuh: ring-buffered QFUTURE_IRF in
slot-major arrays, reach unit hydrographs from mr_uh.h, lakes via mr_lake.cuh), compiled for the host and stepped in stage
order, against the CPU oracle: REACH_Q of both methods, REACH_VOL(1), the water balance and the future-flow series in the
restart layout must agree BIT FOR BIT (the GPU tests find the same on the device for SUM / IRF)."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as orc
from oracle.oracle import Oracle
from tests import emul
from tests.util import case, gauge_series

CASES = [
    dict(kind="random", n=120, seed=5, dt=3600.0, steps=40, zero_area_frac=0.1),
    dict(kind="random", n=80, seed=6, dt=86400.0, steps=20),
    dict(kind="conus", n=1500, seed=4, dt=3600.0, steps=30),
    dict(kind="binary", n=255, seed=2, dt=900.0, steps=30, hw_drain_point=1),
    dict(kind="random", n=60, seed=8, dt=3600.0, steps=12, min_length_route=1500.0),
    dict(kind="conus", n=900, seed=4, dt=86400.0, steps=14, lakes=9),
    dict(kind="tiny:one_reach", dt=3600.0, steps=20), dict(kind="tiny:isolated_reaches", dt=86400.0, steps=10),
    dict(kind="tiny:chain_of_two", dt=3600.0, steps=20), dict(kind="tiny:middle_reach_without_hru", dt=3600.0, steps=20),
    dict(kind="tiny:star_of_five", dt=900.0, steps=20),
]


@pytest.mark.parametrize("kw", CASES, ids=lambda k: "%s-%g%s" % (k["kind"], k["dt"], "-lakes" if k.get("lakes") else ""))
def test_sum_and_irf_device_source_match_oracle_bit_for_bit(kw):
    net, params, opts, ro = case(route_opt="01", **kw)
    K = ro.shape[0]
    o = Oracle(net, params, opts)
    qr = np.empty((K + 1, net.nRch)); qo = np.empty((2, K, net.nRch))
    qr[0] = o.get(orc.F_BASIN_QR1)
    for t in range(K):
        o.step(ro[t])
        qr[t + 1] = o.get(orc.F_BASIN_QR1)
        qo[0, t] = o.get(orc.F_REACH_Q, orc.M_SUM); qo[1, t] = o.get(orc.F_REACH_Q, orc.M_IRF)
    L = emul.load_irf()
    qs = np.empty((K, net.nRch)); qi = np.empty((K, net.nRch)); ve = np.empty(net.nRch); we = np.empty(net.nRch)
    qf = np.empty((net.nRch, 240)); mx = C.c_int(0)
    msg = C.create_string_buffer(256)
    p = lambda a, ct: None if a is None else np.ascontiguousarray(a).ctypes.data_as(C.POINTER(ct))
    lake = opts.is_lake_sim and net.islake is not None
    zeros = np.zeros(net.nRch)
    ierr = L.irf_emul_run(C.c_int(net.nRch), C.c_int(net.nHRU), p(net.segId, C.c_int), p(net.downSegId, C.c_int), p(net.hruSegId, C.c_int),
                          p(net.area, C.c_double), p(net.length, C.c_double), p(net.slope, C.c_double),
                          p(net.islake if lake else None, C.c_int), p(net.lakeModelType if lake else None, C.c_int),
                          p(net.D03_MaxStorage if lake else zeros, C.c_double), p(net.D03_Coefficient if lake else zeros, C.c_double),
                          p(net.D03_Power if lake else zeros, C.c_double), p(net.D03_S0 if lake else zeros, C.c_double),
                          C.c_double(params.wscale), C.c_double(opts.dt), C.c_double(params.velo), C.c_double(params.diff),
                          C.c_int(opts.hw_drain_point), C.c_double(opts.min_length_route), C.c_int(opts.LakeInputOption), C.c_int(K),
                          p(qr, C.c_double), None, None, None, C.c_int(0), p(qs, C.c_double), p(qi, C.c_double), p(ve, C.c_double), p(we, C.c_double), p(qf, C.c_double),
                          C.byref(mx), msg)
    assert ierr == 0, msg.value.decode()
    assert np.array_equal(qs, qo[0])
    assert np.array_equal(qi, qo[1])
    assert np.array_equal(ve, o.get(orc.F_REACH_VOL1, orc.M_IRF))
    assert np.array_equal(we, o.get(orc.F_WB, orc.M_IRF))
    ptr, _ = o.reach_uh()
    flat = np.empty(int(ptr[-1])); orc.lib().mro_get_qfuture_irf(o.h, flat.ctypes.data_as(C.POINTER(C.c_double)))
    assert mx.value == int(np.diff(ptr).max())
    for r in range(net.nRch):
        assert np.array_equal(qf[r, :ptr[r + 1] - ptr[r]], flat[ptr[r]:ptr[r + 1]]), r


@pytest.mark.parametrize("lakes", [0, 9])
def test_water_management_device_source_matches_oracle(lakes):
    """mr_upload_wm path: the abstraction cascade of irf_reach<EXT> and, with lakes, the lake fluxes, the target volumes of
    the lakes flagged LakeTargVol and the volume jump start in lake_reach<M, EXT> -- against Oracle.set_wm, bit for bit."""
    net, params, opts, ro = case("conus", n=900, seed=4, dt=86400.0, route_opt="01", steps=12, lakes=lakes)
    K = ro.shape[0]
    rng = np.random.default_rng(11)
    lake = bool(lakes)
    targ = None
    if lake:
        lk = np.flatnonzero(net.islake == 1)
        net.lake_params = {"LakeTargVol": np.isin(np.arange(net.nRch), lk[:3]).astype(np.float64)}
        targ = net.lake_params["LakeTargVol"]
    flux = np.full((K, net.nRch), -9999.0)
    pick = rng.random((K, net.nRch)) < 0.4
    flux[pick] = rng.choice([-1.0, 1.0], pick.sum()) * rng.lognormal(np.log(0.05), 1.5, pick.sum())
    vol = np.where(net.islake == 1, rng.uniform(1e6, 5e7, (K, net.nRch)), 0.0) if lake else None
    o = Oracle(net, params, opts)
    qr = np.empty((K + 1, net.nRch)); qo = np.empty((K, net.nRch))
    qr[0] = o.get(orc.F_BASIN_QR1)
    for t in range(K):
        o.set_wm(flux[t], None if vol is None else vol[t], vol_jumpstart=True)
        o.step(ro[t])
        qr[t + 1] = o.get(orc.F_BASIN_QR1); qo[t] = o.get(orc.F_REACH_Q, orc.M_IRF)
    L = emul.load_irf()
    qs = np.empty((K, net.nRch)); qi = np.empty((K, net.nRch)); ve = np.empty(net.nRch); we = np.empty(net.nRch)
    qf = np.empty((net.nRch, 240)); mx = C.c_int(0)
    msg = C.create_string_buffer(256)
    p = lambda a, ct: None if a is None else np.ascontiguousarray(a).ctypes.data_as(C.POINTER(ct))
    zeros = np.zeros(net.nRch)
    ierr = L.irf_emul_run(C.c_int(net.nRch), C.c_int(net.nHRU), p(net.segId, C.c_int), p(net.downSegId, C.c_int), p(net.hruSegId, C.c_int),
                          p(net.area, C.c_double), p(net.length, C.c_double), p(net.slope, C.c_double),
                          p(net.islake if lake else None, C.c_int), p(net.lakeModelType if lake else None, C.c_int),
                          p(net.D03_MaxStorage if lake else zeros, C.c_double), p(net.D03_Coefficient if lake else zeros, C.c_double),
                          p(net.D03_Power if lake else zeros, C.c_double), p(net.D03_S0 if lake else zeros, C.c_double),
                          C.c_double(params.wscale), C.c_double(opts.dt), C.c_double(params.velo), C.c_double(params.diff),
                          C.c_int(opts.hw_drain_point), C.c_double(opts.min_length_route), C.c_int(opts.LakeInputOption), C.c_int(K),
                          p(qr, C.c_double), p(flux, C.c_double), p(vol, C.c_double), p(targ, C.c_double), C.c_int(1),
                          p(qs, C.c_double), p(qi, C.c_double), p(ve, C.c_double), p(we, C.c_double), p(qf, C.c_double), C.byref(mx), msg)
    assert ierr == 0, msg.value.decode()
    assert np.array_equal(qi, qo)
    assert np.array_equal(ve, o.get(orc.F_REACH_VOL1, orc.M_IRF))
    assert np.array_equal(we, o.get(orc.F_WB, orc.M_IRF))
    assert (qo[:, (net.islake != 1) if lake else slice(None)] == 0.0).any()          # somewhere more was asked for than there was


@pytest.mark.parametrize("trend", [1, 2, 3, 4])
def test_direct_insertion_device_source_matches_oracle(trend):
    """mr_set_da / mr_upload_obs path: da_rows (the body of k_da_rows) turns the gauge records of a batch into the Qobs /
    Qelapsed every (reach, step) sees, irf_reach<EXT> ends in direct_insertion (mr_dev.h) instead of the water balance --
    against Oracle.set_da / set_obs, bit for bit: REACH_Q, REACH_VOL(1), Qerror, the future-flow series."""
    net, params, opts, ro = case("conus", n=900, seed=4, dt=86400.0, route_opt="01", steps=24)
    K = ro.shape[0]
    base = Oracle(net, params, opts).run(ro)[1]
    obs, has, gauges = gauge_series(net, K, seed=trend, base=base)
    blend = 4
    o = Oracle(net, params, opts); o.set_da(1, blend, trend)
    qr = np.empty((K + 1, net.nRch)); qo = np.empty((K, net.nRch))
    qr[0] = o.get(orc.F_BASIN_QR1)
    for t in range(K):
        o.set_obs(obs[t] if has[t] else None)
        o.step(ro[t])
        qr[t + 1] = o.get(orc.F_BASIN_QR1); qo[t] = o.get(orc.F_REACH_Q, orc.M_IRF)
    L = emul.load_irf()
    qs = np.empty((K, net.nRch)); qi = np.empty((K, net.nRch)); ve = np.empty(net.nRch); we = np.empty(net.nRch); qerr = np.empty(net.nRch)
    qf = np.empty((net.nRch, 240)); mx = C.c_int(0)
    msg = C.create_string_buffer(256)
    p = lambda a, ct: None if a is None else np.ascontiguousarray(a).ctypes.data_as(C.POINTER(ct))
    zeros = np.zeros(net.nRch)
    L.irf_emul_set_da(C.c_int(blend), C.c_int(trend), p(has, C.c_int), p(obs, C.c_double), p(qerr, C.c_double))
    ierr = L.irf_emul_run(C.c_int(net.nRch), C.c_int(net.nHRU), p(net.segId, C.c_int), p(net.downSegId, C.c_int), p(net.hruSegId, C.c_int),
                          p(net.area, C.c_double), p(net.length, C.c_double), p(net.slope, C.c_double), None, None,
                          p(zeros, C.c_double), p(zeros, C.c_double), p(zeros, C.c_double), p(zeros, C.c_double),
                          C.c_double(params.wscale), C.c_double(opts.dt), C.c_double(params.velo), C.c_double(params.diff),
                          C.c_int(opts.hw_drain_point), C.c_double(opts.min_length_route), C.c_int(opts.LakeInputOption), C.c_int(K),
                          p(qr, C.c_double), None, None, None, C.c_int(0),
                          p(qs, C.c_double), p(qi, C.c_double), p(ve, C.c_double), p(we, C.c_double), p(qf, C.c_double), C.byref(mx), msg)
    assert ierr == 0, msg.value.decode()
    assert np.array_equal(qi, qo)
    assert np.array_equal(ve, o.get(orc.F_REACH_VOL1, orc.M_IRF))
    assert np.array_equal(qerr, o.get(orc.F_QERROR, orc.M_IRF))
    assert (we == 0.0).all()                                   # the water balance is not evaluated under qmodOption 1
    assert not np.array_equal(qo, base)                        # and the observations did change the flow
    assert np.array_equal(qs, Oracle(net, params, opts).run(ro)[0])      # runoff accumulation is not corrected
