"""The Euler routing schemes as the GPU runs them (mizuroute_b200/csrc/mr_euler.cuh), compiled for the host and stepped
reach by reach in stage order, against the CPU oracle: REACH_Q, REACH_VOL(1) and the molecules must agree BIT FOR BIT
(the host build uses libm's pow like the oracle; on the device pow may differ in the last ulp, which the GPU tests
allow for with a tolerance)."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as orc
from oracle.oracle import Oracle
from tests import emul
from tests.util import case, gauge_series

CASES = [
    dict(kind="random", n=80, seed=5, dt=3600.0, steps=40, zero_area_frac=0.08),
    dict(kind="random", n=60, seed=6, dt=86400.0, steps=20),                      # Muskingum-Cunge sub-steps
    dict(kind="conus", n=800, seed=4, dt=3600.0, steps=24),
    dict(kind="random", n=50, seed=7, dt=900.0, steps=30, hw_drain_point=1),
    dict(kind="binary", n=127, seed=2, dt=10800.0, steps=20, floodplain=True),    # over-bank branch
    dict(kind="random", n=40, seed=8, dt=3600.0, steps=12, min_length_route=1500.0),   # pass-through reaches
    dict(kind="tiny:one_reach", dt=3600.0, steps=20), dict(kind="tiny:isolated_reaches", dt=86400.0, steps=10),     # degenerate networks
    dict(kind="tiny:chain_of_two", dt=3600.0, steps=20), dict(kind="tiny:middle_reach_without_hru", dt=86400.0, steps=12),
    dict(kind="tiny:star_of_five", dt=900.0, steps=20),
]


def _run(kw, method, noise_seed=0):
    kw = dict(kw)
    fp = kw.pop("floodplain", False)
    net, params, opts, ro = case(route_opt=str(method), **kw)
    opts.floodplain = fp
    K = ro.shape[0]
    o = Oracle(net, params, opts)
    qr = np.empty((K + 1, net.nRch)); qo = np.empty((K, net.nRch))
    qr[0] = o.get(orc.F_BASIN_QR1)
    for t in range(K):
        o.step(ro[t])
        qr[t + 1] = o.get(orc.F_BASIN_QR1)
        qo[t] = o.get(orc.F_REACH_Q, method)
    L = emul.load_euler(noise_seed)
    nm = orc.N_MOLECULE[method]
    qe = np.empty((K, net.nRch)); ve = np.empty(net.nRch); me = np.empty((net.nRch, nm))
    msg = C.create_string_buffer(256)
    p = lambda a, ct: a.ctypes.data_as(C.POINTER(ct))
    ierr = L.euler_emul_run(C.c_int(method), C.c_int(net.nRch), C.c_int(net.nHRU), p(net.segId, C.c_int), p(net.downSegId, C.c_int),
                            p(net.hruSegId, C.c_int), p(net.area, C.c_double), p(net.length, C.c_double), p(net.slope, C.c_double),
                            C.c_double(params.mann_n), C.c_double(params.wscale), C.c_double(opts.dt), C.c_int(opts.hw_drain_point),
                            C.c_double(opts.min_length_route), C.c_int(int(fp)), C.c_int(K), p(qr, C.c_double), None, p(qe, C.c_double),
                            p(ve, C.c_double), p(me, C.c_double), msg)
    assert ierr == 0, msg.value.decode()
    return o, qo, qe, ve, me


@pytest.mark.parametrize("method", [orc.M_KW, orc.M_MC, orc.M_DW], ids=["kw", "mc", "dw"])
@pytest.mark.parametrize("kw", CASES, ids=lambda k: "%s-%g" % (k["kind"], k["dt"]))
def test_device_source_matches_oracle_bit_for_bit(kw, method):
    o, qo, qe, ve, me = _run(kw, method)
    assert np.array_equal(qe, qo)
    assert np.array_equal(ve, o.get(orc.F_REACH_VOL1, method))
    assert np.array_equal(me, o.molecule(method))


@pytest.mark.parametrize("method", [orc.M_KW, orc.M_MC, orc.M_DW], ids=["kw", "mc", "dw"])
def test_last_ulp_of_pow_moves_discharge_by_less_than_the_gpu_tolerance(method):
    """Conditioning of the GPU parity bar.  flow_depth stops its Newton iteration at a 0.5 % change (hydraulic.f90:35), so
    a last-ulp difference in pow() (device vs libm) can add or drop one iteration and move the depth by ~1e-5 relative.
    With pow() perturbed by +-1 ulp the host build of the device code stays within 1e-5 of the oracle on every case;
    the GPU tests hold the Euler schemes to 1e-4."""
    worst = 0.0
    for kw in CASES:
        o, qo, qe, ve, me = _run(kw, method, noise_seed=1)
        worst = max(worst, float(np.max(np.abs(qe - qo) / np.maximum(np.abs(qo), 1e-300))))
    assert worst < 1e-5, worst


@pytest.mark.parametrize("method", [orc.M_KW, orc.M_MC, orc.M_DW], ids=["kw", "mc", "dw"])
def test_dry_channels_then_a_flood_pulse(method):
    """runoffMin = 0 and no runoff at first: discharge is exactly zero (flow_depth returns 0 below Qmin = 1e-50, celerity 0,
    Muskingum-Cunge takes its Qbar <= Qmin branch), then a pulse 1e6 times the usual runoff arrives; no NaN/Inf, the
    device source still equals the oracle bit for bit and the twin agrees."""
    from oracle.twin import Twin
    net, params, opts, ro = case("random", n=40, seed=9, dt=3600.0, route_opt=str(method), steps=30)
    opts.runoffMin = 0.0
    ro = ro.copy(); ro[:6] = 0.0; ro[12:14] *= 1.0e6; ro[20:] = 0.0
    kw = dict(kind="random", n=40, seed=9, dt=3600.0, steps=30)

    def run_with(ro_):
        o = Oracle(net, params, opts)
        K = ro_.shape[0]
        qr = np.empty((K + 1, net.nRch)); qo = np.empty((K, net.nRch))
        qr[0] = o.get(orc.F_BASIN_QR1)
        for t in range(K):
            o.step(ro_[t]); qr[t + 1] = o.get(orc.F_BASIN_QR1); qo[t] = o.get(orc.F_REACH_Q, method)
        return o, qr, qo
    o, qr, qo = run_with(ro)
    assert np.isfinite(qo).all() and (qo >= 0.0).all()
    assert not qo[:1].any() and qo[12:16].max() > 1.0                     # dry start, then the pulse
    L = emul.load_euler()
    K = ro.shape[0]
    qe = np.empty((K, net.nRch)); ve = np.empty(net.nRch); me = np.empty((net.nRch, orc.N_MOLECULE[method]))
    msg = C.create_string_buffer(256)
    p = lambda a, ct: a.ctypes.data_as(C.POINTER(ct))
    ierr = L.euler_emul_run(C.c_int(method), C.c_int(net.nRch), C.c_int(net.nHRU), p(net.segId, C.c_int), p(net.downSegId, C.c_int),
                            p(net.hruSegId, C.c_int), p(net.area, C.c_double), p(net.length, C.c_double), p(net.slope, C.c_double),
                            C.c_double(params.mann_n), C.c_double(params.wscale), C.c_double(opts.dt), C.c_int(opts.hw_drain_point),
                            C.c_double(opts.min_length_route), C.c_int(0), C.c_int(K), p(qr, C.c_double), None, p(qe, C.c_double),
                            p(ve, C.c_double), p(me, C.c_double), msg)
    assert ierr == 0, msg.value.decode()
    assert np.array_equal(qe, qo) and np.array_equal(me, o.molecule(method))
    t = Twin(net, params, opts)
    for k in range(K):
        t.step(ro[k])
    assert np.array_equal(np.array(t.Q[method]), qo[-1])


@pytest.mark.parametrize("method", [orc.M_KW, orc.M_MC, orc.M_DW], ids=["kw", "mc", "dw"])
def test_water_management_cascade_in_the_euler_schemes(method):
    """The abstraction / injection cascade of kw_dw_reach<M, EXT> / mc_reach<EXT> (kwe_route.f90:118-146 and siblings) against
    Oracle.set_wm, bit for bit -- REACH_Q, REACH_VOL(1) and the molecules."""
    net, params, opts, ro = case("conus", n=700, seed=3, dt=3600.0, route_opt=str(method), steps=24)
    K = ro.shape[0]
    rng = np.random.default_rng(5)
    flux = np.full((K, net.nRch), -9999.0)
    pick = rng.random((K, net.nRch)) < 0.4
    flux[pick] = rng.choice([-1.0, 1.0], pick.sum()) * rng.lognormal(np.log(0.02), 1.5, pick.sum())
    o = Oracle(net, params, opts)
    qr = np.empty((K + 1, net.nRch)); qo = np.empty((K, net.nRch))
    qr[0] = o.get(orc.F_BASIN_QR1)
    for t in range(K):
        o.set_wm(flux[t])
        o.step(ro[t])
        qr[t + 1] = o.get(orc.F_BASIN_QR1); qo[t] = o.get(orc.F_REACH_Q, method)
    L = emul.load_euler()
    nm = orc.N_MOLECULE[method]
    qe = np.empty((K, net.nRch)); ve = np.empty(net.nRch); me = np.empty((net.nRch, nm))
    msg = C.create_string_buffer(256)
    p = lambda a, ct: np.ascontiguousarray(a).ctypes.data_as(C.POINTER(ct))
    ierr = L.euler_emul_run(C.c_int(method), C.c_int(net.nRch), C.c_int(net.nHRU), p(net.segId, C.c_int), p(net.downSegId, C.c_int),
                            p(net.hruSegId, C.c_int), p(net.area, C.c_double), p(net.length, C.c_double), p(net.slope, C.c_double),
                            C.c_double(params.mann_n), C.c_double(params.wscale), C.c_double(opts.dt), C.c_int(opts.hw_drain_point),
                            C.c_double(opts.min_length_route), C.c_int(0), C.c_int(K), p(qr, C.c_double), p(flux, C.c_double),
                            p(qe, C.c_double), p(ve, C.c_double), p(me, C.c_double), msg)
    assert ierr == 0, msg.value.decode()
    assert np.array_equal(qe, qo)
    assert np.array_equal(ve, o.get(orc.F_REACH_VOL1, method))
    assert np.array_equal(me, o.molecule(method))
    assert not np.array_equal(qo, Oracle(net, params, opts).run(ro)[0])


@pytest.mark.parametrize("trend", [1, 2, 3, 4])
@pytest.mark.parametrize("method", [orc.M_KW, orc.M_MC, orc.M_DW], ids=["kw", "mc", "dw"])
def test_direct_insertion_in_the_euler_schemes(method, trend):
    """kw_dw_reach<M, EXT> / mc_reach<EXT> ending in direct_insertion (kwe_route.f90:183-197 and siblings) on the rows da_rows
    builds from the gauge records -- against Oracle.set_da / set_obs, bit for bit: REACH_Q, REACH_VOL(1), molecules, Qerror."""
    net, params, opts, ro = case("conus", n=500, seed=3, dt=3600.0, route_opt=str(method), steps=24)
    K = ro.shape[0]
    base = Oracle(net, params, opts).run(ro)[0]
    obs, has, gauges = gauge_series(net, K, seed=10 + trend, base=base)
    blend = 5
    o = Oracle(net, params, opts); o.set_da(1, blend, trend)
    qr = np.empty((K + 1, net.nRch)); qo = np.empty((K, net.nRch))
    qr[0] = o.get(orc.F_BASIN_QR1)
    for t in range(K):
        o.set_obs(obs[t] if has[t] else None)
        o.step(ro[t])
        qr[t + 1] = o.get(orc.F_BASIN_QR1); qo[t] = o.get(orc.F_REACH_Q, method)
    L = emul.load_euler()
    nm = orc.N_MOLECULE[method]
    qe = np.empty((K, net.nRch)); ve = np.empty(net.nRch); me = np.empty((net.nRch, nm)); qerr = np.empty(net.nRch)
    msg = C.create_string_buffer(256)
    p = lambda a, ct: None if a is None else np.ascontiguousarray(a).ctypes.data_as(C.POINTER(ct))
    L.euler_emul_set_da(C.c_int(blend), C.c_int(trend), p(has, C.c_int), p(obs, C.c_double), p(qerr, C.c_double))
    ierr = L.euler_emul_run(C.c_int(method), C.c_int(net.nRch), C.c_int(net.nHRU), p(net.segId, C.c_int), p(net.downSegId, C.c_int),
                            p(net.hruSegId, C.c_int), p(net.area, C.c_double), p(net.length, C.c_double), p(net.slope, C.c_double),
                            C.c_double(params.mann_n), C.c_double(params.wscale), C.c_double(opts.dt), C.c_int(opts.hw_drain_point),
                            C.c_double(opts.min_length_route), C.c_int(0), C.c_int(K), p(qr, C.c_double), None,
                            p(qe, C.c_double), p(ve, C.c_double), p(me, C.c_double), msg)
    assert ierr == 0, msg.value.decode()
    assert np.array_equal(qe, qo)
    assert np.array_equal(ve, o.get(orc.F_REACH_VOL1, method))
    assert np.array_equal(me, o.molecule(method))
    assert np.array_equal(qerr, o.get(orc.F_QERROR, method))
    assert not np.array_equal(qo, base)
