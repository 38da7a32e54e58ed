"""The KWT device code -- the thread-per-task path (mr_kwt_scalar.cuh) backed by the warp-cooperative path (mr_kwt.cuh, one
lane per team here) for the tasks it defers, and the warp-cooperative path alone -- compiled for the host, against the CPU
oracle.  Both evaluate the same operations on the same operands with libm pow(), so REACH_Q and the live
particle counts must agree bit for bit -- including steps that thin (>20 particles) and merge shocks."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as orc
from oracle.oracle import Oracle
from tests import emul
from tests.util import case


def _emul_vs_oracle(net, params, opts, ro, mode=0):
    """mode 0: scalar path first, team path for the deferred tasks (what k_route_kwt does); 1: team path only"""
    K = ro.shape[0]
    o = Oracle(net, params, opts)
    qr = np.empty((K + 1, net.nRch)); qo = np.empty((K, net.nRch))
    qr[0] = o.get(orc.F_BASIN_QR1)
    for t in range(K):
        o.step(ro[t])
        qr[t + 1] = o.get(orc.F_BASIN_QR1)
        qo[t] = o.get(orc.F_REACH_Q, orc.M_KWT)
    L = emul.load()
    L.kwt_emul_count.restype = C.c_long
    L.kwt_emul_set_mode(C.c_int(mode))
    qe = np.empty((K, net.nRch)); ne = np.empty(net.nRch, dtype=np.int32)
    msg = C.create_string_buffer(256)
    p = lambda a, ct: a.ctypes.data_as(C.POINTER(ct))
    ierr = L.kwt_emul_run(C.c_int(net.nRch), C.c_int(net.nHRU), p(net.segId, C.c_int), p(net.downSegId, C.c_int), p(net.hruSegId, C.c_int),
                          p(net.area, C.c_double), p(net.length, C.c_double), p(net.slope, C.c_double), C.c_double(params.mann_n),
                          C.c_double(params.wscale), C.c_double(opts.dt), C.c_int(K), p(qr, C.c_double), None, p(qe, C.c_double), p(ne, C.c_int), msg)
    L.kwt_emul_set_mode(C.c_int(0))
    assert ierr == 0, msg.value.decode()
    _emul_vs_oracle.retries = int(msg.value.decode().split("=")[1])
    _emul_vs_oracle.scalar, _emul_vs_oracle.deferred, _emul_vs_oracle.heavy = int(L.kwt_emul_count(0)), int(L.kwt_emul_count(1)), int(L.kwt_emul_count(2))
    return o, qo, qe, ne


@pytest.mark.parametrize("kind,n,dt,steps", [("random", 200, 3600.0, 60), ("random", 120, 86400.0, 30), ("conus", 3000, 3600.0, 48),
                                             ("binary", 1023, 86400.0, 25), ("tiny:one_reach", 1, 3600.0, 20), ("tiny:isolated_reaches", 4, 86400.0, 10),
                                             ("tiny:chain_of_two", 2, 3600.0, 25), ("tiny:middle_reach_without_hru", 3, 3600.0, 25),
                                             ("tiny:star_of_five", 6, 900.0, 30)])
@pytest.mark.parametrize("mode", [0, 1])
def test_team_kwt_bit_exact_vs_oracle(kind, n, dt, steps, mode):
    net, params, opts, ro = case(kind, n=n, seed=21, dt=dt, route_opt="2", steps=steps)
    orc.lib().mro_reset_counters()
    o, qo, qe, ne = _emul_vs_oracle(net, params, opts, ro, mode)
    if kind != "binary" and not kind.startswith("tiny:"):
        assert orc.lib().mro_counter(0) > 0, "thinning was not exercised"
        if dt < 86400.0:
            assert orc.lib().mro_counter(1) > 0, "wave breaking was not exercised"
        if mode == 0:                                        # all three paths carry a real share of the tasks
            assert _emul_vs_oracle.scalar > 0.3 * (_emul_vs_oracle.scalar + _emul_vs_oracle.heavy + _emul_vs_oracle.deferred)
            assert _emul_vs_oracle.heavy > 0 and _emul_vs_oracle.deferred > 0
    if mode == 1:
        assert _emul_vs_oracle.scalar == 0 and _emul_vs_oracle.heavy == 0
    assert np.array_equal(qe, qo)
    assert np.array_equal(ne, o.get_state()["kwt_n"])


def test_team_kwt_full_scratch_retry():
    """A star confluence (12 interior reaches draining into one) overflows the shared-memory-sized scratch and must be
    re-run with the full-capacity one, with the same result."""
    from tests.util import star_network
    net, params, opts, ro = star_network()
    o, qo, qe, ne = _emul_vs_oracle(net, params, opts, ro)
    assert _emul_vs_oracle.retries > 0, "the wide confluence did not exercise the full-capacity scratch"
    assert np.array_equal(qe, qo)


def test_team_kwt_zero_area_parents():
    net, params, opts, ro = case("random", n=150, seed=5, dt=3600.0, route_opt="2", steps=30, zero_area_frac=0.15)
    o, qo, qe, ne = _emul_vs_oracle(net, params, opts, ro)
    assert np.array_equal(qe, qo)


def test_water_management_extract_from_rch():
    """extract_from_rch in the team code (kwt_reach_team<.., EXT>, between thinning and routing) against Oracle.set_wm, bit for
    bit.  Reaches whose wave series has no point strictly inside the step make the reference's interp_rch return 0, hence
    "no water", hence zero flow and kinwav_rch's stop (kwt_route.f90:1603-1606, 421) -- they are found by trial and left out."""
    net, params, opts, ro = case("random", n=200, seed=21, dt=3600.0, route_opt="2", steps=30)
    K = ro.shape[0]
    ob = Oracle(net, params, opts)
    inflow = []
    for k in range(K):
        ob.step(ro[k]); inflow.append(ob.get(orc.F_REACH_INFLOW, orc.M_KWT))
    low = np.min(np.array(inflow)[4:], axis=0)
    rng = np.random.default_rng(2)
    want = np.where(low > 0.0, rng.uniform(-0.2, 0.2, net.nRch) * low, -9999.0)
    ok = []
    for j in np.flatnonzero(want != -9999.0)[:60]:
        one = np.full(net.nRch, -9999.0); one[j] = want[j]
        oj = Oracle(net, params, opts)
        try:
            finite = True
            for k in range(K):
                if k == 4:
                    oj.set_wm(one)
                oj.step(ro[k])
                finite = finite and bool(np.isfinite(oj.get(orc.F_REACH_Q, orc.M_KWT)).all())
            if finite:                                       # (an injection into "no water" divides by zero: inf / NaN waves)
                ok.append(j)
        except orc.OracleError:
            pass
    assert len(ok) >= 5
    L = emul.load()
    p = lambda a, ct: a.ctypes.data_as(C.POINTER(ct))
    for j in ok[:5]:
        flux = np.full((K, net.nRch), -9999.0); flux[4:, j] = want[j]
        o = Oracle(net, params, opts)
        qr = np.empty((K + 1, net.nRch)); qo = np.empty((K, net.nRch))
        qr[0] = o.get(orc.F_BASIN_QR1)
        for t in range(K):
            o.set_wm(flux[t])
            o.step(ro[t])
            qr[t + 1] = o.get(orc.F_BASIN_QR1); qo[t] = o.get(orc.F_REACH_Q, orc.M_KWT)
        qe = np.empty((K, net.nRch)); ne = np.empty(net.nRch, dtype=np.int32)
        msg = C.create_string_buffer(256)
        ierr = L.kwt_emul_run(C.c_int(net.nRch), C.c_int(net.nHRU), p(net.segId, C.c_int), p(net.downSegId, C.c_int), p(net.hruSegId, C.c_int),
                              p(net.area, C.c_double), p(net.length, C.c_double), p(net.slope, C.c_double), C.c_double(params.mann_n),
                              C.c_double(params.wscale), C.c_double(opts.dt), C.c_int(K), p(qr, C.c_double), p(flux, C.c_double), p(qe, C.c_double),
                              p(ne, C.c_int), msg)
        assert ierr == 0, msg.value.decode()
        assert np.isfinite(qo).all() and np.array_equal(qe, qo), j
        assert not np.array_equal(qo[-1], ob.get(orc.F_REACH_Q, orc.M_KWT))
