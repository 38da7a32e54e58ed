"""The lake reach step as the GPU runs it (mizuroute_b200/csrc/mr_lake.cuh), compiled for the host and stepped in stage
order inside a kinematic-wave network, against the CPU oracle -- bit for bit: REACH_Q, REACH_VOL(1), the lake water
balance and the evaporation left after a lake ran dry; without lake forcing (exact zeros, the path every earlier GPU lake
test takes) and with evaporation / precipitation for LakeInputOption 0, 1, 2."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as orc
from oracle.oracle import Oracle
from tests import emul
from tests.util import case


HYP_ORDER = ["HYP_E_emr", "HYP_E_lim", "HYP_E_min", "HYP_E_zero", "HYP_Qrate_emr", "HYP_Erate_emr", "HYP_Qrate_prim", "HYP_Qrate_amp",
             "HYP_Qrate_phs", "HYP_prim_F", "HYP_A_avg", "HYP_Qsim_mode"]


@pytest.mark.parametrize("option,forcing,hype", [(1, False, None), (0, False, None), (2, False, None), (0, True, None), (1, True, None), (2, True, None),
                                                 (1, False, ("standard", (2000, 2, 25, 0.0))), (2, True, ("noleap", (2001, 12, 28, 43200.0)))])
def test_lake_reach_device_source_matches_oracle(option, forcing, hype):
    net, params, opts, ro = case("conus", n=700, seed=4, dt=86400.0, route_opt="3", steps=14, lakes=9)
    opts.LakeInputOption = option
    hyp = None
    start = (0, 1, 1, 0.0)
    if hype:                                          # HYPE reservoirs + the simulation calendar (run crosses a leap day / a year end)
        from mizuroute_b200 import synth
        assert synth.make_hype_lakes(net, np.random.default_rng(5)) >= 2
        opts.calendar, opts.sim_start = hype
        start = opts.sim_start
        ro = ro * 30.0
        hyp = np.stack([net.lake_params[k] for k in HYP_ORDER])
    K = ro.shape[0]
    ev = pr = None
    if forcing:
        rng = np.random.default_rng(3)
        ev = np.abs(rng.lognormal(np.log(3e-5), 0.5, size=ro.shape)); pr = np.abs(rng.lognormal(np.log(2e-5), 0.8, size=ro.shape))
        lakes = np.flatnonzero(net.islake == 1)
        ev[:, np.isin(net.hruSegId, net.segId[lakes[:2]])] *= 3.0e4          # two lakes run dry
    o = Oracle(net, params, opts)
    qr = np.empty((K + 1, net.nRch)); qo = np.empty((K, net.nRch))
    qr[0] = o.get(orc.F_BASIN_QR1)
    for t in range(K):
        o.step(ro[t], None if ev is None else ev[t], None if pr is None else pr[t])
        qr[t + 1] = o.get(orc.F_BASIN_QR1); qo[t] = o.get(orc.F_REACH_Q, orc.M_KW)
    tc, lc = opts.conv()
    L = emul.load_lake()
    qe = np.empty((K, net.nRch)); ve = np.empty(net.nRch); we = np.empty(net.nRch); ee = np.empty(net.nRch)
    msg = C.create_string_buffer(256)
    p = lambda a, ct: None if a is None else np.ascontiguousarray(a).ctypes.data_as(C.POINTER(ct))
    arrs = [np.ascontiguousarray(a) for a in (net.islake, net.lakeModelType, net.D03_MaxStorage, net.D03_Coefficient, net.D03_Power, net.D03_S0)]
    ierr = L.lake_emul_run(C.c_int(net.nRch), C.c_int(net.nHRU), p(net.segId, C.c_int), p(net.downSegId, C.c_int), p(net.hruSegId, C.c_int),
                           p(net.area, C.c_double), p(net.length, C.c_double), p(net.slope, C.c_double), p(arrs[0], C.c_int), p(arrs[1], C.c_int),
                           p(arrs[2], C.c_double), p(arrs[3], C.c_double), p(arrs[4], C.c_double), p(arrs[5], C.c_double),
                           C.c_double(params.mann_n), C.c_double(params.wscale), C.c_double(opts.dt), C.c_int(option), C.c_double(opts.runoffMin),
                           C.c_double(tc), C.c_double(lc), C.c_int(K), p(qr, C.c_double), p(ev, C.c_double), p(pr, C.c_double),
                           p(hyp, C.c_double), C.c_int(start[0]), C.c_int(start[1]), C.c_int(start[2]), C.c_double(start[3]), C.c_int(int(opts.calendar == "noleap")),
                           p(qe, C.c_double), p(ve, C.c_double), p(we, C.c_double), p(ee, C.c_double), msg)
    assert ierr == 0, msg.value.decode()
    assert np.array_equal(qe, qo)
    assert np.array_equal(ve, o.get(orc.F_REACH_VOL1, orc.M_KW))
    assert np.array_equal(we, o.get(orc.F_WB, orc.M_KW))
    if hype:
        hy = (net.islake == 1) & (net.lakeModelType == 3)
        assert (qo[:, hy] > 0.0).any()
    if forcing:
        lk = net.islake == 1
        assert np.array_equal(ee[lk], o.lake_forcing()[0][lk])
        if option != 1 and not hype:
            assert (ve[lk] == 0.0).any()                                     # a lake did run dry


H06_ORDER = (["H06_Smax", "H06_alpha", "H06_envfact", "H06_c1", "H06_c2", "H06_exponent", "H06_denominator", "H06_c_compare", "H06_frac_Sdead",
              "H06_E_rel_ini"] + ["H06_I_" + m for m in ("Jan", "Feb", "Mar", "Apr", "May", "Jun", "Jul", "Aug", "Sep", "Oct", "Nov", "Dec")]
             + ["H06_D_" + m for m in ("Jan", "Feb", "Mar", "Apr", "May", "Jun", "Jul", "Aug", "Sep", "Oct", "Nov", "Dec")]
             + ["H06_purpose", "H06_I_mem_F", "H06_I_mem_L"])


@pytest.mark.parametrize("memory,calendar,start,dt,steps,K", [(False, "standard", (2000, 5, 20, 0.0), 86400.0, 30, 7),
                                                              (True, "standard", (2000, 2, 20, 0.0), 86400.0, 24, 5),
                                                              (True, "noleap", (2001, 12, 25, 0.0), 43200.0, 30, 30)])
def test_hanasaki_reservoirs_two_methods_in_device_order(memory, calendar, start, dt, steps, K):
    """Hanasaki-2006 reservoirs (h06_release in mr_lake.cuh: ring-buffer memory, only the month's mean recomputed) under two
    routing methods, stepped in the order route_device uses when lake state is shared -- batches of K steps, headwaters,
    then wavefront by wavefront and method by method -- against the oracle's step-by-step, method-by-method order: bit for bit."""
    from mizuroute_b200 import synth
    net, params, opts, ro = case("conus", n=500, seed=4, dt=dt, route_opt="35", steps=steps, lakes=10)
    assert synth.make_h06_lakes(net, np.random.default_rng(6), frac=0.7, memory=memory) >= 2
    opts.sim_start, opts.calendar = start, calendar
    ro = ro * 20.0
    o = Oracle(net, params, opts)
    qr = np.empty((steps + 1, net.nRch)); qo = np.empty((2, steps, net.nRch))
    qr[0] = o.get(orc.F_BASIN_QR1)
    for t in range(steps):
        o.step(ro[t])
        qr[t + 1] = o.get(orc.F_BASIN_QR1)
        qo[0, t] = o.get(orc.F_REACH_Q, orc.M_KW); qo[1, t] = o.get(orc.F_REACH_Q, orc.M_DW)
    h06 = np.ascontiguousarray(np.stack([net.lake_params[k] for k in H06_ORDER] + [np.zeros(net.nRch)]))
    L = emul.load_lake()
    qe = np.empty((2, steps, net.nRch)); ve = np.empty((2, net.nRch))
    msg = C.create_string_buffer(256)
    p = lambda a, ct: np.ascontiguousarray(a).ctypes.data_as(C.POINTER(ct))
    arrs = [np.ascontiguousarray(a) for a in (net.islake, net.lakeModelType, net.D03_MaxStorage, net.D03_Coefficient, net.D03_Power, net.D03_S0)]
    ierr = L.lake_emul_run_h06(C.c_int(net.nRch), C.c_int(net.nHRU), p(net.segId, C.c_int), p(net.downSegId, C.c_int), p(net.hruSegId, C.c_int),
                               p(net.area, C.c_double), p(net.length, C.c_double), p(net.slope, C.c_double), p(arrs[0], C.c_int), p(arrs[1], C.c_int),
                               p(arrs[2], C.c_double), p(arrs[3], C.c_double), p(arrs[4], C.c_double), p(arrs[5], C.c_double),
                               C.c_double(params.mann_n), C.c_double(params.wscale), C.c_double(opts.dt), C.c_int(opts.LakeInputOption),
                               C.c_int(steps), C.c_int(K), p(qr, C.c_double), p(h06, C.c_double),
                               C.c_int(start[0]), C.c_int(start[1]), C.c_int(start[2]), C.c_double(start[3]), C.c_int(int(calendar == "noleap")),
                               p(qe, C.c_double), p(ve, C.c_double), msg)
    assert ierr == 0, msg.value.decode()
    assert np.array_equal(qe, qo)
    assert np.array_equal(ve[0], o.get(orc.F_REACH_VOL1, orc.M_KW)) and np.array_equal(ve[1], o.get(orc.F_REACH_VOL1, orc.M_DW))
    hl = (net.islake == 1) & (net.lakeModelType == 2)
    assert (qo[0][:, hl] > 0.0).any()
