"""The stand-alone host `route_runoff` (control file + namelist + NetCDF-3 network/runoff, as the reference's
route_runoff.exe): input parsing without a GPU (--dry-run), and a full run on the GPU against the oracle."""
import json
import os
import subprocess

import numpy as np
import pytest

from mizuroute_b200 import build as mrbuild
from mizuroute_b200 import casefiles
from mizuroute_b200.network import RouteOptions, RouteParams
from tests.util import case


def _host():
    return mrbuild.build_host()


# The host tests that route run twice: against the real library on a GPU ("cuda", the parity tests proper), and -- so that
# the host's own I/O and bookkeeping are checked in the CPU suite too -- with a copy of the host linked against a test-only
# stand-in for the library built on the oracle (tests/stub/mr_stub.c; says nothing about the CUDA path).
BACKENDS = [pytest.param("oracle-stub", id="stub"), pytest.param("cuda", marks=pytest.mark.gpu, id="cuda")]


def _routing_host(backend):
    from tests import stub
    return stub.build() if backend == "oracle-stub" else _host()


def test_dry_run_parses_reference_style_case(tmp_path):
    net, params, opts, ro = case("random", n=90, seed=3, dt=3600.0, route_opt="012", steps=30)
    params = RouteParams(fshape=2.2, tscale=70000.0, velo=1.2, diff=4000.0, mann_n=0.02, wscale=0.0015)
    ctl = casefiles.write_case(str(tmp_path), net, params, opts, ro, case_name="dry", split_forcing=3, shuffle_hru_seed=1)
    r = subprocess.run([_host(), ctl, "--dry-run"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    info = json.loads(r.stdout.strip().splitlines()[0])
    assert (info["nRch"], info["nHRU"], info["nSteps"], info["route_opt"], info["dt"]) == (net.nRch, net.nHRU, 30, "012", 3600.0)
    assert (info["fshape"], info["tscale"], info["velo"], info["diff"], info["mann_n"], info["wscale"]) == (2.2, 70000.0, 1.2, 4000.0, 0.02, 0.0015)
    assert info["length_conv"] == 1e-3 and info["time_conv"] == 1.0 and info["first_record"] == 0


def test_control_file_errors(tmp_path):
    net, params, opts, ro = case("random", n=40, seed=3, dt=86400.0, route_opt="1", steps=4)
    ctl = casefiles.write_case(str(tmp_path), net, params, opts, ro)
    txt = open(ctl).read()
    bad = tmp_path / "bad.control"
    bad.write_text(txt + "<no_such_key>   1   ! unknown\n")
    r = subprocess.run([_host(), str(bad), "--dry-run"], capture_output=True, text=True)
    assert r.returncode == 81 and "unknown control key" in r.stderr          # read_control.f90:374-377
    import re
    bad.write_text(txt + "<is_flux_wm>   T   ! water management needs its file\n")
    r = subprocess.run([_host(), str(bad), "--dry-run"], capture_output=True, text=True)
    assert r.returncode != 0 and "<fname_wm> must be given" in r.stderr
    bad.write_text(re.sub(r"(<ro_time_stamp>\s+)start", r"\g<1>front", txt))
    r = subprocess.run([_host(), str(bad), "--dry-run"], capture_output=True, text=True)
    assert r.returncode != 0 and "must be start, end, or middle" in r.stderr        # read_control.f90:514-518


def _time_map(tmp_path, dt, forcing_dt, records, sim_steps, stamp=None, name="tm"):
    net, params, opts, ro = case("random", n=20, seed=3, dt=dt, route_opt="1", steps=records)
    ctl = casefiles.write_case(str(tmp_path), net, params, opts, ro, case_name=name, forcing_dt=forcing_dt, sim_steps=sim_steps, ro_time_stamp=stamp)
    r = subprocess.run([_host(), ctl, "--dry-run"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    a, b = (json.loads(x) for x in r.stdout.strip().splitlines()[:2])
    return a, b


def test_time_map_between_simulation_steps_and_forcing_records(tmp_path):
    """timeMap_sim_forc (get_basin_runoff.f90:256-369): a simulation step inside one record uses that record; a step over
    several records takes them with their time fractions."""
    a, b = _time_map(tmp_path, 3600.0, 86400.0, records=3, sim_steps=72, name="fine")        # hourly steps, daily runoff
    assert a["nSteps"] == 72 and b["dt_ro"] == 86400.0 and b["time_map"] == [[[0, 1.0]]] * 6
    a, b = _time_map(tmp_path, 86400.0, 3600.0, records=72, sim_steps=3, name="coarse")      # daily steps, hourly runoff
    assert a["nSteps"] == 3
    assert [len(m) for m in b["time_map"]] == [24, 24, 24] and [m[0][0] for m in b["time_map"]] == [0, 24, 48]
    assert all(abs(f - 1.0 / 24.0) < 1e-9 for m in b["time_map"] for _, f in m)
    a, b = _time_map(tmp_path, 7200.0, 10800.0, records=4, sim_steps=6, name="ragged")       # 2-hourly steps, 3-hourly runoff
    assert a["nSteps"] == 6
    want = [[[0, 1.0]], [[0, 0.5], [1, 0.5]], [[1, 1.0]], [[2, 1.0]], [[2, 0.5], [3, 0.5]], [[3, 1.0]]]
    assert b["time_map"] == want
    # records stamped at the middle / end of the interval they cover (casefiles shifts the stamps, not the data): same map
    for stamp in ("middle", "end"):
        a, b = _time_map(tmp_path, 3600.0, 3600.0, records=5, sim_steps=4, stamp=stamp, name=stamp)
        assert a["first_record"] == 0 and a["nSteps"] == 4 and b["time_map"] == [[[k, 1.0]] for k in range(4)]
    # a simulation period longer than the forcing is cut to the steps the forcing covers completely
    a, b = _time_map(tmp_path, 3600.0, 3600.0, records=5, sim_steps=9, name="clip")
    assert a["nSteps"] == 5


def test_history_file_plan(tmp_path):
    """<newFileFrequency>: a new history file when the day / month / year of a step's start changes (newFileAlarm,
    write_simoutput_pio.f90:110-135), stamped like get_hfilename (:329-392); records per file follow <outputFrequency>."""
    net, params, opts, ro = case("random", n=20, seed=3, dt=21600.0, route_opt="1", steps=14)
    def plan(freq, out_freq="1", start="2000-12-30 12:00:00"):
        ctl = casefiles.write_case(str(tmp_path), net, params, opts, ro, case_name="hp", start=start, new_file_frequency=freq, output_frequency=out_freq)
        r = subprocess.run([_host(), ctl, "--dry-run"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        return json.loads(r.stdout.strip().splitlines()[1])["history_plan"]
    assert plan("single") == [["hp.h.2000-12-30-43200.nc", 14]]
    assert plan("daily") == [["hp.h.2000-12-30-43200.nc", 2], ["hp.h.2000-12-31-00000.nc", 4], ["hp.h.2001-01-01-00000.nc", 4], ["hp.h.2001-01-02-00000.nc", 4]]
    assert plan("monthly") == [["hp.h.2000-12.nc", 6], ["hp.h.2001-01.nc", 8]]
    assert plan("yearly") == [["hp.h.2000.nc", 6], ["hp.h.2001.nc", 8]]
    assert plan("monthly", "2") == [["hp.h.2000-12.nc", 3], ["hp.h.2001-01.nc", 4]]
    assert plan("daily", start="2000-02-28 00:00:00")[1][0] == "hp.h.2000-02-29-00000.nc"          # leap day in the standard calendar
    ctl = casefiles.write_case(str(tmp_path), net, params, opts, ro, case_name="bad", new_file_frequency="weekly")
    r = subprocess.run([_host(), ctl, "--dry-run"], capture_output=True, text=True)
    assert r.returncode != 0 and "new output files" in r.stderr


def test_restart_write_plan(tmp_path):
    """<restart_write> specified / daily / monthly / yearly / last: the state is written after the step whose END is the
    restart time, and the file carries that time (restart_alarm + restart_fname, write_restart_pio.f90:110-253)."""
    net, params, opts, ro = case("random", n=20, seed=3, dt=3600.0, route_opt="1", steps=80)
    def plan(rw, **extra):
        ctl = casefiles.write_case(str(tmp_path), net, params, opts, ro, case_name="rp", start="2000-01-30 22:00:00", restart_write=rw, extra_keys=extra)
        r = subprocess.run([_host(), ctl, "--dry-run"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        return json.loads(r.stdout.strip().splitlines()[1])["restart_plan"]
    assert plan("never") == []
    assert plan("last") == [[79, "rp.r.2000-02-03-21600.nc"]]
    assert plan("specified", restart_date="2000-01-31 06:00:00") == [[7, "rp.r.2000-01-31-21600.nc"]]
    assert plan("daily") == [[1, "rp.r.2000-01-31-00000.nc"], [25, "rp.r.2000-02-01-00000.nc"], [49, "rp.r.2000-02-02-00000.nc"], [73, "rp.r.2000-02-03-00000.nc"]]
    assert plan("daily", restart_hour=12)[0] == [13, "rp.r.2000-01-31-43200.nc"]
    assert plan("monthly") == [[25, "rp.r.2000-02-01-00000.nc"]]
    assert plan("monthly", restart_day=31) == [[1, "rp.r.2000-01-31-00000.nc"]]
    assert plan("yearly", restart_month=2, restart_day=2) == [[49, "rp.r.2000-02-02-00000.nc"]]
    assert plan("yearly") == []
    ctl = casefiles.write_case(str(tmp_path), net, params, opts, ro, case_name="bad", restart_write="specified")
    r = subprocess.run([_host(), ctl, "--dry-run"], capture_output=True, text=True)
    assert r.returncode != 0 and "<restart_date> must be provided" in r.stderr


def _dump_forcing(ctl, tmp_path, cols):
    path = os.path.join(str(tmp_path), "forcing.f64")
    r = subprocess.run([_host(), ctl, "--dry-run", "--dump-forcing", path], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return np.fromfile(path, dtype=np.float64).reshape(-1, cols)


@pytest.mark.parametrize("dt,forcing_dt,records,sim_steps", [(3600.0, 3600.0, 12, 12), (3600.0, 10800.0, 10, 30), (10800.0, 3600.0, 36, 12),
                                                             (7200.0, 10800.0, 8, 12)])
def test_forcing_rows_fed_to_the_library(tmp_path, dt, forcing_dt, records, sim_steps):
    """get_hru_runoff on the host side: the rows the time loop hands to mr_step_batch (--dump-forcing) are the forcing
    records mapped onto the simulation steps (record the step lies in, or the time-weighted mean of the records it spans),
    in network HRU order, negatives zeroed (sort_flux, process_remap.f90:271-311)."""
    net, params, opts, ro = case("random", n=60, seed=4, dt=dt, route_opt="1", steps=records)
    ro = ro.copy(); ro[1, 3] = -2.0                                        # a negative value: removed by sort_flux
    same = forcing_dt == dt
    ctl = casefiles.write_case(str(tmp_path), net, params, opts, ro, case_name="rows", forcing_dt=None if same else forcing_dt,
                               sim_steps=None if same else sim_steps, split_forcing=3 if same else 1, shuffle_hru_seed=7 if same else None)
    got = _dump_forcing(ctl, tmp_path, net.nHRU)
    fine = 1800.0
    k = int(dt / fine)
    ro_fine = np.repeat(ro, int(forcing_dt / fine), axis=0)[:sim_steps * k]
    want = np.maximum(ro_fine.reshape(sim_steps, k, -1).mean(axis=1), 0.0) if not same else np.maximum(ro, 0.0)
    assert got.shape == want.shape
    if same or dt < forcing_dt and forcing_dt % dt == 0:
        assert np.array_equal(got, want)                                     # one record per step: values pass through untouched
    else:
        mask = np.ones_like(want, dtype=bool)
        if forcing_dt < dt: mask[0, 3] = False                              # the -2.0 enters a mean there (weighted like any value)
        np.testing.assert_allclose(got[mask], want[mask], rtol=1e-14)


def test_forcing_fill_values_are_skipped_in_the_time_mean(tmp_path):
    """read_1D_forcing (read_runoff.f90:312-322): records holding the fill value drop out of a step's mean and the
    remaining weights are renormalised; all records missing -> missing -> 0 after sort_flux."""
    net, params, opts, ro = case("random", n=30, seed=5, dt=10800.0, route_opt="1", steps=12)
    ro = ro.copy(); ro[0, 2] = -9999.0; ro[3:6, 4] = -9999.0
    ctl = casefiles.write_case(str(tmp_path), net, params, opts, ro, case_name="fill", forcing_dt=3600.0, sim_steps=4)
    got = _dump_forcing(ctl, tmp_path, net.nHRU)
    want = ro.reshape(4, 3, -1).mean(axis=1)
    want[0, 2] = ro[1:3, 2].mean(); want[1, 4] = 0.0
    np.testing.assert_allclose(got, want, rtol=1e-14)


def test_runoff_scale_and_offset(tmp_path):
    """<scale_factor_runoff> / <offset_value_runoff> (scale_forcing, get_basin_runoff.f90:375-423): applied to every value that
    is not missing, before negatives are removed; scale 0 without an offset switches the runoff off (:71-73)."""
    net, params, opts, ro = case("random", n=30, seed=5, dt=86400.0, route_opt="1", steps=4)
    ro = ro.copy(); ro[2, 5] = -9999.0
    rows = lambda **extra: _dump_forcing(casefiles.write_case(str(tmp_path), net, params, opts, ro, case_name="sc", extra_keys=extra), tmp_path, net.nHRU)
    want = 0.5 * ro + 1e-6; want[2, 5] = 0.0
    assert np.array_equal(rows(scale_factor_runoff=0.5, offset_value_runoff=1e-6), want)
    want = ro - 2e-5; want[2, 5] = 0.0
    assert np.array_equal(rows(offset_value_runoff=-2e-5), np.maximum(want, 0.0))
    assert not rows(scale_factor_runoff=0.0).any()


def test_netcdf3_writer_roundtrip_through_scipy(tmp_path):
    """nc3.h writes what scipy reads: exercised by the host's history file in the GPU test; here the network and
    runoff files written by scipy are the reader's input (dry run) with float32 runoff and 64-bit offsets."""
    net, params, opts, ro = case("binary", n=63, seed=2, dt=86400.0, route_opt="2", steps=6)
    ctl = casefiles.write_case(str(tmp_path), net, params, opts, ro.astype(np.float32).astype(np.float64))
    r = subprocess.run([_host(), ctl, "--dry-run"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_reader_accepts_classic_cdf1_and_rejects_netcdf4(tmp_path):
    """NetCDF-3 classic (CDF-1, 32-bit offsets) network file is read like the 64-bit-offset one; an HDF5 (netCDF-4)
    file is refused with a clear message."""
    from scipy.io import netcdf_file
    net, params, opts, ro = case("random", n=30, seed=8, dt=86400.0, route_opt="1", steps=3)
    ctl = casefiles.write_case(str(tmp_path), net, params, opts, ro, case_name="cdf1")
    path = os.path.join(str(tmp_path), "ancillary", "ntopo.nc")
    f = netcdf_file(path, "w", version=1)                   # rewrite the network as CDF-1
    f.createDimension("seg", net.nRch); f.createDimension("hru", net.nHRU)
    for nm, dat, dim, typ in (("segId", net.segId, "seg", "i"), ("downSegId", net.downSegId, "seg", "i"), ("length", net.length, "seg", "d"),
                              ("slope", net.slope.astype(np.float32), "seg", "f"), ("HRUid", net.hruId, "hru", "i"),
                              ("hruSegId", net.hruSegId, "hru", "i"), ("area", net.area, "hru", "d")):
        v = f.createVariable(nm, typ, (dim,)); v[:] = dat
    f.close()
    assert open(path, "rb").read(4) == b"CDF\x01"
    r = subprocess.run([_host(), ctl, "--dry-run"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert json.loads(r.stdout.strip().splitlines()[0])["nRch"] == net.nRch
    with open(path, "wb") as fh:
        fh.write(b"\x89HDF\r\n\x1a\n" + b"\0" * 64)
    r = subprocess.run([_host(), ctl, "--dry-run"], capture_output=True, text=True)
    assert r.returncode != 0 and "netCDF-4/HDF5 is not supported" in r.stderr


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("route,dt,lakes", [("012", 3600.0, 0), ("12", 86400.0, 5), ("345", 3600.0, 4)])
def test_host_run_matches_oracle(tmp_path, backend, route, dt, lakes):
    from oracle.oracle import Oracle
    net, params, opts, ro = case("conus", n=600, seed=8, dt=dt, route_opt=route, steps=40, lakes=lakes)
    ctl = casefiles.write_case(str(tmp_path), net, params, opts, ro, case_name="gpu", split_forcing=2, shuffle_hru_seed=5)
    r = subprocess.run([_routing_host(backend), ctl, "--batch", "16"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    hist = json.loads(r.stdout.strip().splitlines()[-1])["history"]
    out = casefiles.read_history(hist)
    qo = Oracle(net, params, opts).run(ro)
    assert np.array_equal(out["reachID"], net.segId)
    assert np.array_equal(out["time"], np.arange(40) * dt)
    names = {"0": "sumUpstreamRunoff", "1": "IRFroutedRunoff", "2": "KWTroutedRunoff", "3": "KWroutedRunoff", "4": "MCroutedRunoff", "5": "DWroutedRunoff"}
    for i, c in enumerate(route):
        got = out[names[c]]
        assert got.dtype == np.float32 and got.shape == (40, net.nRch)       # history is float32 [time, seg]
        np.testing.assert_allclose(got, qo[i].astype(np.float32), rtol=2e-6 if c in "01" else 1e-4, atol=1e-30)


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("dt,forcing_dt,records,sim_steps", [(3600.0, 10800.0, 10, 30), (10800.0, 3600.0, 36, 12), (7200.0, 10800.0, 8, 12)])
def test_host_maps_forcing_records_onto_simulation_steps(tmp_path, backend, dt, forcing_dt, records, sim_steps):
    """dt_qsim != dt_ro: the host feeds each step the record it lies in, or the time-weighted mean of the records it
    spans (timeMap_sim_forc + read_1D_forcing); the oracle is run on the runoff averaged the same way here."""
    from oracle.oracle import Oracle
    net, params, opts, ro = case("conus", n=300, seed=4, dt=dt, route_opt="12", steps=records)
    ctl = casefiles.write_case(str(tmp_path), net, params, opts, ro, case_name="tmap", forcing_dt=forcing_dt, sim_steps=sim_steps)
    r = subprocess.run([_routing_host(backend), ctl, "--batch", "5"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = casefiles.read_history(json.loads(r.stdout.strip().splitlines()[-1])["history"])
    fine = 1800.0                                                          # common divisor of every interval used here
    ro_fine = np.repeat(ro, int(forcing_dt / fine), axis=0)
    k = int(dt / fine)
    ro_sim = ro_fine[:sim_steps * k].reshape(sim_steps, k, -1).mean(axis=1)
    qo = Oracle(net, params, opts).run(ro_sim)
    assert out["IRFroutedRunoff"].shape == (sim_steps, net.nRch)
    np.testing.assert_allclose(out["IRFroutedRunoff"], qo[0].astype(np.float32), rtol=3e-6, atol=1e-30)
    np.testing.assert_allclose(out["KWTroutedRunoff"], qo[1].astype(np.float32), rtol=1e-4, atol=1e-30)


@pytest.mark.parametrize("backend", BACKENDS)
def test_daily_history_files_hold_the_single_file_run(tmp_path, backend):
    net, params, opts, ro = case("random", n=150, seed=6, dt=10800.0, route_opt="12", steps=20)
    d = str(tmp_path)
    run = lambda ctl: subprocess.run([_routing_host(backend), ctl, "--batch", "6"], capture_output=True, text=True)
    r = run(casefiles.write_case(d, net, params, opts, ro, case_name="one", start="2000-02-28 06:00:00")); assert r.returncode == 0, r.stderr
    one = casefiles.read_history(json.loads(r.stdout.strip().splitlines()[-1])["history"])
    r = run(casefiles.write_case(d, net, params, opts, ro, case_name="many", start="2000-02-28 06:00:00", new_file_frequency="daily")); assert r.returncode == 0, r.stderr
    files = json.loads(r.stdout.strip().splitlines()[-1])["history_files"]
    assert [os.path.basename(f) for f in files] == ["many.h.2000-02-28-21600.nc", "many.h.2000-02-29-00000.nc", "many.h.2000-03-01-00000.nc"]
    parts = [casefiles.read_history(f) for f in files]
    assert [len(p["time"]) for p in parts] == [6, 8, 6]
    for v in ("time", "IRFroutedRunoff", "KWTroutedRunoff", "dlayRunoff"):
        assert np.array_equal(np.concatenate([p[v] for p in parts]), one[v]), v
    assert all(np.array_equal(p["reachID"], net.segId) for p in parts)


@pytest.mark.parametrize("backend", BACKENDS)
def test_exact_restart(tmp_path, backend):
    """ERS, the reference's main regression idea (cime_config/testdefs/testlist_mizuRoute.xml): 40 steps in one run ==
    20 steps + restart file (reference schema: qfuture, irf_qfuture, numWaves, tentry/texit/qwave/routed, ...) + 20 steps."""
    net, params, opts, ro = case("conus", n=500, seed=9, dt=3600.0, route_opt="012", steps=40, lakes=4)
    d = str(tmp_path)
    full = casefiles.write_case(d, net, params, opts, ro, case_name="full")
    first = casefiles.write_case(d, net, params, opts, ro[:20], case_name="first", restart_write="last")
    run = lambda ctl: subprocess.run([_routing_host(backend), ctl, "--batch", "7"], capture_output=True, text=True)
    r = run(full); assert r.returncode == 0, r.stderr
    h_full = casefiles.read_history(json.loads(r.stdout.strip().splitlines()[-1])["history"])
    r = run(first); assert r.returncode == 0, r.stderr
    lines = [json.loads(x) for x in r.stdout.strip().splitlines()]
    h_first = casefiles.read_history(next(x["history"] for x in lines if "history" in x))
    rfile = next(x["restart"] for x in lines if "restart" in x)
    rst = casefiles.read_history(rfile)
    assert {"reachID", "basin_q", "qfuture", "numQF", "irf_qfuture", "volume_irf", "numWaves", "tentry", "texit", "qwave", "routed"} <= set(rst)
    assert rst["tentry"].shape == (22, net.nRch) and rst["qfuture"].shape[1] == net.nRch          # Fortran (seg, wave) = file (wave, seg)
    second = casefiles.write_case(d, net, params, opts, ro[20:], case_name="second", fname_state_in=os.path.basename(rfile), first_step=20)
    r = run(second); assert r.returncode == 0, r.stderr
    h_second = casefiles.read_history(json.loads(r.stdout.strip().splitlines()[-1])["history"])
    for v in ("sumUpstreamRunoff", "IRFroutedRunoff", "KWTroutedRunoff"):
        assert np.array_equal(np.concatenate([h_first[v], h_second[v]]), h_full[v]), v


@pytest.mark.parametrize("backend", BACKENDS)
def test_restart_files_written_during_the_run_continue_it_exactly(tmp_path, backend):
    """<restart_write> daily on a 6-hourly run with --batch 5: batches are cut where a restart file is due, and a run
    continued from the second file reproduces the rest of the uninterrupted run bit for bit."""
    net, params, opts, ro = case("conus", n=300, seed=3, dt=21600.0, route_opt="12", steps=14, lakes=3)
    d = str(tmp_path)
    run = lambda ctl: subprocess.run([_routing_host(backend), ctl, "--batch", "5"], capture_output=True, text=True)
    r = run(casefiles.write_case(d, net, params, opts, ro, case_name="full", start="2000-03-01 12:00:00", restart_write="daily")); assert r.returncode == 0, r.stderr
    lines = [json.loads(x) for x in r.stdout.strip().splitlines()]
    rfiles = [x["restart"] for x in lines if "restart" in x]
    assert [os.path.basename(f) for f in rfiles] == ["full.r.2000-03-02-00000.nc", "full.r.2000-03-03-00000.nc", "full.r.2000-03-04-00000.nc", "full.r.2000-03-05-00000.nc"]
    h_full = casefiles.read_history(next(x["history"] for x in lines if "history" in x))
    k0 = 6                                                                  # 2000-03-03 00:00 = 6 steps after the start
    assert list(casefiles.read_history(rfiles[1])["time_bound"]) == [(k0 - 1) * 21600.0, k0 * 21600.0]      # TSEC(1:2) of the last step routed (write_restart_pio.f90:812)
    ctl = casefiles.write_case(d, net, params, opts, ro[k0:], case_name="cont", start="2000-03-01 12:00:00", first_step=k0, fname_state_in=os.path.basename(rfiles[1]))
    r = run(ctl); assert r.returncode == 0, r.stderr
    h_cont = casefiles.read_history(json.loads(r.stdout.strip().splitlines()[-1])["history"])
    for v in ("IRFroutedRunoff", "KWTroutedRunoff", "dlayRunoff"):
        assert np.array_equal(h_cont[v], h_full[v][k0:]), v


def test_restart_inside_an_output_period_carries_the_partial_means(tmp_path):
    """A restart file written inside an output period holds the reference's history state -- `nt`, `history_time` and the running
    sums under their history-file names (write_restart_pio.f90:315-324,1324-1480) -- and the continuation run finishes the period:
    its history records equal the uninterrupted run's bit for bit.  Host logic only (stand-in library): the device path is the
    one of the exact-restart tests."""
    net, params, opts, ro = case("conus", n=200, seed=8, dt=21600.0, route_opt="12", steps=22)
    d = str(tmp_path)
    run = lambda ctl: subprocess.run([_routing_host("oracle-stub"), ctl, "--batch", "5"], capture_output=True, text=True)
    keys = {"basRunoff": "T"}
    # daily means of 6-hourly steps from 12:00 on: the periods run noon to noon, the daily restart files fall at midnight
    r = run(casefiles.write_case(d, net, params, opts, ro, case_name="full", start="2000-03-01 12:00:00", restart_write="daily", output_frequency="daily", extra_keys=keys))
    assert r.returncode == 0, r.stderr
    lines = [json.loads(x) for x in r.stdout.strip().splitlines()]
    rfiles = [x["restart"] for x in lines if "restart" in x]
    h_full = casefiles.read_history(next(x["history"] for x in lines if "history" in x))
    k0 = 6                                                                  # 2000-03-03 00:00: two steps into the second period
    st = casefiles.read_history(rfiles[1])
    assert int(st["nt"][0]) == 2 and "KWTroutedRunoff" in st and "dlayRunoff" in st and "basRunoff" in st
    assert np.allclose(st["history_time"][1] - st["history_time"][0], 2 * 21600.0)
    ctl = casefiles.write_case(d, net, params, opts, ro[k0:], case_name="cont", start="2000-03-01 12:00:00", first_step=k0, fname_state_in=os.path.basename(rfiles[1]),
                               output_frequency="daily", extra_keys=keys)
    r = run(ctl); assert r.returncode == 0, r.stderr
    h_cont = casefiles.read_history(json.loads(r.stdout.strip().splitlines()[-1])["history"])
    for v in ("IRFroutedRunoff", "KWTroutedRunoff", "dlayRunoff", "basRunoff"):
        assert np.array_equal(h_cont[v], h_full[v][1:]), v                  # from the period the restart fell into
    assert np.array_equal(h_cont["time"] + k0 * 21600.0, h_full["time"][1:])   # (time is "since <sim_start>", six steps later here)


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("option", [0, 2])
def test_host_feeds_lake_evaporation_and_precipitation(tmp_path, backend, option):
    """<is_lake_sim> T with <LakeInputOption> 0 / 2: evaporation and precipitation are read from the forcing file beside the
    runoff (get_basin_runoff.f90:136-197), with <scale_factor_Ep> applied, and reach lake_route through
    mr_upload_lake_forcing; two methods, so the evaporation cut of a dried lake is shared (methods run in route_opt order)."""
    from oracle.oracle import Oracle
    net, params, opts, ro = case("conus", n=500, seed=4, dt=86400.0, route_opt="14", steps=12, lakes=7)
    opts.LakeInputOption = option
    rng = np.random.default_rng(3)
    ev = np.abs(rng.lognormal(np.log(3e-5), 0.5, size=ro.shape)); pr = np.abs(rng.lognormal(np.log(2e-5), 0.8, size=ro.shape))
    lakes = np.flatnonzero(net.islake == 1)
    ev[:, np.isin(net.hruSegId, net.segId[lakes[:2]])] *= 3.0e4                 # two lakes run dry
    ctl = casefiles.write_case(str(tmp_path), net, params, opts, ro, case_name="lakeep", lake_forcing=(ev, pr), extra_keys={"scale_factor_Ep": 0.5})
    r = subprocess.run([_routing_host(backend), ctl, "--batch", "5"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = casefiles.read_history(json.loads(r.stdout.strip().splitlines()[-1])["history"])
    qo = Oracle(net, params, opts).run(ro, evapo=0.5 * ev, precip=pr)
    q_plain = Oracle(net, params, opts).run(ro)
    assert not np.array_equal(qo, q_plain)
    np.testing.assert_allclose(out["IRFroutedRunoff"], qo[0].astype(np.float32), rtol=2e-6, atol=1e-30)
    np.testing.assert_allclose(out["MCroutedRunoff"], qo[1].astype(np.float32), rtol=1e-4, atol=1e-30)


@pytest.mark.parametrize("backend", BACKENDS)
def test_hype_reservoirs_and_their_calendar_survive_a_restart(tmp_path, backend):
    """lakeModelType 3: HYP_* are read from the river-network file and the day of year follows <sim_start>; a run continued
    from a restart file -- whose library handle counts steps from the cold start -- keeps the calendar (2000-02-25 + 14 days
    crosses the leap day) and reproduces the uninterrupted run bit for bit."""
    from mizuroute_b200 import synth
    from oracle.oracle import Oracle
    net, params, opts, ro = case("conus", n=400, seed=4, dt=86400.0, route_opt="13", steps=14, lakes=8)
    assert synth.make_hype_lakes(net, np.random.default_rng(5), frac=0.7) >= 2
    ro = ro * 30.0
    d = str(tmp_path)
    start = "2000-02-25 00:00:00"
    run = lambda ctl: subprocess.run([_routing_host(backend), ctl, "--batch", "4"], capture_output=True, text=True)
    r = run(casefiles.write_case(d, net, params, opts, ro, case_name="full", start=start)); assert r.returncode == 0, r.stderr
    h_full = casefiles.read_history(json.loads(r.stdout.strip().splitlines()[-1])["history"])
    opts.sim_start = (2000, 2, 25, 0.0)
    qo = Oracle(net, params, opts).run(ro)
    np.testing.assert_allclose(h_full["IRFroutedRunoff"], qo[0].astype(np.float32), rtol=2e-6, atol=1e-30)
    np.testing.assert_allclose(h_full["KWroutedRunoff"], qo[1].astype(np.float32), rtol=1e-4, atol=1e-30)
    r = run(casefiles.write_case(d, net, params, opts, ro[:6], case_name="first", start=start, restart_write="last")); assert r.returncode == 0, r.stderr
    rfile = next(json.loads(x)["restart"] for x in r.stdout.strip().splitlines() if "restart" in x)
    r = run(casefiles.write_case(d, net, params, opts, ro[6:], case_name="second", start=start, first_step=6, fname_state_in=os.path.basename(rfile)))
    assert r.returncode == 0, r.stderr
    h_second = casefiles.read_history(json.loads(r.stdout.strip().splitlines()[-1])["history"])
    for v in ("IRFroutedRunoff", "KWroutedRunoff"):
        assert np.array_equal(h_second[v], h_full[v][6:]), v


@pytest.mark.parametrize("backend", BACKENDS)
def test_hanasaki_reservoirs_from_the_network_file(tmp_path, backend):
    """lakeModelType 2: the H06_* variables of the river-network file reach the routing, two methods share the reservoirs'
    state, and the inflow memory makes the release depend on the run's own history."""
    from mizuroute_b200 import synth
    from oracle.oracle import Oracle
    net, params, opts, ro = case("conus", n=400, seed=4, dt=86400.0, route_opt="13", steps=20, lakes=8)
    assert synth.make_h06_lakes(net, np.random.default_rng(6), frac=0.7, memory=True) >= 2
    ro = ro * 20.0
    ctl = casefiles.write_case(str(tmp_path), net, params, opts, ro, case_name="h06", start="2000-02-20 00:00:00")
    r = subprocess.run([_routing_host(backend), ctl, "--batch", "6"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = casefiles.read_history(json.loads(r.stdout.strip().splitlines()[-1])["history"])
    opts.sim_start = (2000, 2, 20, 0.0)
    qo = Oracle(net, params, opts).run(ro)
    np.testing.assert_allclose(out["IRFroutedRunoff"], qo[0].astype(np.float32), rtol=2e-6, atol=1e-30)
    np.testing.assert_allclose(out["KWroutedRunoff"], qo[1].astype(np.float32), rtol=1e-4, atol=1e-30)


@pytest.mark.parametrize("backend", BACKENDS)
def test_host_reads_the_water_management_file(tmp_path, backend):
    """<is_flux_wm> / <is_vol_wm> T: fluxes and target volumes [time, seg] of <fname_wm>, given for a shuffled subset of the
    reaches (the others get "none"), reach the routing through mr_upload_wm; LakeTargVol comes with the river network."""
    from oracle.oracle import Oracle
    net, params, opts, ro = case("conus", n=500, seed=4, dt=86400.0, route_opt="14", steps=12, lakes=7)
    K = ro.shape[0]
    rng = np.random.default_rng(11)
    lk = np.flatnonzero(net.islake == 1)
    net.lake_params = {"LakeTargVol": np.isin(np.arange(net.nRch), lk[:3]).astype(np.float64)}
    sub = rng.permutation(net.nRch)[: int(0.6 * net.nRch)]                     # reaches the file knows, in file order
    sub = np.union1d(sub, lk)[rng.permutation(np.union1d(sub, lk).size)]
    flux_f = rng.choice([-1.0, 1.0], (K, sub.size)) * rng.lognormal(np.log(0.05), 1.5, (K, sub.size))
    vol_f = np.where(net.islake[sub] == 1, rng.uniform(1e6, 5e7, (K, sub.size)), 0.0)
    ctl = casefiles.write_case(str(tmp_path), net, params, opts, ro, case_name="wm", wm=(net.segId[sub], flux_f, vol_f))
    r = subprocess.run([_routing_host(backend), ctl, "--batch", "5"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = casefiles.read_history(json.loads(r.stdout.strip().splitlines()[-1])["history"])
    o = Oracle(net, params, opts)
    qo = np.empty((2, K, net.nRch))
    for t in range(K):
        flux = np.full(net.nRch, -9999.0); flux[sub] = flux_f[t]
        vol = np.zeros(net.nRch); vol[sub] = vol_f[t]
        o.set_wm(flux, vol, vol_jumpstart=True)
        o.step(ro[t])
        qo[0, t] = o.get(0, 1); qo[1, t] = o.get(0, 4)          # F_REACH_Q of IRF and MC
    np.testing.assert_allclose(out["IRFroutedRunoff"], qo[0].astype(np.float32), rtol=2e-6, atol=1e-30)
    np.testing.assert_allclose(out["MCroutedRunoff"], qo[1].astype(np.float32), rtol=1e-4, atol=1e-30)
    assert not np.array_equal(qo, Oracle(net, params, opts).run(ro))


@pytest.mark.parametrize("backend", BACKENDS)
def test_exact_restart_of_the_euler_schemes(tmp_path, backend):
    """q_sub_kw / q_sub_mc / q_sub_dw [mol, seg] and volume_* in the restart file (popMetadat.f90:283-295): 24 steps in one
    run == 12 steps + restart + 12 steps, bit for bit, with <floodplain> T."""
    net, params, opts, ro = case("conus", n=400, seed=6, dt=3600.0, route_opt="345", steps=24, floodplain=True)
    d = str(tmp_path)
    run = lambda ctl: subprocess.run([_routing_host(backend), ctl, "--batch", "5"], capture_output=True, text=True)
    fp = {"floodplain": "T"}
    r = run(casefiles.write_case(d, net, params, opts, ro, case_name="full", extra_keys=fp)); assert r.returncode == 0, r.stderr
    h_full = casefiles.read_history(json.loads(r.stdout.strip().splitlines()[-1])["history"])
    r = run(casefiles.write_case(d, net, params, opts, ro[:12], case_name="first", restart_write="last", extra_keys=fp)); assert r.returncode == 0, r.stderr
    lines = [json.loads(x) for x in r.stdout.strip().splitlines()]
    rfile = next(x["restart"] for x in lines if "restart" in x)
    rst = casefiles.read_history(rfile)
    assert rst["q_sub_kw"].shape == (20, net.nRch) and rst["q_sub_mc"].shape == (2, net.nRch) and rst["q_sub_dw"].shape == (20, net.nRch)
    assert {"volume_kw", "volume_mc", "volume_dw"} <= set(rst)
    r = run(casefiles.write_case(d, net, params, opts, ro[12:], case_name="second", fname_state_in=os.path.basename(rfile), first_step=12, extra_keys=fp))
    assert r.returncode == 0, r.stderr
    h_second = casefiles.read_history(json.loads(r.stdout.strip().splitlines()[-1])["history"])
    for v in ("KWroutedRunoff", "MCroutedRunoff", "DWroutedRunoff"):
        assert np.array_equal(h_second[v], h_full[v][12:]), v
    if backend == "oracle-stub":                      # <floodplain> T reached the routing: differs from the default geometry
        from oracle.oracle import Oracle
        q_fp = Oracle(net, params, opts).run(ro)
        np.testing.assert_allclose(h_full["MCroutedRunoff"], q_fp[1].astype(np.float32), rtol=2e-6, atol=1e-30)


@pytest.mark.parametrize("backend", BACKENDS)
def test_daily_mean_output_and_delayed_runoff(tmp_path, backend):
    """<outputFrequency> daily on an hourly run: every history record is the mean of 24 steps (histVars_data.f90:154-246);
    dlayRunoff = BASIN_QR(1)."""
    from oracle import oracle as orc
    from oracle.oracle import Oracle
    net, params, opts, ro = case("random", n=120, seed=2, dt=3600.0, route_opt="1", steps=60)
    ctl = casefiles.write_case(str(tmp_path), net, params, opts, ro, case_name="agg", output_frequency="daily")
    r = subprocess.run([_routing_host(backend), ctl, "--batch", "25"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = casefiles.read_history(json.loads(r.stdout.strip().splitlines()[-1])["history"])
    o = Oracle(net, params, opts)
    q, qr = [], []
    for t in range(60):
        o.step(ro[t]); q.append(o.get(orc.F_REACH_Q, orc.M_IRF)); qr.append(o.get(orc.F_BASIN_QR1))
    q, qr = np.array(q), np.array(qr)
    want_q = np.stack([q[0:24].mean(0), q[24:48].mean(0), q[48:60].mean(0)])
    want_r = np.stack([qr[0:24].mean(0), qr[24:48].mean(0), qr[48:60].mean(0)])
    assert np.array_equal(out["time"], np.array([0.0, 86400.0, 172800.0]))
    assert np.array_equal(out["time_bounds"], np.array([[0.0, 86400.0], [86400.0, 172800.0], [172800.0, 60 * 3600.0]]))    # historyFile.f90:372
    np.testing.assert_allclose(out["IRFroutedRunoff"], want_q, rtol=3e-6, atol=1e-30)
    np.testing.assert_allclose(out["dlayRunoff"], want_r, rtol=3e-6, atol=1e-30)


@pytest.mark.parametrize("backend", BACKENDS)
def test_host_reads_the_gauge_files_for_direct_insertion(tmp_path, backend):
    """<qmodOption> 1: gauge ids of <fname_gageObs> are linked to reaches through the csv <gageMetaFile> (one site is not in
    the csv, one points at a reach outside the network), records exist only at some of the steps (the others count up
    Qelapsed), fill values and negative flows are skipped -- the observations reach the routing through mr_upload_obs."""
    from oracle import oracle as orc
    from oracle.oracle import Oracle
    net, params, opts, ro = case("conus", n=400, seed=4, dt=3600.0, route_opt="14", steps=20)
    K = ro.shape[0]
    rng = np.random.default_rng(5)
    base = Oracle(net, params, opts).run(ro)[0]
    rch = rng.choice(net.nRch, 12, replace=False)
    gage_ids = ["G%05d" % i for i in range(14)]
    csv_ids, csv_rch = gage_ids[:13], list(net.segId[rch]) + [int(net.segId.max()) + 77]     # G00012 -> unknown reach; G00013 not in the csv
    rec_steps = [2, 3, 4, 9, 15]
    flow = np.full((len(rec_steps), 14), np.nan)
    for i, k in enumerate(rec_steps):
        flow[i, :12] = base[k, rch] * rng.uniform(0.4, 2.0, 12)
        flow[i, 12:] = 5.0
    flow[1, 3] = np.nan; flow[2, 5] = -3.0
    blend, trend = 3, 2
    ctl = casefiles.write_case(str(tmp_path), net, params, opts, ro, case_name="da",
                               gauges=(gage_ids, csv_ids, csv_rch, [k * opts.dt for k in rec_steps], flow, blend, trend))
    r = subprocess.run([_routing_host(backend), ctl, "--batch", "6"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = casefiles.read_history(json.loads(r.stdout.strip().splitlines()[-1])["history"])
    o = Oracle(net, params, opts); o.set_da(1, blend, trend)
    qo = np.empty((2, K, net.nRch))
    for t in range(K):
        if t in rec_steps:
            obs = np.full(net.nRch, np.nan); obs[rch] = flow[rec_steps.index(t), :12]
            o.set_obs(obs)
        else:
            o.set_obs(None)
        o.step(ro[t])
        qo[0, t] = o.get(orc.F_REACH_Q, 1); qo[1, t] = o.get(orc.F_REACH_Q, 4)
    np.testing.assert_allclose(out["IRFroutedRunoff"], qo[0].astype(np.float32), rtol=2e-6, atol=1e-30)
    np.testing.assert_allclose(out["MCroutedRunoff"], qo[1].astype(np.float32), rtol=1e-4, atol=1e-30)
    assert not np.array_equal(qo[0], base)


@pytest.mark.parametrize("backend", BACKENDS)
def test_restart_under_data_assimilation_carries_the_discharge_error(tmp_path, backend):
    """qerror_irf / qerror_mc are written to the restart file only under <qmodOption> 1 (write_restart_pio.f90:1021-1030) and read
    back when present (read_restart.f90:358-366); Qobs / Qelapsed are not part of the reference's restart file and start from
    zero again (init_model_data.f90:511-512) -- so a restarted run equals the oracle told the same: Qerror kept, Qobs = 0,
    Qelapsed = 0."""
    from oracle import oracle as orc
    from oracle.oracle import Oracle
    net, params, opts, ro = case("conus", n=300, seed=6, dt=3600.0, route_opt="14", steps=16)
    rng = np.random.default_rng(8)
    base = Oracle(net, params, opts).run(ro)[0]
    rch = rng.choice(net.nRch, 10, replace=False)
    gage_ids = ["S%03d" % i for i in range(10)]
    rec_steps = [1, 5, 6, 11]
    flow = np.stack([base[k, rch] * rng.uniform(0.5, 1.8, 10) for k in rec_steps])
    blend, trend = 6, 1
    d = str(tmp_path)
    g = lambda steps: (gage_ids, gage_ids, list(net.segId[rch]), [k * opts.dt for k in steps], flow[[rec_steps.index(k) for k in steps]], blend, trend)
    run = lambda ctl: subprocess.run([_routing_host(backend), ctl, "--batch", "5"], capture_output=True, text=True)
    r = run(casefiles.write_case(d, net, params, opts, ro[:8], case_name="first", restart_write="last", gauges=g(rec_steps)))
    assert r.returncode == 0, r.stderr
    rfile = next(x["restart"] for x in (json.loads(l) for l in r.stdout.strip().splitlines()) if "restart" in x)
    rst = casefiles.read_history(rfile)
    assert {"qerror_irf", "qerror_mc"} <= set(rst) and np.abs(rst["qerror_irf"]).max() > 0.0
    # second leg: the gauge file's time axis counts from the start of the whole run
    ctl = casefiles.write_case(d, net, params, opts, ro[8:], case_name="second", fname_state_in=os.path.basename(rfile), first_step=8,
                               gauges=(gage_ids, gage_ids, list(net.segId[rch]), [(k - 8) * opts.dt for k in rec_steps if k >= 8],
                                       flow[[i for i, k in enumerate(rec_steps) if k >= 8]], blend, trend))
    r = run(ctl); assert r.returncode == 0, r.stderr
    h2 = casefiles.read_history(json.loads(r.stdout.strip().splitlines()[-1])["history"])
    o = Oracle(net, params, opts); o.set_da(1, blend, trend)
    qo = np.empty((2, 16, net.nRch))
    for t in range(16):
        if t == 8:                                              # what the restart file does not carry
            o.set(orc.F_QOBS, np.zeros(net.nRch)); lib_reset_elapsed(o)
        if t in rec_steps:
            obs = np.full(net.nRch, np.nan); obs[rch] = flow[rec_steps.index(t)]
            o.set_obs(obs)
        else:
            o.set_obs(None)
        o.step(ro[t])
        qo[0, t] = o.get(orc.F_REACH_Q, 1); qo[1, t] = o.get(orc.F_REACH_Q, 4)
    np.testing.assert_allclose(h2["IRFroutedRunoff"], qo[0, 8:].astype(np.float32), rtol=2e-6, atol=1e-30)
    np.testing.assert_allclose(h2["MCroutedRunoff"], qo[1, 8:].astype(np.float32), rtol=1e-4, atol=1e-30)


def lib_reset_elapsed(o):
    import ctypes as C
    from oracle import oracle as orc
    z = np.zeros(o.net.nRch, dtype=np.int32)
    orc.lib().mro_set_qelapsed(o.h, z.ctypes.data_as(C.POINTER(C.c_int)))


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("freq,keys", [("daily", {"IRFvolume": "T", "KWvolume": "T"}), ("5", {"IRFinflow": "T", "instRunoff": "T", "KWvolume": "T"}),
                                       ("1", {"IRFvolume": "T", "KWinflow": "T"})], ids=["daily-volumes", "5-inflow-inst", "1-mixed"])
def test_history_volume_inflow_and_instantaneous_runoff(tmp_path, backend, freq, keys):
    """<IRFvolume> ... = REACH_VOL(1) at the END of the output period, <IRFinflow> ... = mean REACH_INFLOW, <instRunoff> = mean
    BASIN_QI over it (histVars_data.f90:200-246) -- the host cuts its batches where the library's last-step values are needed."""
    from oracle import oracle as orc
    from oracle.oracle import Oracle
    net, params, opts, ro = case("random", n=90, seed=3, dt=3600.0, route_opt="13", steps=52)
    ctl = casefiles.write_case(str(tmp_path), net, params, opts, ro, case_name="vars", output_frequency=freq, extra_keys=keys)
    r = subprocess.run([_routing_host(backend), ctl, "--batch", "7"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = casefiles.read_history(json.loads(r.stdout.strip().splitlines()[-1])["history"])
    o = Oracle(net, params, opts)
    ser = {k: [] for k in ("IRFvolume", "KWvolume", "IRFinflow", "KWinflow", "instRunoff", "IRFroutedRunoff")}
    for t in range(52):
        o.step(ro[t])
        ser["IRFvolume"].append(o.get(orc.F_REACH_VOL1, 1)); ser["KWvolume"].append(o.get(orc.F_REACH_VOL1, 3))
        ser["IRFinflow"].append(o.get(orc.F_REACH_INFLOW, 1)); ser["KWinflow"].append(o.get(orc.F_REACH_INFLOW, 3))
        ser["instRunoff"].append(o.get(orc.F_BASIN_QI)); ser["IRFroutedRunoff"].append(o.get(orc.F_REACH_Q, 1))
    n = 24 if freq == "daily" else int(freq)
    bounds = [(a, min(a + n, 52)) for a in range(0, 52, n)]
    for name in list(keys) + ["IRFroutedRunoff"]:
        a = np.array(ser[name])
        want = np.stack([a[hi - 1] if name.endswith("volume") else a[lo:hi].mean(0) for lo, hi in bounds])
        np.testing.assert_allclose(out[name], want, rtol=1e-4 if name.startswith("KW") else 3e-6, atol=1e-12, err_msg=name)
    assert set(out) >= set(keys) and "KWTvolume" not in out
    want_b = np.stack([ro[lo:hi].mean(0) for lo, hi in bounds])                  # <basRunoff>: the HRU runoff as read, period mean
    np.testing.assert_allclose(out["basRunoff"], want_b, rtol=3e-6, atol=1e-30)
    assert np.array_equal(out["basinID"], net.hruId)


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("dt,forcing_dt,records,sim_steps,bas", [(3600.0, 3600.0, 12, 12, "F"), (10800.0, 3600.0, 36, 12, "T"), (7200.0, 10800.0, 8, 12, "F")])
def test_device_ingest_gives_the_same_history_as_host_built_rows(tmp_path, backend, dt, forcing_dt, records, sim_steps, bas):
    """--device-ingest: the runoff records travel as they are and mr_ingest_records builds the rows (time-weighted mean, scale,
    offset, sort_flux of shuffled forcing HRUs) -- the history file equals the one of the default path byte for byte."""
    net, params, opts, ro = case("conus", n=300, seed=4, dt=dt, route_opt="01", steps=records)
    ro = ro.copy(); ro[1, 3] = -2.0; ro[2, 5] = -9999.0
    same = forcing_dt == dt
    keys = {"scale_factor_runoff": "1.25", "offset_value_runoff": "1.e-10", "basRunoff": bas}
    outs = []
    for tag, flags in (("host", []), ("dev", ["--device-ingest"])):
        ctl = casefiles.write_case(os.path.join(str(tmp_path), tag), net, params, opts, ro, case_name="ing", forcing_dt=None if same else forcing_dt,
                                   sim_steps=None if same else sim_steps, shuffle_hru_seed=7 if same else None, extra_keys=keys)
        r = subprocess.run([_routing_host(backend), ctl, "--batch", "5"] + flags, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        outs.append(casefiles.read_history(json.loads(r.stdout.strip().splitlines()[-1])["history"]))
    assert set(outs[0]) == set(outs[1])
    for v in outs[0]:
        assert np.array_equal(outs[0][v], outs[1][v]), v
    assert ("basRunoff" in outs[0]) == (bas == "T")


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("route_opt,freq,batch,restart", [("12", "daily", 25, "never"), ("012", "5", 7, "never"), ("1", "daily", 64, "daily")])
def test_device_history_gives_the_same_files_as_host_aggregation(tmp_path, backend, route_opt, freq, batch, restart):
    """--device-history: the period means of the discharges and of dlayRunoff are formed on the device (mr_history_means) and
    only they travel -- periods that span batches, a last period cut short by the end of the run; with restart files due inside
    an output period the host aggregates instead (their history state is the host's); the history files equal the ones of the
    default path byte for byte."""
    net, params, opts, ro = case("conus", n=260, seed=6, dt=3600.0, route_opt=route_opt, steps=60)
    outs = []
    for tag, flags in (("host", []), ("dev", ["--device-history"])):
        ctl = casefiles.write_case(os.path.join(str(tmp_path), tag), net, params, opts, ro, case_name="dh", output_frequency=freq, restart_write=restart,
                                   start="2001-02-27 06:00:00")
        r = subprocess.run([_routing_host(backend), ctl, "--batch", str(batch)] + flags, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        # restart files inside an output period need the running sums on the host: the flag yields then, and says so
        assert ("the host aggregates" in r.stderr) == (bool(flags) and restart != "never"), r.stderr
        outs.append(casefiles.read_history(json.loads(r.stdout.strip().splitlines()[-1])["history"]))
    assert set(outs[0]) == set(outs[1]) and "dlayRunoff" in outs[0]
    for v in outs[0]:
        assert np.array_equal(outs[0][v], outs[1][v]), v


@pytest.mark.parametrize("backend", BACKENDS)
def test_history_at_gauges_only(tmp_path, backend):
    """<outputAtGage> T with <gageMetaFile>: the history files hold the gauged reaches only, in the order of the csv (gauges
    outside the network dropped) -- the columns of the full output."""
    net, params, opts, ro = case("random", n=70, seed=3, dt=3600.0, route_opt="01", steps=12)
    rng = np.random.default_rng(1)
    rch = rng.choice(net.nRch, 6, replace=False)
    gage_ids = ["Q%02d" % i for i in range(7)]
    csv_rch = list(net.segId[rch[:3]]) + [int(net.segId.max()) + 5] + list(net.segId[rch[3:]])      # the fourth gauge is not on the network
    d = str(tmp_path)
    full = casefiles.write_case(os.path.join(d, "full"), net, params, opts, ro, case_name="g", extra_keys={"IRFvolume": "T"})
    # the gauge files are written by the `gauges` option; qmodOption is switched back off so that only the output changes
    sub = casefiles.write_case(os.path.join(d, "sub"), net, params, opts, ro, case_name="g",
                               gauges=(gage_ids, gage_ids, csv_rch, [0.0], np.full((1, 7), np.nan), 5, 1),
                               extra_keys={"IRFvolume": "T", "outputAtGage": "T", "qmodOption": "0"})
    outs = []
    for ctl in (full, sub):
        r = subprocess.run([_routing_host(backend), ctl, "--batch", "5"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        outs.append(casefiles.read_history(json.loads(r.stdout.strip().splitlines()[-1])["history"]))
    assert np.array_equal(outs[1]["reachID"], net.segId[rch])
    for v in ("sumUpstreamRunoff", "IRFroutedRunoff", "dlayRunoff", "IRFvolume"):
        assert np.array_equal(outs[1][v], outs[0][v][:, rch]), v
    assert np.array_equal(outs[1]["basRunoff"], outs[0]["basRunoff"])                  # HRU-dimensioned output is not subset


_REF = "/root/reference/route"


@pytest.mark.skipif(not os.path.exists(os.path.join(_REF, "settings", "SAMPLE.control")), reason="reference checkout not present (CPU container only)")
def test_reference_sample_control_file_drives_the_host(tmp_path):
    """Drop-in check of the control-file reader: the reference's own route/settings/SAMPLE.control -- every line, comments and
    layout as shipped -- with only its placeholder values (CASE_NAME, NTOPO_NC, ...) filled in, and its param.nml.default,
    runs through this host (stand-in library) and gives the oracle's flows for the options the sample asks for
    (route_opt 5, daily steps, daily output, monthly files)."""
    import re
    import shutil
    from oracle.oracle import Oracle
    net, params, opts, ro = case("random", n=60, seed=3, dt=86400.0, route_opt="5", steps=40)
    opts.units_qsim = "mm/s"
    ro = ro * 1000.0                                                                  # the sample's <units_qsim> is mm/s
    ctl0 = casefiles.write_case(str(tmp_path), net, params, opts, ro, case_name="sample", restart_write="last")
    mine = dict(re.findall(r"^<([A-Za-z0-9_]+)>\s+(.*?)\s*!", open(ctl0).read(), flags=re.M))
    anc = mine["ancil_dir"]
    shutil.copy(os.path.join(_REF, "ancillary_data", "param.nml.default"), os.path.join(anc, "param.nml.default"))
    fill = {"param_nml": "param.nml.default", "is_remap": "F", "fname_state_in": "coldstart", "seg_outlet": "-9999",
                 "varname_area": "area", "varname_length": "length", "varname_slope": "slope", "varname_HRUid": "HRUid",
                 "varname_hruSegId": "hruSegId", "varname_segId": "segId", "varname_downSegId": "downSegId",
                 "varname_islake": "islake", "varname_lakeModelType": "lakeModelType"}
    placeholder = lambda v: bool(re.fullmatch(r"[A-Z][A-Z_0-9]*", v)) or v.startswith("yyyy")
    lines = []
    for line in open(os.path.join(_REF, "settings", "SAMPLE.control")):
        m = re.match(r"^(<([A-Za-z0-9_]+)>\s+)(\S.*?)(\s*!.*)$", line.rstrip("\n"))
        if m and (m.group(2) in fill or (placeholder(m.group(3).strip()) and m.group(2) in mine)):
            line = m.group(1) + str(fill.get(m.group(2), mine.get(m.group(2)))) + "   " + m.group(4).strip() + "\n"
        lines.append(line)
    ctl = os.path.join(str(tmp_path), "SAMPLE_filled.control")
    open(ctl, "w").writelines(lines)
    kept = dict(re.findall(r"^<([A-Za-z0-9_]+)>\s+(.*?)\s*!", "".join(lines), flags=re.M))
    assert kept["route_opt"] == "5" and kept["outputFrequency"] == "daily" and kept["newFileFrequency"] == "monthly"      # the sample's own choices
    r = subprocess.run([_routing_host("oracle-stub"), ctl, "--batch", "16"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    info = json.loads(r.stdout.strip().splitlines()[-1])
    files = info.get("history_files", [info["history"]])
    assert [os.path.basename(f) for f in files] == ["sample.h.2000-01.nc", "sample.h.2000-02.nc"]                  # get_hfilename, monthly
    out = np.concatenate([casefiles.read_history(f)["DWroutedRunoff"] for f in files])
    qo = Oracle(net, params, opts).run(ro)
    np.testing.assert_allclose(out, qo[0].astype(np.float32), rtol=2e-6, atol=1e-30)


@pytest.mark.skipif(not os.path.exists(os.path.join(_REF, "build", "src", "public_var.f90")), reason="reference checkout not present (CPU container only)")
def test_option_defaults_equal_the_reference_declarations():
    """The defaults of RouteOptions / RouteParams (and so of every test case and of the host, which share them) against the
    declarations in public_var.f90 and the namelist the reference ships (param.nml.default)."""
    import re
    from mizuroute_b200 import capi
    src = open(os.path.join(_REF, "build", "src", "public_var.f90")).read()

    def declared(name):
        m = re.search(r"::\s*%s\s*=\s*([^!\n]+)" % re.escape(name), src)
        assert m, name
        v = m.group(1).strip().lower()
        if v in (".true.", ".false."):
            return v == ".true."
        return float(re.sub(r"_dp|_i4b", "", v).replace("d", "e"))
    o = RouteOptions()
    for name in ("doesBasinRoute", "hw_drain_point", "min_length_route", "is_lake_sim", "lakeRegulate", "LakeInputOption", "runoffMin", "floodplain"):
        assert float(getattr(o, name)) == float(declared(name)), name
    assert declared("MAXQPAR") == capi.MR_KW_SLOTS - 2 == 20 and declared("negRunoffTol") == -1e-3
    assert (declared("qBlendPeriod"), declared("QerrTrend"), declared("qmodOption")) == (10, 1, 0)
    nml = open(os.path.join(_REF, "ancillary_data", "param.nml.default")).read()
    p = RouteParams()
    for name in ("fshape", "tscale", "velo", "diff", "mann_n", "wscale"):
        assert float(re.search(r"%s\s*=\s*([0-9.eE+-]+)" % name, nml).group(1)) == getattr(p, name), name


@pytest.mark.skipif(not os.path.exists(os.path.join(_REF, "build", "src", "read_control.f90")), reason="reference checkout not present (CPU container only)")
def test_every_control_key_of_the_reference_is_accepted():
    """read_control.f90 stops at an unknown key (:374-377), and so does this host -- so every key the reference's reader knows
    (and every key its sample control files use) must be known here, or a working control file would be refused."""
    import glob
    import re
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mizuroute_b200", "csrc", "route_runoff.cpp")).read()
    i = src.index("KNOWN_KEYS[] = {")
    known = set(re.findall(r'"([A-Za-z0-9_]+)"', src[i:src.index("};", i)]))
    free = re.compile(r"(varname_|vname_|dname_|fname_)")                         # name keys are accepted by prefix
    ref = set(re.findall(r"case\('<([A-Za-z0-9_]+)>'", open(os.path.join(_REF, "build", "src", "read_control.f90")).read()))
    assert len(ref) > 100
    for f in glob.glob(os.path.join(_REF, "settings", "*.control")):
        ref |= set(re.findall(r"^<([A-Za-z0-9_]+)>", open(f).read(), flags=re.M))
    assert sorted(k for k in ref if k not in known and not free.match(k)) == []
