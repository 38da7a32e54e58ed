"""Writes a stand-alone mizuRoute case in the reference's file formats: control file (`<key> value ! comment`,
route/settings/SAMPLE.control), parameter namelist (route/ancillary_data/param.nml.default), river-network netCDF
(read_streamSeg.f90:44) and runoff netCDF [time, hru] (read_runoff.f90) -- NetCDF-3 64-bit offset through
scipy.io.netcdf_file (there is no libnetcdf in this image).  Used by the synthetic-configuration generators and the
host tests; `mizuroute_b200/route_runoff <control>` runs the case."""
from __future__ import annotations

import os

import numpy as np
from scipy.io import netcdf_file

from .network import RiverNetwork, RouteOptions, RouteParams


def write_network(path: str, net: RiverNetwork):
    f = netcdf_file(path, "w", version=2)
    f.createDimension("seg", net.nRch)
    f.createDimension("hru", net.nHRU)

    def put(name, data, dim, typ):
        v = f.createVariable(name, typ, (dim,))
        v[:] = data

    put("segId", net.segId, "seg", "i"); put("downSegId", net.downSegId, "seg", "i")
    put("length", net.length, "seg", "d"); put("slope", net.slope, "seg", "d")
    put("HRUid", net.hruId, "hru", "i"); put("hruSegId", net.hruSegId, "hru", "i"); put("area", net.area, "hru", "d")
    if net.width is not None:
        put("width", net.width, "seg", "d")
    if net.man_n is not None:
        put("man_n", net.man_n, "seg", "d")
    if net.islake is not None:
        put("islake", net.islake, "seg", "i")
        put("lakeModelType", net.lakeModelType if net.lakeModelType is not None else np.ones(net.nRch, np.int32), "seg", "i")
        for k in ("D03_MaxStorage", "D03_Coefficient", "D03_Power", "D03_S0"):
            if getattr(net, k) is not None:
                put(k, getattr(net, k), "seg", "d")
        for k, v in (getattr(net, "lake_params", None) or {}).items():      # HYP_* etc. by their reference names
            put(k, v, "seg", "d")
    f.close()


def write_runoff(path: str, hru_ids: np.ndarray, runoff: np.ndarray, dt: float, start: str = "2000-01-01 00:00:00",
                 t_offset_steps: int = 0, dtype: str = "d", extra=None):
    """runoff[time, hru]; time in seconds since `start` (stamped at the start of each step).  `extra` = {name: array[time, hru]}
    adds further forcing variables (lake evaporation / precipitation)."""
    f = netcdf_file(path, "w", version=2)
    f.createDimension("time", None)
    f.createDimension("hru", len(hru_ids))
    t = f.createVariable("time", "d", ("time",))
    t.units = "seconds since " + start
    t.calendar = "standard"
    h = f.createVariable("hruid", "i", ("hru",))
    h[:] = hru_ids
    q = f.createVariable("runoff", dtype, ("time", "hru"))
    q.units = "mm/s"
    more = {}
    for nm in (extra or {}):
        more[nm] = f.createVariable(nm, dtype, ("time", "hru"))
        more[nm].units = "mm/s"
    for k in range(runoff.shape[0]):
        t[k] = (t_offset_steps + k) * dt
        q[k, :] = runoff[k]
        for nm, v in more.items():
            v[k, :] = extra[nm][k]
    f.close()


def write_wm(path: str, seg_ids: np.ndarray, flux, vol, dt: float, start: str = "2000-01-01 00:00:00"):
    """Water-management file: flux_wm / vol_wm [time, seg] (either may be None), reach ids, time in seconds since `start`."""
    f = netcdf_file(path, "w", version=2)
    f.createDimension("time", None)
    f.createDimension("seg", len(seg_ids))
    t = f.createVariable("time", "d", ("time",))
    t.units = "seconds since " + start
    t.calendar = "standard"
    sid = f.createVariable("seg_id", "i", ("seg",))
    sid[:] = seg_ids
    vf = f.createVariable("flux_wm", "d", ("time", "seg")) if flux is not None else None
    vv = f.createVariable("vol_wm", "d", ("time", "seg")) if vol is not None else None
    n = (flux if flux is not None else vol).shape[0]
    for k in range(n):
        t[k] = k * dt
        if vf is not None:
            vf[k, :] = flux[k]
        if vv is not None:
            vv[k, :] = vol[k]
    f.close()


def write_runoff_grid(path: str, runoff: np.ndarray, dt: float, start: str = "2000-01-01 00:00:00", fill=None):
    """Gridded runoff[time, lat, lon] (the layout read_2D_forcing expects, read_runoff.f90:331-396)."""
    f = netcdf_file(path, "w", version=2)
    f.createDimension("time", None)
    f.createDimension("lat", runoff.shape[1])
    f.createDimension("lon", runoff.shape[2])
    t = f.createVariable("time", "d", ("time",))
    t.units = "seconds since " + start
    t.calendar = "standard"
    q = f.createVariable("runoff", "d", ("time", "lat", "lon"))
    q.units = "mm/s"
    if fill is not None:
        q._FillValue = float(fill)
    for k in range(runoff.shape[0]):
        t[k] = k * dt
        q[k] = runoff[k]
    f.close()


def _stamp(seconds: float, start: str) -> str:
    import datetime as _dt
    t0 = _dt.datetime.strptime(start, "%Y-%m-%d %H:%M:%S")
    return (t0 + _dt.timedelta(seconds=seconds)).strftime("%Y-%m-%d %H:%M:%S")


def write_gauges(csv_path: str, nc_path: str, site_ids, gage_ids, reach_ids, times_sec, flow, start: str = "2000-01-01 00:00:00", strlen: int = 30,
                 fill: float = -9999.0):
    """Gauge metadata csv (<gageMetaFile>: header row, columns gage_id, reach_id, as gageMeta_data.f90:58-68 reads them) and the
    gauge netCDF (<fname_gageObs>: flow [time, site], site names as a character variable [site, strlen], time in seconds since
    `start` -- records only at `times_sec`; NaN in `flow` is written as the fill value).  `site_ids` are the gauges of the netCDF,
    `gage_ids` / `reach_ids` the rows of the csv (they need not hold the same gauges)."""
    with open(csv_path, "w") as f:
        f.write("gage_id,reach_id,lat,lon\n")
        for g, r in zip(gage_ids, reach_ids):
            f.write("%s,%d,0.0,0.0\n" % (g, int(r)))
    f = netcdf_file(nc_path, "w", version=2)
    f.createDimension("time", None)
    f.createDimension("site", len(site_ids))
    f.createDimension("strlen", strlen)
    t = f.createVariable("time", "d", ("time",))
    t.units = "seconds since " + start
    t.calendar = "standard"
    sv = f.createVariable("site", "c", ("site", "strlen"))
    for i, g in enumerate(site_ids):
        sv[i, :] = np.frombuffer(str(g).ljust(strlen)[:strlen].encode(), dtype="S1")
    q = f.createVariable("flow", "d", ("time", "site"))
    q._FillValue = fill
    for k, ts in enumerate(times_sec):
        t[k] = float(ts)
        q[k, :] = np.where(np.isnan(flow[k]), fill, flow[k])
    f.close()


def write_case(case_dir: str, net: RiverNetwork, params: RouteParams, opts: RouteOptions, runoff: np.ndarray, case_name: str = "case",
               start: str = "2000-01-01 00:00:00", split_forcing: int = 1, shuffle_hru_seed=None, restart_write: str = "never",
               fname_state_in: str = "coldstart", first_step: int = 0, remap=None, output_frequency="1", forcing_dt=None, sim_steps=None,
               ro_time_stamp=None, new_file_frequency="single", extra_keys=None, lake_forcing=None, wm=None, gauges=None) -> str:
    """Creates <case_dir>/{ancillary,input,output} and returns the control-file path.  `first_step` > 0 writes a
    continuation run: the forcing records and <sim_start> begin `first_step` steps after `start`.  `forcing_dt` != dt_qsim
    writes the runoff records on their own interval (`sim_steps` simulation steps of opts.dt are then asked for);
    `ro_time_stamp` = start|middle|end shifts the record time stamps within their interval."""
    t_first = _stamp(first_step * opts.dt, start)
    anc, inp, out = (os.path.join(case_dir, d) + "/" for d in ("ancillary", "input", "output"))
    for d in (anc, inp, out):
        os.makedirs(d, exist_ok=True)
    write_network(anc + "ntopo.nc", net)
    with open(anc + "param.nml", "w") as f:
        f.write("&HSLOPE\n  ! hillslope gamma UH\n  fshape = %r\n  tscale = %r\n/\n&IRF_UH\n  velo = %r\n  diff = %r\n/\n&KWT\n  mann_n = %r\n  wscale = %r\n/\n"
                % (params.fshape, params.tscale, params.velo, params.diff, params.mann_n, params.wscale))
    ids, ro = net.hruId, runoff
    grid = remap is not None and isinstance(remap[0], str) and remap[0] == "grid"
    if grid:                                                # ("grid", map_ids, num_qhru, i_index, j_index, weight): runoff[time, lat, lon]
        _, map_ids, num_q, i_index, j_index, wgt = remap
        f = netcdf_file(anc + "remap.nc", "w", version=2)
        f.createDimension("hru", len(map_ids)); f.createDimension("data", len(wgt))
        for nm, dat, dim, typ in (("RN_hruId", map_ids, "hru", "i"), ("nOverlaps", num_q, "hru", "i"), ("i_index", i_index, "data", "i"),
                                  ("j_index", j_index, "data", "i"), ("weight", wgt, "data", "d")):
            v = f.createVariable(nm, typ, (dim,)); v[:] = dat
        f.close()
    elif remap is not None:                                 # (map_ids, num_qhru, qhru_ids, weight, forcing_ids): runoff is on forcing polygons
        map_ids, num_q, q_ids, wgt, ids = remap
        f = netcdf_file(anc + "remap.nc", "w", version=2)
        f.createDimension("hru", len(map_ids)); f.createDimension("data", len(q_ids))
        for nm, dat, dim, typ in (("RN_hruId", map_ids, "hru", "i"), ("nOverlaps", num_q, "hru", "i"), ("overlapPolyId", q_ids, "data", "i"), ("weight", wgt, "data", "d")):
            v = f.createVariable(nm, typ, (dim,)); v[:] = dat
        f.close()
    elif shuffle_hru_seed is not None:                       # forcing HRUs in a different order than the network's
        perm = np.random.default_rng(shuffle_hru_seed).permutation(net.nHRU)
        ids, ro = ids[perm], runoff[:, perm]
    K = runoff.shape[0] if sim_steps is None else sim_steps
    dt_ro = opts.dt if forcing_dt is None else float(forcing_dt)
    if grid:
        assert split_forcing <= 1 and first_step == 0
        write_runoff_grid(inp + "runoff_%s.nc" % case_name, runoff, dt_ro, start)
        fname_qsim = "runoff_%s.nc" % case_name
    elif forcing_dt is not None or ro_time_stamp is not None:
        assert split_forcing <= 1 and first_step == 0
        shift = {None: 0.0, "start": 0.0, "middle": 0.5, "end": 1.0}[ro_time_stamp]
        write_runoff(inp + "runoff_%s.nc" % case_name, ids, ro, dt_ro, start, t_offset_steps=shift)
        fname_qsim = "runoff_%s.nc" % case_name
    elif split_forcing <= 1:
        extra = None
        if lake_forcing is not None:                         # (evapo[time, hru], precip[time, hru]) in river-network HRU order
            assert remap is None and shuffle_hru_seed is None
            extra = {"evapo": lake_forcing[0], "precip": lake_forcing[1]}
        write_runoff(inp + "runoff_%s.nc" % case_name, ids, ro, opts.dt, start, t_offset_steps=first_step, extra=extra)
        fname_qsim = "runoff_%s.nc" % case_name
    else:
        bounds = np.linspace(0, runoff.shape[0], split_forcing + 1).astype(int)
        names = []
        for i in range(split_forcing):
            nm = "runoff_%02d.nc" % i
            write_runoff(inp + nm, ids, ro[bounds[i]:bounds[i + 1]], opts.dt, start, t_offset_steps=first_step + int(bounds[i]))
            names.append(nm)
        with open(inp + "runoff_files.txt", "w") as f:
            f.write("! forcing files in chronological order\n" + "\n".join(names) + "\n")
        fname_qsim = "runoff_files.txt"
    keys = [
        ("case_name", case_name, "name of simulation"),
        ("ancil_dir", anc, "directory containing ancillary data"),
        ("input_dir", inp, "directory containing input data"),
        ("output_dir", out, "directory containing output data"),
        ("sim_start", t_first, "time of simulation start"),
        ("sim_end", _stamp((first_step + K - 1) * opts.dt, start), "time of simulation end"),
        ("route_opt", opts.route_opt, "routing schemes"),
        ("doesBasinRoute", opts.doesBasinRoute, "hillslope routing"),
        ("dt_qsim", int(opts.dt), "simulation time interval [sec]"),
        ("hw_drain_point", opts.hw_drain_point, "lateral runoff to headwater reaches"),
        ("min_length_route", opts.min_length_route, "minimum reach length for routing"),
        ("is_lake_sim", "T" if opts.is_lake_sim else "F", "lake simulation"),
        ("lakeRegulate", "T" if opts.lakeRegulate else "F", "parametric lake models"),
        ("LakeInputOption", opts.LakeInputOption, "fluxes for lake simulation"),
        ("runoffMin", repr(float(opts.runoffMin)), "minimum runoff [m3/s]"),
        ("fname_ntopOld", "ntopo.nc", "river network netCDF"),
        ("dname_sseg", "seg", "dimension of segments"),
        ("dname_nhru", "hru", "dimension of HRUs"),
        ("fname_qsim", fname_qsim, "runoff netCDF or list of netCDFs"),
        ("vname_qsim", "runoff", "runoff variable"),
        ("vname_time", "time", "time variable"),
        ("vname_evapo", "evapo", "lake evaporation variable (<is_lake_sim> T, <LakeInputOption> 0 or 2)"),
        ("vname_precip", "precip", "lake precipitation variable"),
        ("vname_hruid", "hruid", "forcing HRU id variable"),
        ("dname_time", "time", "time dimension"),
        ("dname_hruid", "hru", "HRU dimension"),
        ("units_qsim", opts.units_qsim, "units of runoff"),
        ("dt_ro", int(dt_ro), "forcing interval [sec]"),
        ("ro_time_stamp", ro_time_stamp or "start", "time stamp of a forcing record within its interval"),
        ("is_remap", "T" if remap is not None else "F", "runoff remapping"),
        ("fname_remap", "remap.nc", "runoff mapping netCDF"),
        ("vname_hruid_in_remap", "RN_hruId", "river-network HRU ids in the mapping"),
        ("vname_weight", "weight", "areal weights"),
        ("vname_qhruid", "overlapPolyId", "forcing polygon ids"),
        ("vname_num_qhru", "nOverlaps", "overlapping polygons per river-network HRU"),
        ("dname_hru_remap", "hru", "mapping HRU dimension"),
        ("dname_data_remap", "data", "mapping data dimension"),
        ("vname_i_index", "i_index", "x (lon) index of the overlapping grid cells, 1-based"),
        ("vname_j_index", "j_index", "y (lat) index of the overlapping grid cells, 1-based"),
        ("dname_xlon", "lon", "x dimension of gridded runoff"),
        ("dname_ylat", "lat", "y dimension of gridded runoff"),
        ("param_nml", "param.nml", "spatially constant parameters"),
        ("restart_write", restart_write, "restart write option"),
        ("fname_state_in", fname_state_in, "input restart netCDF ('coldstart' = none)"),
        ("newFileFrequency", new_file_frequency, "history file frequency: single, daily, monthly or yearly"),
        ("outputFrequency", output_frequency, "output frequency: number of steps or daily"),
    ]
    if wm is not None:                                      # (seg_ids, flux_wm[time, seg] or None, vol_wm[time, seg] or None)
        wm_ids, wm_flux, wm_vol = wm
        write_wm(inp + "wm_%s.nc" % case_name, wm_ids, wm_flux, wm_vol, opts.dt, t_first)
        keys += [("is_flux_wm", "T" if wm_flux is not None else "F", "abstraction / injection fluxes"),
                 ("is_vol_wm", "T" if wm_vol is not None else "F", "target lake volumes"), ("is_vol_wm_jumpstart", "T", "start the target-volume lakes at their target"),
                 ("fname_wm", "wm_%s.nc" % case_name, "water-management netCDF"), ("vname_flux_wm", "flux_wm", ""), ("vname_vol_wm", "vol_wm", ""),
                 ("vname_time_wm", "time", ""), ("vname_segid_wm", "seg_id", ""), ("dname_time_wm", "time", ""), ("dname_segid_wm", "seg", "")]
    if gauges is not None:                                  # (site_ids, csv gage_ids, csv reach_ids, times_sec, flow[time, site], qBlendPeriod, QerrTrend)
        g_sites, g_ids, g_rch, g_t, g_flow, blend, trend = gauges
        write_gauges(anc + "gauges_%s.csv" % case_name, anc + "gauge_obs_%s.nc" % case_name, g_sites, g_ids, g_rch, g_t, g_flow, t_first)
        keys += [("qmodOption", 1, "direct insertion"), ("qBlendPeriod", int(blend), "steps over which the correction fades"),
                 ("QerrTrend", int(trend), "1 constant, 2 linear, 3 logistic, 4 exponential"),
                 ("gageMetaFile", "gauges_%s.csv" % case_name, "gauge metadata csv"), ("fname_gageObs", "gauge_obs_%s.nc" % case_name, "gauge netCDF"),
                 ("vname_gageFlow", "flow", ""), ("vname_gageSite", "site", ""), ("vname_gageTime", "time", ""), ("dname_gageSite", "site", ""),
                 ("dname_gageTime", "time", ""), ("strlen_gageSite", 30, "")]
    for k, v in (extra_keys or {}).items():
        keys.append((k, v, "extra key"))
    ctl = os.path.join(case_dir, case_name + ".control")
    with open(ctl, "w") as f:
        f.write("! mizuRoute control file written by mizuroute_b200.casefiles\n! format: <key>  value  ! comment\n")
        for k, v, cmt in keys:
            f.write("%-24s %-40s ! %s\n" % ("<" + k + ">", v, cmt))
    return ctl


def read_history(path: str) -> dict:
    f = netcdf_file(path, "r", mmap=False)
    out = {k: np.array(v[:]) for k, v in f.variables.items()}
    out = {k: a.astype(a.dtype.newbyteorder("=")) for k, a in out.items()}       # netCDF is big-endian on disk
    f.close()
    return out
