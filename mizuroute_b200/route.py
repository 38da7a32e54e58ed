"""Host-side mirror of the reference's routing interface on top of the C ABI.

`Router.main_route(basinRunoff)` is one call of the reference's `main_route` (main_route.f90:29-268) for the
current time step `TSEC` (globalData.f90:112); the caller then advances time exactly as `update_time` does
(init_model_data.f90:311-312).  `Router.route_batch(runoff)` is the same for a block of steps in one
time-skewed wavefront on the device.  Errors surface as `RoutingError(ierr, message)` -- the reference's
`ierr, message` pair -- instead of `handle_err -> MPI_Abort` (model_utils.f90:45-57).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import capi
from .network import RiverNetwork, RouteOptions, RouteParams


class RoutingError(RuntimeError):
    def __init__(self, ierr: int, message: str):
        super().__init__(f"ierr={ierr}: {message}")
        self.ierr = ierr
        self.message = message


def _ptr(a: Optional[np.ndarray], ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


class Router:
    def __init__(self, net: RiverNetwork, params: RouteParams, opts: RouteOptions, device: int = 0, max_batch: int = 1,
                 ghosts=None):
        """`ghosts` = (segId[], kind[], totArea[], width[]) marks reaches of `net` that are copies of tributary outlets
        routed in another domain (mr_set_ghosts)."""
        self._L = capi.load()
        self._h = C.c_void_p()
        self._msg = C.create_string_buffer(capi.MR_STRLEN)
        self.net, self.params, self.opts = net, params, opts
        self.methods = [int(c) for c in opts.route_opt]          # read_control.f90:583-597
        tc, lc = opts.conv()
        o = capi.mr_options()
        o.dt = float(opts.dt)
        o.n_routes = len(self.methods)
        for i, m in enumerate(self.methods[:8]):
            o.route_methods[i] = m
        o.doesBasinRoute = int(opts.doesBasinRoute)
        o.hw_drain_point = int(opts.hw_drain_point)
        o.min_length_route = float(opts.min_length_route)
        o.is_lake_sim = int(bool(opts.is_lake_sim))
        o.lakeRegulate = int(bool(opts.lakeRegulate))
        o.LakeInputOption = int(opts.LakeInputOption)
        o.runoffMin = float(opts.runoffMin)
        o.time_conv, o.length_conv = tc, lc
        o.fshape, o.tscale = params.fshape, params.tscale
        o.velo, o.diff = params.velo, params.diff
        o.mann_n, o.wscale = params.mann_n, params.wscale
        o.device = int(device)
        o.max_batch = int(max_batch)
        o.floodplain = int(bool(getattr(opts, "floodplain", False)))     # dscale / floodplainSlope stay 0 = the reference's constants
        self.max_batch = int(max_batch)
        self._check(self._L.mr_create(C.byref(o), C.byref(self._h), self._msg))
        n = net
        for name, vals in (getattr(net, "lake_params", None) or {}).items():     # HYPE reservoirs etc., before mr_set_network
            v = np.ascontiguousarray(vals, dtype=np.float64)
            self._check(self._L.mr_set_lake_param(self._h, name.encode(), int(v.size), _ptr(v, C.c_double), self._msg))
        if getattr(opts, "sim_start", None):
            y, mo, d, sec = opts.sim_start
            self._check(self._L.mr_set_sim_start(self._h, int(y), int(mo), int(d), float(sec), int(opts.calendar == "noleap"), self._msg))
        if ghosts is not None:
            gid, gkind, garea, gwidth = (np.ascontiguousarray(ghosts[0], dtype=np.int32), np.ascontiguousarray(ghosts[1], dtype=np.int32),
                                         np.ascontiguousarray(ghosts[2], dtype=np.float64), np.ascontiguousarray(ghosts[3], dtype=np.float64))
            self._check(self._L.mr_set_ghosts(self._h, len(gid), _ptr(gid, C.c_int), _ptr(gkind, C.c_int), _ptr(garea, C.c_double),
                                              _ptr(gwidth, C.c_double), self._msg))
        self._check(self._L.mr_set_network(
            self._h, n.nRch, n.nHRU, _ptr(n.segId, C.c_int), _ptr(n.downSegId, C.c_int), _ptr(n.hruSegId, C.c_int),
            _ptr(n.area, C.c_double), _ptr(n.length, C.c_double), _ptr(n.slope, C.c_double), _ptr(n.width, C.c_double),
            _ptr(n.man_n, C.c_double), _ptr(n.islake, C.c_int), _ptr(n.lakeModelType, C.c_int),
            _ptr(n.D03_MaxStorage, C.c_double), _ptr(n.D03_Coefficient, C.c_double), _ptr(n.D03_Power, C.c_double),
            _ptr(n.D03_S0, C.c_double), self._msg))
        self.nRch, self.nHRU = n.nRch, n.nHRU
        self.TSEC = [0.0, float(opts.dt)]                        # init_model_data.f90:600

    # ------------------------------------------------------------------------------------------
    def _check(self, ierr: int):
        if ierr != 0:
            raise RoutingError(ierr, self._msg.value.decode(errors="replace"))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.mr_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _advance(self, n: int):
        for _ in range(n):
            self.TSEC[0] = self.TSEC[1]
            self.TSEC[1] = self.TSEC[0] + float(self.opts.dt)

    # ------------------------------------------------------------------------------------------
    def main_route(self, basinRunoff: np.ndarray):
        """One reference `main_route` call for [TSEC(1),TSEC(2)], then `update_time`."""
        r = np.ascontiguousarray(basinRunoff, dtype=np.float64)
        if r.shape != (self.nHRU,):
            raise ValueError("basinRunoff must have one value per HRU")
        self._check(self._L.mr_step(self._h, self.TSEC[0], self.TSEC[1], _ptr(r, C.c_double), self._msg))
        self._advance(1)

    def route_batch(self, runoff, out=None, want_q: bool = True):
        """Route runoff[K, nHRU]; returns REACH_Q[n_routes, K, nRch] (caller's reach order) or None.

        `runoff` / `out` may be numpy arrays or anything exposing `data_ptr()` (pinned torch tensors)."""
        K, rp = self._host_ptr(runoff, self.nHRU)
        if want_q and out is None:
            out = np.empty((len(self.methods), K, self.nRch))
        op = None
        if want_q:
            _, op = self._host_ptr(out, self.nRch, rows=len(self.methods) * K)
        self._check(self._L.mr_step_batch(self._h, K, self.TSEC[0], rp, op, self._msg))
        self._advance(K)
        return out if want_q else None

    def route_batch_async(self, runoff, out=None):
        """`route_batch` without host waits (mr_step_batch_async): upload, routing and download of consecutive calls
        overlap.  `runoff` / `out` must be pinned host buffers that stay alive until `wait()`."""
        K, rp = self._host_ptr(runoff, self.nHRU)
        op = None
        if out is not None:
            _, op = self._host_ptr(out, self.nRch, rows=len(self.methods) * K)
        self._check(self._L.mr_step_batch_async(self._h, K, self.TSEC[0], rp, op, self._msg))
        self._advance(K)

    def upload_wm(self, flux_wm=None, vol_wm=None, vol_jumpstart: bool = False):
        """Water management [K, nRch] of the NEXT routing call (mr_upload_wm): abstraction (+) / injection (-) fluxes, -9999 =
        none, and / or target volumes of the lakes flagged by the lake parameter "LakeTargVol"."""
        f = None if flux_wm is None else np.ascontiguousarray(np.atleast_2d(flux_wm), dtype=np.float64)
        v = None if vol_wm is None else np.ascontiguousarray(np.atleast_2d(vol_wm), dtype=np.float64)
        k = (f if f is not None else v).shape[0]
        assert all(a is None or a.shape == (k, self.nRch) for a in (f, v))
        self._check(self._L.mr_upload_wm(self._h, int(k), _ptr(f, C.c_double), _ptr(v, C.c_double), int(bool(vol_jumpstart)), self._msg))

    def set_ingest(self, n_forcing: int, forcing_of_hru, scale: float = -9999.0, offset: float = -9999.0, fill: float = -9999.0):
        """Forcing ingest on the device (mr_set_ingest): forcing_of_hru [nHRU] = column of every river-network HRU in the raw
        forcing records (-1 = none); scale / offset = <scale_factor_runoff> / <offset_value_runoff> (-9999 = not given)."""
        f = np.ascontiguousarray(forcing_of_hru, dtype=np.int32)
        assert f.shape == (self.nHRU,)
        self._check(self._L.mr_set_ingest(self._h, int(n_forcing), _ptr(f, C.c_int), float(scale), float(offset), float(fill), self._msg))
        self._n_forcing = int(n_forcing)

    def ingest_records(self, records, rec_ptr, rec_idx, rec_frac=None) -> int:
        """Raw forcing records [nRec, nForcing] -> the runoff rows of K = len(rec_ptr) - 1 steps, resident on the device
        (mr_ingest_records; time-weighted mean over the records under a step, scale / offset, sort_flux).  Follow with
        route_resident(K)."""
        r = np.ascontiguousarray(records, dtype=np.float64)
        p = np.ascontiguousarray(rec_ptr, dtype=np.int32); i = np.ascontiguousarray(rec_idx, dtype=np.int32)
        f = None if rec_frac is None else np.ascontiguousarray(rec_frac, dtype=np.float64)
        assert r.ndim == 2 and r.shape[1] == self._n_forcing and i.size == p[-1] and (f is None or f.size == i.size)
        self._check(self._L.mr_ingest_records(self._h, int(p.size - 1), int(r.shape[0]), _ptr(r, C.c_double), _ptr(p, C.c_int), _ptr(i, C.c_int),
                                              _ptr(f, C.c_double), self._msg))
        return int(p.size - 1)

    def set_da(self, qmod_option: int = 1, q_blend_period: int = 10, q_err_trend: int = 1):
        """Data assimilation by direct insertion (<qmodOption>, <qBlendPeriod>, <QerrTrend>; mr_set_da)."""
        self._check(self._L.mr_set_da(self._h, int(qmod_option), int(q_blend_period), int(q_err_trend), self._msg))

    def upload_obs(self, obs, has_record=None):
        """Gauge observations [K, nRch] (m3/s; NaN or negative = none at that reach) of the NEXT routing call, which must
        route K steps (mr_upload_obs); has_record [K], 0 = the gauge file has no record at that step."""
        o = np.ascontiguousarray(np.atleast_2d(obs), dtype=np.float64)
        assert o.shape[1] == self.nRch
        r = None if has_record is None else np.ascontiguousarray(has_record, dtype=np.int32)
        assert r is None or r.shape == (o.shape[0],)
        self._check(self._L.mr_upload_obs(self._h, int(o.shape[0]), _ptr(r, C.c_int), _ptr(o, C.c_double), self._msg))

    def upload_lake_forcing(self, evapo, precip):
        """Lake evaporation / precipitation [K, nHRU] (runoff units, river-network HRU order) of the NEXT routing call, which
        must route K steps (mr_upload_lake_forcing; basinEvapo_in / basinPrecip_in of main_route)."""
        e = np.ascontiguousarray(evapo, dtype=np.float64); p = np.ascontiguousarray(precip, dtype=np.float64)
        if e.ndim == 1:
            e, p = e[None, :], p[None, :]
        assert e.shape == p.shape and e.shape[1] == self.net.nHRU
        self._check(self._L.mr_upload_lake_forcing(self._h, int(e.shape[0]), _ptr(e, C.c_double), _ptr(p, C.c_double), self._msg))

    def upload_runoff(self, runoff):
        K, rp = self._host_ptr(runoff, self.nHRU)
        self._check(self._L.mr_upload_runoff(self._h, K, rp, self._msg))
        return K

    def route_resident(self, K: int):
        self._check(self._L.mr_route_resident(self._h, int(K), self.TSEC[0], self._msg))
        self._advance(K)

    def route_resident_async(self, K: int):
        """Enqueue K steps and return; `wait()` blocks and raises a device-side RoutingError, if any."""
        self._check(self._L.mr_route_resident_async(self._h, int(K), self.TSEC[0], self._msg))
        self._advance(K)

    def wait(self):
        self._check(self._L.mr_wait(self._h, self._msg))

    def download_q(self, K: int, out=None):
        if out is None:
            out = np.empty((len(self.methods), K, self.nRch))
        _, op = self._host_ptr(out, self.nRch, rows=len(self.methods) * K)
        self._check(self._L.mr_download_q(self._h, int(K), op, self._msg))
        return out

    def download_basin_q(self, K: int) -> np.ndarray:
        """BASIN_QR(1) (hillslope-routed lateral inflow, the reference's `dlayRunoff`) of the last batch, [K, nRch]."""
        out = np.empty((int(K), self.nRch))
        self._check(self._L.mr_download_basin_q(self._h, int(K), C.c_void_p(out.ctypes.data), self._msg))
        return out

    def history_means(self, K: int, n_agg: int, want_dlay: bool = False, flush: bool = False) -> np.ndarray:
        """Period means of REACH_Q per routing method (and of BASIN_QR(1) with want_dlay) over groups of n_agg steps, formed on
        the device from the first K steps of the last batch and rounded to float32 (mr_history_means; histVars_data.f90:154-246).
        A period may span calls; flush also closes the one still open.  Returns [nPeriods, n_routes (+1), nRch] float32."""
        n_series = len(self.methods) + (1 if want_dlay else 0)
        cap = int(K) // int(n_agg) + 2
        out = np.empty((cap, n_series, self.nRch), dtype=np.float32)
        n = C.c_int(0)
        self._check(self._L.mr_history_means(self._h, int(K), int(n_agg), int(bool(want_dlay)), int(bool(flush)), cap, C.c_void_p(out.ctypes.data),
                                             C.byref(n), self._msg))
        return out[:n.value]

    @staticmethod
    def _host_ptr(a, ncol: int, rows: Optional[int] = None):
        if hasattr(a, "data_ptr"):                               # torch tensor (pinned host memory)
            if a.dtype.itemsize != 8 or not a.is_contiguous() or a.device.type != "cpu":
                raise ValueError("expected a contiguous float64 host tensor")
            n = a.numel()
            ptr = C.c_void_p(a.data_ptr())
        else:
            if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous):
                raise ValueError("expected a C-contiguous float64 array")
            n = a.size
            ptr = C.c_void_p(a.ctypes.data)
        if n % ncol != 0:
            raise ValueError("array size is not a multiple of the reach/HRU count")
        k = n // ncol
        if rows is not None and k != rows:
            raise ValueError("output array has the wrong number of rows")
        return k, ptr

    # ------------------------------------------------------------------------------------------
    def flux(self, field: int, method: int = 1) -> np.ndarray:
        out = np.empty(self.nRch)
        self._check(self._L.mr_get_flux(self._h, int(method), int(field), _ptr(out, C.c_double), self._msg))
        return out

    def info(self, key: int) -> int:
        return int(self._L.mr_get_info(self._h, int(key)))

    def timing(self) -> dict:
        ms = (C.c_double * 8)()
        self._L.mr_get_timing(self._h, ms)
        out = {"total": ms[0], "basin": ms[1], "route_network": ms[2], "h2d": ms[3], "d2h": ms[4]}
        for i, m in enumerate(self.methods):
            out["route_%s" % ("sum", "irf", "kwt")[m]] = ms[5 + i]
        return out

    def set_stream(self, cuda_stream: int):
        """Launch on the caller's CUDA stream (e.g. torch.cuda.current_stream().cuda_stream); 0 = own stream."""
        self._check(self._L.mr_set_stream(self._h, C.c_void_p(cuda_stream or None), self._msg))

    def set_remap(self, n_forcing: int, map_hru_index, num_qhru, qhru_index, weight):
        """Runoff arrives on `n_forcing` polygons and is remapped to the network HRUs on the device (mr_set_remap,
        remap_1D_runoff of the reference); afterwards every runoff array is [K, n_forcing]."""
        a = np.ascontiguousarray(map_hru_index, dtype=np.int32); b = np.ascontiguousarray(num_qhru, dtype=np.int32)
        c = np.ascontiguousarray(qhru_index, dtype=np.int32); w = np.ascontiguousarray(weight, dtype=np.float64)
        self._check(self._L.mr_set_remap(self._h, int(n_forcing), len(a), _ptr(a, C.c_int), _ptr(b, C.c_int), _ptr(c, C.c_int),
                                         _ptr(w, C.c_double), self._msg))
        self.nHRU = int(n_forcing)                           # columns of the runoff arrays from now on

    # ---- multi-domain hand-off -------------------------------------------------------------------
    def set_export(self, seg_ids):
        ids = np.ascontiguousarray(seg_ids, dtype=np.int32)
        self._check(self._L.mr_set_export(self._h, len(ids), _ptr(ids, C.c_int), self._msg))

    def exchange_bytes(self, which: int) -> int:
        return int(self._L.mr_exchange_bytes(self._h, int(which)))

    def set_exchange_buffer(self, which: int, dev_ptr: int = 0, nbytes: int = 0):
        """which: 0 export / 1 import; dev_ptr 0 lets the library allocate."""
        self._check(self._L.mr_set_exchange_buffer(self._h, int(which), C.c_void_p(dev_ptr or None), int(nbytes), self._msg))

    def copy_exchange_to(self, dst: "Router", src_slot0: int, dst_slot0: int, n_slots: int):
        self._check(self._L.mr_copy_exchange(self._h, dst._h, int(src_slot0), int(dst_slot0), int(n_slots), self._msg))

    def set_counting(self, on: bool):
        self._check(self._L.mr_set_counting(self._h, int(on), self._msg))

    def basin_uh(self) -> np.ndarray:
        out = np.empty(self.info(capi.INFO_NTDH_BAS))
        self._check(self._L.mr_get_basin_uh(self._h, _ptr(out, C.c_double), self._msg))
        return out

    def reach_uh(self):
        ntdh = np.empty(self.nRch, dtype=np.int32)
        uh = np.empty((self.nRch, self.info(capi.INFO_MAXTDH)))
        self._check(self._L.mr_get_reach_uh(self._h, _ptr(ntdh, C.c_int), _ptr(uh, C.c_double), self._msg))
        return ntdh, uh

    def _state_shape(self, var: int):
        n, w = self.nRch, capi.MR_KW_SLOTS
        return {
            capi.ST_BASIN_QFUTURE: ((n, self.info(capi.INFO_NTDH_BAS)), np.float64),
            capi.ST_BASIN_QR: ((n, 2), np.float64),
            capi.ST_IRF_QFUTURE: ((n, self.info(capi.INFO_MAXTDH)), np.float64),
            capi.ST_IRF_VOL: ((n,), np.float64),
            capi.ST_KWT_NWAVE: ((n,), np.int32),
            capi.ST_KWT_QWAVE: ((n, w), np.float64),
            capi.ST_KWT_TENTRY: ((n, w), np.float64),
            capi.ST_KWT_TEXIT: ((n, w), np.float64),
            capi.ST_KWT_ROUTED: ((n, w), np.int32),
            capi.ST_LAKE_VOL: ((len(self.methods), n), np.float64),
            capi.ST_MOLECULE_KW: ((n, capi.N_MOLECULE[3]), np.float64),
            capi.ST_MOLECULE_MC: ((n, capi.N_MOLECULE[4]), np.float64),
            capi.ST_MOLECULE_DW: ((n, capi.N_MOLECULE[5]), np.float64),
            capi.ST_QERROR: ((len(self.methods), n), np.float64),
            capi.ST_DA_QOBS: ((n,), np.float64),
            capi.ST_DA_QELAPSED: ((n,), np.int32),
        }[var]

    def get_state(self, var: int) -> np.ndarray:
        shape, dt = self._state_shape(var)
        a = np.empty(shape, dtype=dt)
        self._check(self._L.mr_get_state(self._h, int(var), C.c_void_p(a.ctypes.data), a.nbytes, self._msg))
        return a

    def set_state(self, var: int, a: np.ndarray):
        shape, dt = self._state_shape(var)
        a = np.ascontiguousarray(a, dtype=dt)
        if a.shape != shape:
            raise ValueError(f"state variable {var} expects shape {shape}")
        self._check(self._L.mr_set_state(self._h, int(var), C.c_void_p(a.ctypes.data), a.nbytes, self._msg))

    def set_steps_done(self, steps: int):
        self._check(self._L.mr_set_steps_done(self._h, int(steps), self._msg))
        self.TSEC = [float(steps) * float(self.opts.dt), float(steps + 1) * float(self.opts.dt)]
