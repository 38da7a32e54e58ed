"""ctypes front-end of oracle/libmr_oracle.so (CPU restatement of the reference routing path).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
Parity unpinned (see mr_oracle.c header).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

M_SUM, M_IRF, M_KWT, M_KW, M_MC, M_DW = range(6)        # = the digits of <route_opt> (public_var.f90:74-80)
METHOD_OF_DIGIT = {str(m): m for m in range(6)}
N_MOLECULE = {M_KW: 20, M_MC: 2, M_DW: 20}               # init_model_data.f90:386-393
F_REACH_Q, F_REACH_VOL1, F_REACH_INFLOW, F_WB, F_BASIN_QI, F_BASIN_QR1, F_BASIN_QR0, F_REACH_VOL0 = range(8)
F_QERROR, F_QOBS = 8, 9
F_WIDTH, F_TOTAREA, F_BASAREA, F_SLOPE = 10, 11, 12, 13
KW_CAP = 24


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libmr_oracle.so")
    src = os.path.join(_HERE, "mr_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.mro_create.restype = C.c_void_p
        _LIB.mro_message.restype = C.c_char_p
        _LIB.mro_gammp.restype = C.c_double
        _LIB.mro_gammp.argtypes = [C.c_double, C.c_double]
    return _LIB


def _p(a, ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


class OracleError(RuntimeError):
    def __init__(self, ierr, msg):
        super().__init__(f"ierr={ierr}: {msg}")
        self.ierr = ierr


class Oracle:
    """One routing domain on the CPU: create -> step()/run() -> get()."""

    def __init__(self, net, params, opts, n_threads: int = 1):
        L = lib()
        self.net, self.params, self.opts = net, params, opts
        tc, lc = opts.conv()
        self.methods = [METHOD_OF_DIGIT[c] for c in opts.route_opt]
        dbl = lambda v: C.c_double(float(v))
        self.h = C.c_void_p(L.mro_create(
            C.c_int(net.nRch), C.c_int(net.nHRU),
            _p(net.segId, C.c_int), _p(net.downSegId, C.c_int), _p(net.hruSegId, C.c_int), _p(net.area, C.c_double),
            _p(net.length, C.c_double), _p(net.slope, C.c_double), _p(net.width, C.c_double), _p(net.man_n, C.c_double),
            _p(net.islake, C.c_int), _p(net.lakeModelType, C.c_int),
            _p(net.D03_MaxStorage, C.c_double), _p(net.D03_Coefficient, C.c_double),
            _p(net.D03_Power, C.c_double), _p(net.D03_S0, C.c_double),
            dbl(opts.dt), C.c_char_p(opts.route_opt.encode()),
            C.c_int(opts.doesBasinRoute), C.c_int(opts.hw_drain_point), dbl(opts.min_length_route),
            C.c_int(int(opts.is_lake_sim)), C.c_int(int(opts.lakeRegulate)), C.c_int(opts.LakeInputOption),
            dbl(opts.runoffMin), dbl(tc), dbl(lc),
            dbl(params.fshape), dbl(params.tscale), dbl(params.velo), dbl(params.diff), dbl(params.mann_n), dbl(params.wscale),
            C.c_int(n_threads)))
        if not self.h:
            raise OracleError(-1, "mro_create failed (bad route_opt or UH construction)")
        for name, vals in (getattr(net, "lake_params", None) or {}).items():
            v = np.ascontiguousarray(vals, dtype=np.float64)
            assert v.shape == (net.nRch,)
            if L.mro_set_lake_param(self.h, C.c_char_p(name.encode()), _p(v, C.c_double)) != 0:
                raise OracleError(20, "unknown lake parameter " + name)
        if getattr(opts, "sim_start", None):
            y, mo, d, sec = opts.sim_start
            L.mro_set_sim_start(self.h, C.c_int(y), C.c_int(mo), C.c_int(d), dbl(sec), C.c_int(int(opts.calendar == "noleap")))
        if getattr(opts, "floodplain", False):          # <floodplain> T: bankfull depth dscale*sqrt(totalArea) (process_ntopo.f90:174-203)
            L.mro_set_channel(self.h, C.c_int(1), dbl(4.5000000682193786e-05), dbl(1000.0))
        self.T0, self.T1 = 0.0, float(opts.dt)      # init_model_data.f90:600

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().mro_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def _check(self, ierr):
        if ierr != 0:
            raise OracleError(ierr, lib().mro_message(self.h).decode())

    def step(self, runoff: np.ndarray, evapo=None, precip=None):
        """evapo / precip [nHRU] (runoff units): lake evaporation / precipitation forcing (main_route.f90:174-199)."""
        r = np.ascontiguousarray(runoff, dtype=np.float64)
        assert r.shape == (self.net.nHRU,)
        e = None if evapo is None else np.ascontiguousarray(evapo, dtype=np.float64)
        p = None if precip is None else np.ascontiguousarray(precip, dtype=np.float64)
        self._check(lib().mro_step_ep(self.h, C.c_double(self.T0), C.c_double(self.T1), _p(r, C.c_double), _p(e, C.c_double), _p(p, C.c_double)))
        self.T0 = self.T1
        self.T1 = self.T0 + float(self.opts.dt)     # init_model_data.f90:311-312

    def run(self, runoff: np.ndarray, want_q: bool = True, evapo=None, precip=None):
        r = np.ascontiguousarray(runoff, dtype=np.float64)
        n = r.shape[0]
        q = np.empty((len(self.methods), n, self.net.nRch)) if want_q else None
        e = None if evapo is None else np.ascontiguousarray(evapo, dtype=np.float64)
        p = None if precip is None else np.ascontiguousarray(precip, dtype=np.float64)
        self._check(lib().mro_run_ep(self.h, C.c_int(n), C.c_double(self.T0), _p(r, C.c_double), _p(e, C.c_double), _p(p, C.c_double), _p(q, C.c_double)))
        for _ in range(n):
            self.T0 = self.T1
            self.T1 = self.T0 + float(self.opts.dt)
        return q

    def set_threads(self, n: int):
        lib().mro_set_threads(self.h, C.c_int(int(n)))

    def get(self, field: int, method: int = M_IRF) -> np.ndarray:
        out = np.empty(self.net.nRch)
        assert lib().mro_get(self.h, C.c_int(method), C.c_int(field), _p(out, C.c_double)) == 0
        return out

    def set(self, field: int, values, method: int = M_IRF):
        v = np.ascontiguousarray(values, dtype=np.float64)
        assert lib().mro_set(self.h, C.c_int(method), C.c_int(field), _p(v, C.c_double)) == 0

    def set_wm(self, flux_wm=None, vol_wm=None, vol_jumpstart: bool = False):
        """Water management of the following steps (is_flux_wm / is_vol_wm): flux_wm [nRch] abstraction (+) / injection (-),
        -9999 = none at that reach; vol_wm [nRch] target volume of the lakes flagged by the lake parameter LakeTargVol."""
        self._wm = (None if flux_wm is None else np.ascontiguousarray(flux_wm, dtype=np.float64),
                    None if vol_wm is None else np.ascontiguousarray(vol_wm, dtype=np.float64))     # kept alive: the C side holds the pointers
        lib().mro_set_wm(self.h, _p(self._wm[0], C.c_double), _p(self._wm[1], C.c_double), C.c_int(int(vol_jumpstart)))

    def set_da(self, qmod_option: int = 1, q_blend_period: int = 10, q_err_trend: int = 1):
        """Data assimilation by direct insertion (<qmodOption> 1, <qBlendPeriod>, <QerrTrend> 1 constant / 2 linear /
        3 logistic / 4 exponential; public_var.f90:189-191)."""
        lib().mro_set_da(self.h, C.c_int(int(qmod_option)), C.c_int(int(q_blend_period)), C.c_int(int(q_err_trend)))

    def set_obs(self, obs=None):
        """Gauge observations [m3/s] of the NEXT step only: obs [nRch], NaN or negative = no value at that reach; None = the
        gauge file has no record at that time (every reach's Qelapsed goes up by one)."""
        self._obs = None if obs is None else np.ascontiguousarray(obs, dtype=np.float64)   # kept alive until the step
        lib().mro_set_obs(self.h, _p(self._obs, C.c_double))

    def qelapsed(self) -> np.ndarray:
        out = np.empty(self.net.nRch, dtype=np.int32)
        lib().mro_get_qelapsed(self.h, _p(out, C.c_int))
        return out

    def lake_forcing(self):
        """(reach evaporation, reach precipitation) [m3/s] of the last step; the evaporation is what lake_route left
        (cut to the lake volume where the lake ran dry)."""
        e, p = np.empty(self.net.nRch), np.empty(self.net.nRch)
        lib().mro_get_lake_forcing(self.h, _p(e, C.c_double), _p(p, C.c_double))
        return e, p

    def molecule(self, method: int) -> np.ndarray:
        """molecule%Q of an Euler scheme, [nRch, N_MOLECULE[method]]."""
        out = np.empty((self.net.nRch, N_MOLECULE[method]))
        lib().mro_get_molecule(self.h, C.c_int(method), _p(out, C.c_double))
        return out

    def set_molecule(self, method: int, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        assert v.shape == (self.net.nRch, N_MOLECULE[method])
        lib().mro_set_molecule(self.h, C.c_int(method), _p(v, C.c_double))

    # --- unit hydrographs -------------------------------------------------------------------
    def frac_future(self) -> np.ndarray:
        out = np.empty(lib().mro_ntdh_bas(self.h))
        lib().mro_get_frac_future(self.h, _p(out, C.c_double))
        return out

    def reach_uh(self):
        ptr = np.empty(self.net.nRch + 1, dtype=np.int32)
        lib().mro_get_uh_ptr(self.h, _p(ptr, C.c_int))
        val = np.empty(int(ptr[-1]))
        lib().mro_get_uh_val(self.h, _p(val, C.c_double))
        return ptr, val

    def down_index(self) -> np.ndarray:
        out = np.empty(self.net.nRch, dtype=np.int32)
        lib().mro_get_down_index(self.h, _p(out, C.c_int))
        return out

    def n_level(self) -> int:
        return int(lib().mro_nlevel(self.h))

    # --- state (restart layout) ---------------------------------------------------------------
    def get_state(self) -> dict:
        L, n = lib(), self.net.nRch
        st = {}
        nb = L.mro_ntdh_bas(self.h)
        qf = np.empty((n, nb)); L.mro_get_qfuture(self.h, _p(qf, C.c_double)); st["qfuture"] = qf
        st["basin_qr1"] = self.get(F_BASIN_QR1); st["basin_qr0"] = self.get(F_BASIN_QR0)
        if M_IRF in self.methods:
            ptr, _ = self.reach_uh()
            q = np.empty(int(ptr[-1])); L.mro_get_qfuture_irf(self.h, _p(q, C.c_double))
            st["irf_qfuture"] = q; st["irf_vol"] = self.get(F_REACH_VOL1, M_IRF)
        if M_KWT in self.methods:
            nw = np.empty(n, dtype=np.int32)
            a = [np.empty((n, KW_CAP)) for _ in range(3)]
            rf = np.empty((n, KW_CAP), dtype=np.uint8)
            L.mro_get_kwt_state(self.h, C.c_int(KW_CAP), _p(nw, C.c_int), _p(a[0], C.c_double), _p(a[1], C.c_double),
                                _p(a[2], C.c_double), _p(rf, C.c_ubyte))
            st.update(kwt_n=nw, kwt_qf=a[0], kwt_ti=a[1], kwt_tr=a[2], kwt_rf=rf)
        return st


def seed_oracle_from_router(o: "Oracle", r) -> None:
    """Copy the complete routing state of a mizuroute_b200 Router (restart schema, mr_get_state) into an Oracle,
    so both continue from the same spun-up state (used for full-size parity samples and the CPU baseline)."""
    from mizuroute_b200 import capi
    L, n = lib(), o.net.nRch
    steps = r.info(capi.INFO_STEPS_DONE)
    qf = np.ascontiguousarray(r.get_state(capi.ST_BASIN_QFUTURE))
    L.mro_set_qfuture(o.h, _p(qf, C.c_double))
    qr = r.get_state(capi.ST_BASIN_QR)
    o.set(F_BASIN_QR0, np.ascontiguousarray(qr[:, 0])); o.set(F_BASIN_QR1, np.ascontiguousarray(qr[:, 1]))
    if M_IRF in o.methods:
        ptr, _ = o.reach_uh()
        irf = r.get_state(capi.ST_IRF_QFUTURE)
        ntdh = np.diff(ptr)
        mask = np.arange(irf.shape[1])[None, :] < ntdh[:, None]
        flat = np.ascontiguousarray(irf[mask])
        assert flat.size == int(ptr[-1])
        L.mro_set_qfuture_irf(o.h, _p(flat, C.c_double))
        o.set(F_REACH_VOL1, r.get_state(capi.ST_IRF_VOL), M_IRF)
    if M_KWT in o.methods:
        nw = np.ascontiguousarray(r.get_state(capi.ST_KWT_NWAVE), dtype=np.int32)
        qw = np.ascontiguousarray(r.get_state(capi.ST_KWT_QWAVE)); ti = np.ascontiguousarray(r.get_state(capi.ST_KWT_TENTRY))
        tr = np.ascontiguousarray(r.get_state(capi.ST_KWT_TEXIT))
        rf = np.ascontiguousarray(r.get_state(capi.ST_KWT_ROUTED).astype(np.uint8))
        L.mro_set_kwt_state(o.h, C.c_int(qw.shape[1]), _p(nw, C.c_int), _p(qw, C.c_double), _p(ti, C.c_double), _p(tr, C.c_double), _p(rf, C.c_ubyte))
    if o.opts.is_lake_sim:
        lv = r.get_state(capi.ST_LAKE_VOL)
        for i, m in enumerate(o.methods):
            if m != M_SUM:
                o.set(F_REACH_VOL1, np.ascontiguousarray(lv[i]), m)
    L.mro_set_itime(o.h, C.c_long(steps + 1))
    o.T0, o.T1 = r.TSEC[0], r.TSEC[1]


def gammp(a, x):
    return lib().mro_gammp(a, x)


def make_uh_one(length, dt, velo, diff):
    out = np.empty(256)
    n = lib().mro_make_uh_one(C.c_double(length), C.c_double(dt), C.c_double(velo), C.c_double(diff), _p(out, C.c_double))
    return out[:n].copy()


def interp_rch(T, Q, T0, T1):
    T = np.ascontiguousarray(T, dtype=np.float64); Q = np.ascontiguousarray(Q, dtype=np.float64)
    out = C.c_double(0.0)
    ierr = lib().mro_interp_rch(_p(T, C.c_double), _p(Q, C.c_double), C.c_int(T.size), C.c_double(T0), C.c_double(T1), C.byref(out))
    return ierr, out.value


def remove_rch(Q, T, X):
    Q = np.array(Q, dtype=np.float64); T = np.array(T, dtype=np.float64); X = np.array(X, dtype=np.float64)
    n = lib().mro_remove_rch(_p(Q, C.c_double), _p(T, C.c_double), _p(X, C.c_double), C.c_int(Q.size))
    return Q[:n], T[:n], X[:n]


def remap_1d(hru_ix, num_qhru, qhru_ix, weight, sim, n_hru):
    """remap_1D_runoff for one time step (process_remap.f90:164-262): forcing vector `sim` -> basinRunoff[n_hru]."""
    a = np.ascontiguousarray(hru_ix, dtype=np.int32); b = np.ascontiguousarray(num_qhru, dtype=np.int32)
    c = np.ascontiguousarray(qhru_ix, dtype=np.int32); w = np.ascontiguousarray(weight, dtype=np.float64)
    s = np.ascontiguousarray(sim, dtype=np.float64)
    out = np.zeros(n_hru)
    lib().mro_remap_1d(C.c_int(len(a)), _p(a, C.c_int), _p(b, C.c_int), _p(c, C.c_int), _p(w, C.c_double), _p(s, C.c_double), _p(out, C.c_double))
    return out
