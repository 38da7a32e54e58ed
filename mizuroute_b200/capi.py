"""ctypes binding of libmizuroute_b200.so (include/mizuroute_b200.h).

This is the binding a Python host uses; a Fortran host would use the bind(C) interface shown in
INTEGRATION.md.  There is no CPU fallback: if the shared library is missing it is built with nvcc, and if
there is no CUDA device `mr_create` fails with ierr 90.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

MR_STRLEN = 256
MR_KW_SLOTS = 22

# flux fields
REACH_Q, REACH_VOL1, REACH_INFLOW, WB, BASIN_QI, BASIN_QR1, BASIN_QR0, REACH_VOL0, QERROR = range(9)
R_WIDTH, TOTAREA, BASAREA, R_SLOPE, NGOOD = 10, 11, 12, 13, 14
KW_PITCH = 24            # particle-row pitch of an exchange record
# state variables
(ST_BASIN_QFUTURE, ST_BASIN_QR, ST_IRF_QFUTURE, ST_IRF_VOL, ST_KWT_NWAVE, ST_KWT_QWAVE, ST_KWT_TENTRY,
 ST_KWT_TEXIT, ST_KWT_ROUTED, ST_LAKE_VOL, ST_MOLECULE_KW, ST_MOLECULE_MC, ST_MOLECULE_DW,
 ST_QERROR, ST_DA_QOBS, ST_DA_QELAPSED) = range(16)
N_MOLECULE = {3: 20, 4: 2, 5: 20}     # nodes of the Euler schemes' molecules (route methods 3 KW, 4 MC, 5 DW)
# info keys
(INFO_NRCH, INFO_NHRU, INFO_NSTAGE, INFO_NTDH_BAS, INFO_MAXTDH, INFO_LAUNCHES_LAST, INFO_STEPS_DONE,
 INFO_MAX_BATCH, INFO_MAX_NUPS, INFO_KWT_PARTICLES, INFO_DEVICE_BYTES, INFO_KWT_TOUCHED, INFO_NHEAD, INFO_SUM_NTDH,
 INFO_SUM_NUPS) = range(15)
INFO_NFORCING = 15

EXPORTS = [
    "mr_create", "mr_set_network", "mr_step", "mr_step_batch", "mr_upload_runoff", "mr_route_resident",
    "mr_download_q", "mr_download_basin_q", "mr_history_means", "mr_selftest_pow04", "mr_get_flux", "mr_get_state", "mr_set_state", "mr_set_steps_done", "mr_get_basin_uh",
    "mr_get_reach_uh", "mr_get_info", "mr_get_timing", "mr_destroy", "mr_set_stream", "mr_set_counting",
    "mr_route_resident_async", "mr_step_batch_async", "mr_wait", "mr_set_remap", "mr_set_ghosts", "mr_set_export", "mr_exchange_bytes", "mr_set_exchange_buffer", "mr_get_exchange_buffer", "mr_copy_exchange",
    "mr_upload_lake_forcing", "mr_set_lake_param", "mr_set_sim_start", "mr_upload_wm", "mr_set_da", "mr_upload_obs", "mr_set_ingest", "mr_ingest_records",
]


class mr_options(C.Structure):
    _fields_ = [
        ("dt", C.c_double),
        ("n_routes", C.c_int),
        ("route_methods", C.c_int * 8),
        ("doesBasinRoute", C.c_int),
        ("hw_drain_point", C.c_int),
        ("min_length_route", C.c_double),
        ("is_lake_sim", C.c_int),
        ("lakeRegulate", C.c_int),
        ("LakeInputOption", C.c_int),
        ("runoffMin", C.c_double),
        ("time_conv", C.c_double),
        ("length_conv", C.c_double),
        ("fshape", C.c_double),
        ("tscale", C.c_double),
        ("velo", C.c_double),
        ("diff", C.c_double),
        ("mann_n", C.c_double),
        ("wscale", C.c_double),
        ("device", C.c_int),
        ("max_batch", C.c_int),
        ("floodplain", C.c_int),
        ("dscale", C.c_double),
        ("floodplainSlope", C.c_double),
    ]


_LIB = None


def lib_path() -> str:
    return _build.LIB


def load(rebuild_if_stale: bool = True):
    """Load (building first if the in-tree .so is missing or older than its sources)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("MR_LIB_PATH", _build.LIB)         # development: load an alternative build of the library
    if path == _build.LIB and rebuild_if_stale and _build.stale():
        if os.path.exists("/usr/local/cuda/bin/nvcc") or os.environ.get("NVCC"):
            _build.build()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing and could not be built: the routing path has no CPU fallback")
    L = C.CDLL(path)
    dp, ip, vp, cp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p, C.c_char_p
    L.mr_create.argtypes = [C.POINTER(mr_options), C.POINTER(vp), cp]
    L.mr_set_network.argtypes = [vp, C.c_int, C.c_int, ip, ip, ip, dp, dp, dp, dp, dp, ip, ip, dp, dp, dp, dp, cp]
    L.mr_step.argtypes = [vp, C.c_double, C.c_double, dp, cp]
    L.mr_step_batch.argtypes = [vp, C.c_int, C.c_double, vp, vp, cp]
    L.mr_upload_runoff.argtypes = [vp, C.c_int, vp, cp]
    L.mr_route_resident.argtypes = [vp, C.c_int, C.c_double, cp]
    L.mr_download_q.argtypes = [vp, C.c_int, vp, cp]
    L.mr_download_basin_q.argtypes = [vp, C.c_int, vp, cp]
    L.mr_selftest_pow04.argtypes = [vp, C.c_int, dp, dp, cp]
    L.mr_history_means.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.POINTER(C.c_int), cp]
    L.mr_route_resident_async.argtypes = [vp, C.c_int, C.c_double, cp]
    L.mr_step_batch_async.argtypes = [vp, C.c_int, C.c_double, vp, vp, cp]
    L.mr_wait.argtypes = [vp, cp]
    L.mr_get_flux.argtypes = [vp, C.c_int, C.c_int, dp, cp]
    L.mr_get_state.argtypes = [vp, C.c_int, vp, C.c_long, cp]
    L.mr_set_state.argtypes = [vp, C.c_int, vp, C.c_long, cp]
    L.mr_set_steps_done.argtypes = [vp, C.c_long, cp]
    L.mr_get_basin_uh.argtypes = [vp, dp, cp]
    L.mr_get_reach_uh.argtypes = [vp, ip, dp, cp]
    L.mr_get_info.argtypes = [vp, C.c_int]
    L.mr_get_info.restype = C.c_long
    L.mr_get_timing.argtypes = [vp, dp]
    L.mr_set_stream.argtypes = [vp, vp, cp]
    L.mr_set_counting.argtypes = [vp, C.c_int, cp]
    L.mr_set_remap.argtypes = [vp, C.c_int, C.c_int, ip, ip, ip, dp, cp]
    L.mr_set_ghosts.argtypes = [vp, C.c_int, ip, ip, dp, dp, cp]
    L.mr_set_export.argtypes = [vp, C.c_int, ip, cp]
    L.mr_exchange_bytes.argtypes = [vp, C.c_int]
    L.mr_exchange_bytes.restype = C.c_long
    L.mr_set_exchange_buffer.argtypes = [vp, C.c_int, vp, C.c_long, cp]
    L.mr_get_exchange_buffer.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_long), cp]
    L.mr_copy_exchange.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, cp]
    L.mr_upload_lake_forcing.argtypes = [vp, C.c_int, dp, dp, cp]
    L.mr_upload_wm.argtypes = [vp, C.c_int, dp, dp, C.c_int, cp]
    L.mr_set_da.argtypes = [vp, C.c_int, C.c_int, C.c_int, cp]
    L.mr_upload_obs.argtypes = [vp, C.c_int, C.POINTER(C.c_int), dp, cp]
    L.mr_set_ingest.argtypes = [vp, C.c_int, C.POINTER(C.c_int), C.c_double, C.c_double, C.c_double, cp]
    L.mr_ingest_records.argtypes = [vp, C.c_int, C.c_int, dp, C.POINTER(C.c_int), C.POINTER(C.c_int), dp, cp]
    L.mr_set_lake_param.argtypes = [vp, cp, C.c_int, dp, cp]
    L.mr_set_sim_start.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, cp]
    L.mr_destroy.argtypes = [vp]
    L.mr_destroy.restype = None
    for name in EXPORTS:
        if name not in ("mr_get_info", "mr_destroy", "mr_exchange_bytes"):
            getattr(L, name).restype = C.c_int
    _LIB = L
    return L
