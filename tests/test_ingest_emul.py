"""Forcing ingest as the GPU runs it (mizuroute_b200/csrc/mr_ingest.h: time-weighted mean over the forcing records under a
step, scale / offset, forcing HRU -> river-network HRU, missing and negative -> 0), compiled for the host, against the rows
the stand-alone host builds on the CPU for the same files (`route_runoff --dry-run --dump-forcing`), BIT FOR BIT."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from mizuroute_b200 import build as mrbuild
from mizuroute_b200 import casefiles
from tests import emul
from tests.util import case
from tests.util_ingest import time_map


def _emul_rows(records, src, ptr, idx, frac, scale=-9999.0, offset=-9999.0, fill=-9999.0):
    L = emul.load_ingest()
    K, nH = ptr.size - 1, src.size
    out = np.empty((K, nH))
    p = lambda a, ct: None if a is None else np.ascontiguousarray(a).ctypes.data_as(C.POINTER(ct))
    rescale = int(scale != -9999.0 or offset != -9999.0)
    L.ingest_emul_run(C.c_int(nH), C.c_int(records.shape[1]), C.c_int(K), p(records, C.c_double), p(src.astype(np.int32), C.c_int), p(ptr, C.c_int),
                      p(idx, C.c_int), p(frac, C.c_double), C.c_int(rescale), C.c_double(1.0 if scale == -9999.0 else scale),
                      C.c_double(0.0 if offset == -9999.0 else offset), C.c_double(fill), p(out, C.c_double))
    return out


def _host_rows(ctl, tmp_path, cols):
    path = os.path.join(str(tmp_path), "forcing.f64")
    r = subprocess.run([mrbuild.build_host(), ctl, "--dry-run", "--dump-forcing", path], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return np.fromfile(path, dtype=np.float64).reshape(-1, cols)


@pytest.mark.parametrize("dt,forcing_dt,records,sim_steps", [(3600.0, 3600.0, 12, 12), (3600.0, 10800.0, 10, 30), (10800.0, 3600.0, 36, 12),
                                                             (7200.0, 10800.0, 8, 12)])
def test_ingest_equals_the_rows_of_the_host(tmp_path, dt, forcing_dt, records, sim_steps):
    net, params, opts, ro = case("random", n=60, seed=4, dt=dt, route_opt="1", steps=records)
    ro = ro.copy(); ro[1, 3] = -2.0; ro[2, 5] = -9999.0; ro[3:6, 7] = -9999.0         # a negative value, fill values
    ctl = casefiles.write_case(str(tmp_path), net, params, opts, ro, case_name="rows", forcing_dt=forcing_dt, sim_steps=sim_steps,
                               extra_keys={"scale_factor_runoff": "1.5", "offset_value_runoff": "1.e-9"})
    want = _host_rows(ctl, tmp_path, net.nHRU)
    ptr, idx, frac = time_map(sim_steps, dt, records, forcing_dt)
    got = _emul_rows(ro, np.arange(net.nHRU), ptr, idx, frac, scale=1.5, offset=1e-9)
    assert np.array_equal(got, want)
    if forcing_dt >= dt and forcing_dt % dt == 0:                                   # one record per step: the share array may be left out
        assert np.array_equal(_emul_rows(ro, np.arange(net.nHRU), ptr, idx, None, scale=1.5, offset=1e-9), want)


def test_ingest_sorts_forcing_columns_into_network_order():
    """sort_flux: shuffled forcing columns, river-network HRUs the forcing does not hold get 0."""
    rng = np.random.default_rng(3)
    nH, nIn, K = 40, 55, 6
    rec = rng.lognormal(0.0, 1.0, (K, nIn)); rec[2, 10] = -1.0
    src = rng.permutation(nIn)[:nH].astype(np.int32); src[[4, 9]] = -1
    ptr, idx, frac = time_map(K, 3600.0, K, 3600.0)
    got = _emul_rows(rec, src, ptr, idx, None)
    want = np.where(src[None, :] >= 0, np.maximum(rec[:, np.maximum(src, 0)], 0.0), 0.0)
    assert np.array_equal(got, want)
