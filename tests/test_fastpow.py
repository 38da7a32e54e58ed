"""mr_pow04_fast (mr_dev.h): x**p2, p2 = (ALFA-1)/ALFA in double precision, the celerity exponent of kinwav_rch
(kwt_route.f90:1296) -- the one pow() of the KWT inner loop.  The routine must be as good as the libm pow() the reference
calls: here its error is measured against extended precision, on the host build and on the device."""
import ctypes as C
import math

import numpy as np
import pytest

P2 = ((5.0 / 3.0) - 1.0) / (5.0 / 3.0)


def _samples(n=400_000, seed=1):
    rng = np.random.default_rng(seed)
    x = np.exp(rng.uniform(-30, 30, n) * np.log(10.0))
    return np.concatenate([x, [1.0, 32.0, 1e-77, 1e77, 2.0 ** -255, 2.0 ** 256, 0.5, 3.0e-9]])


def _ulp_error(x, y):
    exact = np.power(x.astype(np.longdouble), np.longdouble(P2))
    ulp = np.spacing(y)
    return np.abs((y.astype(np.longdouble) - exact) / ulp.astype(np.longdouble)).astype(np.float64)


def test_host_build_within_half_an_ulp_of_the_exact_power():
    from tests import emul
    L = emul.load_fastpow()
    x = _samples(); y = np.empty_like(x)
    L.fastpow_eval(C.c_int(x.size), x.ctypes.data_as(C.POINTER(C.c_double)), y.ctypes.data_as(C.POINTER(C.c_double)))
    err = _ulp_error(x, y)
    assert err.max() <= 0.51, err.max()
    libm = np.array([math.pow(v, P2) for v in x[:60_000]])         # glibc, what the reference's ** calls
    assert np.mean(y[:60_000] == libm) > 0.995         # ... and this routine agree to the last bit nearly always
    assert _ulp_error(x[:60_000], libm).max() <= 0.53  # (glibc's own error, for comparison)


def test_special_values_fall_back_to_pow():
    from tests import emul
    L = emul.load_fastpow()
    x = np.array([0.0, -1.0, np.inf, np.nan, 1e-300, 1e300, 5e-324]); y = np.empty_like(x)
    L.fastpow_eval(C.c_int(x.size), x.ctypes.data_as(C.POINTER(C.c_double)), y.ctypes.data_as(C.POINTER(C.c_double)))
    with np.errstate(invalid="ignore"):
        want = np.power(x, P2)
    assert np.array_equal(y, want, equal_nan=True)


@pytest.mark.gpu
def test_device_within_half_an_ulp_of_the_exact_power():
    from mizuroute_b200 import synth
    from mizuroute_b200.network import RouteOptions, RouteParams
    from mizuroute_b200.route import Router
    r = Router(synth.random_tree(20, seed=1), RouteParams(), RouteOptions(dt=3600.0, route_opt="2", runoffMin=1e-15))
    x = _samples(); y = np.empty_like(x)
    r._check(r._L.mr_selftest_pow04(r._h, int(x.size), x.ctypes.data_as(C.POINTER(C.c_double)), y.ctypes.data_as(C.POINTER(C.c_double)), r._msg))
    err = _ulp_error(x, y)
    assert err.max() <= 0.51, err.max()
    assert np.mean(y[:60_000] == np.array([math.pow(v, P2) for v in x[:60_000]])) > 0.995
