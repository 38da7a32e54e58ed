"""mr_calendar.h -- month, day and day of year of simDatetime(1) for every step of a batch, which the HYPE and Hanasaki
reservoirs read -- against Python's datetime (standard = proleptic Gregorian calendar) and plain 365-day arithmetic
(noleap), over random start dates, step lengths and step counts, across leap days, century years and negative offsets."""
import ctypes as C
import datetime as dt

import numpy as np

from tests import emul


def test_step_calendar_matches_python_datetime():
    L = emul.load_calendar()
    L.calendar_emul.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double, C.c_longlong,
                                C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    rng = np.random.default_rng(1)
    ml = [31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31]
    mo, dy, doy = C.c_int(), C.c_int(), C.c_int()
    for _ in range(4000):
        y = int(rng.choice([1899, 1900, 1968, 1999, 2000, 2001, 2004, 2023, 2024, 2100, 2399, 2400]))
        m = int(rng.integers(1, 13)); d = int(rng.integers(1, 29))
        sec = float(rng.choice([0.0, 3600.0, 43200.0, 86399.0]))
        step_s = float(rng.choice([900.0, 3600.0, 10800.0, 43200.0, 86400.0]))
        k = int(rng.integers(0, 200000)) if step_s < 86400.0 else int(rng.integers(0, 40000))
        # standard calendar
        L.calendar_emul(y, m, d, sec, 0, step_s, k, C.byref(mo), C.byref(dy), C.byref(doy))
        now = dt.datetime(y, m, d) + dt.timedelta(seconds=sec + k * step_s)
        assert (mo.value, dy.value, doy.value) == (now.month, now.day, now.timetuple().tm_yday), (y, m, d, sec, step_s, k)
        # noleap calendar
        L.calendar_emul(y, m, d, sec, 1, step_s, k, C.byref(mo), C.byref(dy), C.byref(doy))
        days = sum(ml[:m - 1]) + (d - 1) + int((sec + k * step_s) // 86400.0)
        yd = days % 365
        mm = 0
        rest = yd
        while rest >= ml[mm]:
            rest -= ml[mm]; mm += 1
        assert (mo.value, dy.value, doy.value) == (mm + 1, rest + 1, yd + 1), ("noleap", y, m, d, sec, step_s, k)
