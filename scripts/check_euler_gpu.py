"""Torch-free GPU check of the Euler routing schemes (route_opt 3/4/5): parity against the oracle on the cases of
tests/test_schemes_gpu.py, the molecule state round trip, and a first throughput figure.  Written to fit a very short
GPU slot: prints one JSON line per item (also into gpurun_out/euler_check.jsonl)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mizuroute_b200 import capi  # noqa: E402
from mizuroute_b200.route import Router  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402
from tests.util import case, rel_err  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
LOG = open(os.path.join(OUT, "euler_check.jsonl"), "a")


def emit(**kw):
    line = json.dumps(kw)
    print(line, flush=True)
    LOG.write(line + "\n"); LOG.flush()


def both(net, params, opts, ro, batch):
    o = Oracle(net, params, opts)
    qo = o.run(ro)
    r = Router(net, params, opts, max_batch=max(batch, 8))
    qg = np.concatenate([r.route_batch(np.ascontiguousarray(ro[s:s + batch])) for s in range(0, ro.shape[0], batch)], axis=1)
    return o, r, qo, qg


def main():
    t_start = time.time()
    cases = [
        ("hourly", dict(kind="random", n=80, seed=5, dt=3600.0, steps=40, zero_area_frac=0.08), 1),
        ("daily-substeps", dict(kind="random", n=60, seed=6, dt=86400.0, steps=20), 7),
        ("conus3000", dict(kind="conus", n=3000, seed=4, dt=3600.0, steps=24), 24),
        ("hw-top", dict(kind="random", n=50, seed=7, dt=900.0, steps=30, hw_drain_point=1), 8),
        ("floodplain", dict(kind="binary", n=1023, seed=2, dt=10800.0, steps=20, floodplain=True), 5),
        ("pass-through", dict(kind="random", n=40, seed=8, dt=3600.0, steps=12, min_length_route=1500.0), 4),
        ("lakes", dict(kind="conus", n=1500, seed=4, dt=86400.0, steps=12, lakes=12), 6),
    ]
    for name, kw, batch in cases:
        try:
            net, params, opts, ro = case(route_opt="345", **kw)
            o, r, qo, qg = both(net, params, opts, ro, batch)
            errs = {}
            for i, m in enumerate((3, 4, 5)):
                errs["q%d" % m] = rel_err(qg[i], qo[i])
                errs["vol%d" % m] = rel_err(r.flux(capi.REACH_VOL1, m), o.get(orc.F_REACH_VOL1, m), floor=1e-6)
                errs["mol%d" % m] = rel_err(r.get_state(capi.ST_MOLECULE_KW + (m - 3)), o.molecule(m), floor=1e-12)
            emit(item="parity", case=name, **errs)
        except Exception as e:          # keep going: every line is evidence
            emit(item="parity", case=name, error=repr(e)[:300])
    try:
        net, params, opts, ro = case("conus", n=6000, seed=7, dt=3600.0, route_opt="012345", steps=20)
        o, r, qo, qg = both(net, params, opts, ro, 10)
        emit(item="six_methods", bit_identical_sum_irf=bool(np.array_equal(qg[:2], qo[:2])), errs=[rel_err(qg[i], qo[i]) for i in range(6)])
    except Exception as e:
        emit(item="six_methods", error=repr(e)[:300])
    try:
        net, params, opts, ro = case("conus", n=800, seed=3, dt=3600.0, route_opt="345", steps=24)
        a = Router(net, params, opts, max_batch=12); a.route_batch(np.ascontiguousarray(ro[:12]))
        b = Router(net, params, opts, max_batch=12); b.set_steps_done(12)
        for v in (capi.ST_BASIN_QFUTURE, capi.ST_BASIN_QR, capi.ST_MOLECULE_KW, capi.ST_MOLECULE_MC, capi.ST_MOLECULE_DW, capi.ST_LAKE_VOL):
            b.set_state(v, a.get_state(v))
        b.TSEC = list(a.TSEC)
        emit(item="state_round_trip", exact=bool(np.array_equal(a.route_batch(np.ascontiguousarray(ro[12:])), b.route_batch(np.ascontiguousarray(ro[12:])))))
    except Exception as e:
        emit(item="state_round_trip", error=repr(e)[:300])
    # first throughput figure (device time of the routing kernels from mr_get_timing, forcing resident)
    try:
        n, K = int(os.environ.get("EULER_N", "200000")), 48
        net, params, opts, ro = case("conus", n=n, seed=2, dt=3600.0, route_opt="345", steps=K)
        r = Router(net, params, opts, max_batch=K)
        r.route_batch(ro)                       # warm-up (also spins the state up)
        t0 = time.time(); r.route_batch(ro); wall = time.time() - t0
        ms = (C.c_double * 8)()
        r._L.mr_get_timing(r._h, ms)
        emit(item="throughput", n_reach=net.nRch, steps=K, wall_ms=wall * 1e3, timing_ms=list(ms),
             reach_steps_per_s_wall=net.nRch * K / wall, note="timing_ms[5..7] = device ms of methods 3, 4, 5 (they run concurrently)")
    except Exception as e:
        emit(item="throughput", error=repr(e)[:300])
    emit(item="done", seconds=time.time() - t_start)


if __name__ == "__main__":
    import ctypes as C
    main()
