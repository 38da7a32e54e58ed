"""Development fuzz: the host build of the Euler routing schemes (tests/emul) against the oracle on random networks and options;
every case must agree bit for bit.  CPU only."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.test_euler_emul import _run
from oracle import oracle as orc
bad = 0; n = 0
t0 = time.time()
rng = np.random.default_rng(123)
for it in range(60):
    kind = ["random", "binary", "conus"][it % 3]
    kw = dict(kind=kind, n=int(rng.integers(20, 400)), seed=int(rng.integers(1, 10000)), dt=float(rng.choice([900.0, 3600.0, 10800.0, 86400.0])),
              steps=int(rng.integers(5, 30)))
    if kind == "random": kw["zero_area_frac"] = float(rng.choice([0.0, 0.1, 0.3]))
    if rng.random() < 0.3: kw["hw_drain_point"] = 1
    if rng.random() < 0.3: kw["min_length_route"] = float(rng.choice([500.0, 2000.0]))
    if rng.random() < 0.3: kw["floodplain"] = True
    for m in (3, 4, 5):
        try:
            o, qo, qe, ve, me = _run(kw, m)
            ok = np.array_equal(qe, qo) and np.array_equal(ve, o.get(orc.F_REACH_VOL1, m)) and np.array_equal(me, o.molecule(m)) and np.isfinite(qo).all()
        except Exception as e:
            ok = False; print("EXC", kw, m, repr(e)[:200])
        n += 1
        if not ok:
            bad += 1; print("MISMATCH", kw, m)
print("cases", n, "bad", bad, "%.0fs" % (time.time() - t0))
