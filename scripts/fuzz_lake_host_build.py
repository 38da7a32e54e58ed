"""Development fuzz: the host build of the lake code (forcing, HYPE, Hanasaki; tests/emul) against the oracle with random networks,
calendars, batch sizes; only equality failures count (the canned tests also assert coverage conditions).  CPU only."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tests.test_lake_emul as T
import tests.util as U
orig = U.case
bad = 0; n = 0; t0 = time.time()
rng = np.random.default_rng(7)
for it in range(40):
    seed = int(rng.integers(1, 10000)); nn = int(rng.integers(200, 900)); lakes = int(rng.integers(4, 14))
    def patched(*a, **k):
        k = dict(k); k["seed"] = seed; k["n"] = nn; k["lakes"] = lakes
        return orig(*a, **k)
    T.case = patched
    combos = [(int(rng.integers(0, 3)), bool(rng.integers(0, 2)), None),
              (int(rng.integers(0, 3)), bool(rng.integers(0, 2)), ("standard", (int(rng.integers(1999, 2005)), int(rng.integers(1, 13)), int(rng.integers(1, 28)), 0.0))),
              (int(rng.integers(0, 3)), bool(rng.integers(0, 2)), ("noleap", (2001, int(rng.integers(1, 13)), int(rng.integers(1, 28)), 43200.0)))]
    for opt, forcing, hype in combos:
        n += 1
        try:
            T.test_lake_reach_device_source_matches_oracle(opt, forcing, hype)
        except AssertionError as e:
            import traceback
            line = traceback.extract_tb(e.__traceback__)[-1].line
            if "array_equal" in line or "ierr" in line:
                bad += 1; print("MISMATCH", seed, nn, lakes, opt, forcing, hype, line)
        except Exception as e:
            bad += 1; print("EXC", seed, nn, lakes, opt, forcing, hype, repr(e)[:200])
    for memory, cal, start, dt, steps, K in [(bool(rng.integers(0, 2)), "standard", (2000, int(rng.integers(1, 13)), int(rng.integers(1, 28)), 0.0), 86400.0, int(rng.integers(10, 40)), int(rng.integers(1, 12))),
                                             (True, "noleap", (2001, int(rng.integers(1, 13)), int(rng.integers(1, 28)), 0.0), 43200.0, int(rng.integers(10, 40)), int(rng.integers(1, 12)))]:
        n += 1
        try:
            T.test_hanasaki_reservoirs_two_methods_in_device_order(memory, cal, start, dt, steps, K)
        except AssertionError as e:
            import traceback
            line = traceback.extract_tb(e.__traceback__)[-1].line
            if "array_equal" in line or "ierr" in line:
                bad += 1; print("MISMATCH H06", seed, nn, lakes, memory, cal, start, dt, steps, K, line)
        except Exception as e:
            bad += 1; print("EXC H06", seed, nn, lakes, memory, cal, start, dt, steps, K, repr(e)[:200])
print("cases", n, "bad", bad, "%.0fs" % (time.time() - t0))
