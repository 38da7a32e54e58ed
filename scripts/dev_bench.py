"""Developer timing probe (not the contract bench): routes a synthetic network and prints device timings."""
import argparse
import json
import sys
import time
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from mizuroute_b200 import capi, synth
from mizuroute_b200.network import RouteOptions, RouteParams
from mizuroute_b200.route import Router

ap = argparse.ArgumentParser()
ap.add_argument("--kind", default="conus")
ap.add_argument("--n", type=int, default=300000)
ap.add_argument("--dt", type=float, default=86400.0)
ap.add_argument("--route", default="2")
ap.add_argument("--K", type=int, default=32)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--spin", type=int, default=1)
a = ap.parse_args()

t0 = time.time()
net = synth.binary_tree(a.n, seed=2) if a.kind == "binary" else synth.conus_like(a.n, seed=3)
ro = synth.runoff_series(net, a.K, seed=11, dt=a.dt)
t1 = time.time()
opts = RouteOptions(dt=a.dt, route_opt=a.route, runoffMin=1e-15)
r = Router(net, RouteParams(), opts, max_batch=a.K)
t2 = time.time()
info = {k: r.info(v) for k, v in [("nStage", capi.INFO_NSTAGE), ("ntdh_bas", capi.INFO_NTDH_BAS), ("maxtdh", capi.INFO_MAXTDH),
                                  ("maxUps", capi.INFO_MAX_NUPS), ("devKiB", capi.INFO_DEVICE_BYTES)]}
print(json.dumps({"nRch": net.nRch, "gen_s": round(t1 - t0, 2), "init_s": round(t2 - t1, 2), **info}))
r.upload_runoff(ro)
for i in range(a.spin + a.reps):
    w0 = time.time()
    r.route_resident(a.K)
    w1 = time.time()
    tm = r.timing()
    units = net.nRch * a.K
    print(json.dumps({"rep": i, "wall_ms": round((w1 - w0) * 1e3, 2), **{k: round(v, 3) for k, v in tm.items()},
                      "launches": r.info(capi.INFO_LAUNCHES_LAST), "Mrs_per_s": round(units / tm["total"] / 1e3, 1),
                      "particles_per_reach": round(r.info(capi.INFO_KWT_PARTICLES) / net.nRch, 2) if "2" in a.route else None}))
w0 = time.time()
q = r.route_batch(ro)
w1 = time.time()
print(json.dumps({"e2e_wall_ms": round((w1 - w0) * 1e3, 2), **{k: round(v, 3) for k, v in r.timing().items()}}))
