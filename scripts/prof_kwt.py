"""Development profile driver: routes warm-up batches of a CONUS-like network, then ONE batch between cudaProfilerStart/Stop, so
that `ncu --profile-from-start off` only sees (and only slows down) that batch.
usage: python scripts/prof_kwt.py [nRch] [tsteps] [warm batches] [route_opt]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from mizuroute_b200 import synth  # noqa: E402
from mizuroute_b200.network import RouteOptions, RouteParams  # noqa: E402
from mizuroute_b200.route import Router  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 64
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 2
route = sys.argv[4] if len(sys.argv) > 4 else "2"
net = synth.conus_like(n, seed=3)
opts = RouteOptions(dt=3600.0, route_opt=route, runoffMin=1e-15)
ro = synth.runoff_series(net, T, seed=11, dt=3600.0)
r = Router(net, RouteParams(), opts, device=0, max_batch=T)
r.upload_runoff(ro)
for _ in range(warm):
    r.route_resident(T)
torch.cuda.synchronize()
rt = torch.cuda.cudart()
rt.cudaProfilerStart()
r.route_resident(T)
torch.cuda.synchronize()
rt.cudaProfilerStop()
print("timing", r.timing())
