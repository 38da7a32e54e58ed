"""Small all-methods case for compute-sanitizer (memcheck / racecheck) runs on the GPU box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.util import case, rel_err
from mizuroute_b200.route import Router
from oracle.oracle import Oracle

net, params, opts, ro = case("random", n=200, seed=21, dt=3600.0, route_opt="012", steps=16)
r = Router(net, params, opts, max_batch=8)
q = np.concatenate([r.route_batch(np.ascontiguousarray(ro[s:s + 8])) for s in (0, 8)], axis=1)
qo = Oracle(net, params, opts).run(ro)
print("rel err", [rel_err(q[i], qo[i]) for i in range(3)])

from tests.util import star_network
net, params, opts, ro = star_network(steps=12)
r = Router(net, params, opts, max_batch=6)
q = np.concatenate([r.route_batch(np.ascontiguousarray(ro[s:s + 6])) for s in (0, 6)], axis=1)
qo = Oracle(net, params, opts).run(ro)
print("star rel err", rel_err(q[0], qo[0]))
