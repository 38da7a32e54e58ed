"""Writes the text fixture ref_replay.f90 reads, from one of the cases of tests/golden/make_golden.py (lake-free, route_opt 0/1/2):

    python oracle/ref_build/export_fixture.py <case name> <fixture.txt>   ->   also returns (q [route, step, reach in processing order])

Reaches go out in a processing order (upstream before downstream) with what `augment_ntopo` derives and the routines read:
upstream reach lists with goodBas (network_topo.f90:769-775: every upstream flag of a reach = "its total area > verySmall"),
basin / upstream / total area, width = wscale * sqrt(total area) (process_ntopo.f90:176-187), slope floor min_slope.  The per-step
input is the reach-level instantaneous runoff BASIN_QI (the output of basin2reach), taken from the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def export(name: str, path: str):
    from oracle import oracle as orc
    from oracle.oracle import Oracle
    from tests.golden.make_golden import CASES
    from tests.util import case
    spec = CASES[name]
    if spec.get("lakes"):
        raise SystemExit("export_fixture: lake cases are not replayed")
    net, params, opts, ro = case(**spec)
    n = net.nRch
    order = np.argsort(net.segId, kind="stable")
    pos = np.clip(np.searchsorted(net.segId[order], net.downSegId), 0, n - 1)
    hit = (net.downSegId > 0) & (net.segId[order][pos] == net.downSegId)
    down = np.where(hit, order[pos], -1)
    # processing order: by distance to the outlet, farthest first
    depth = np.zeros(n, dtype=np.int64)
    for i in range(n):
        d, j = 0, i
        while down[j] >= 0:
            j = down[j]; d += 1
        depth[i] = d
    proc = np.argsort(-depth, kind="stable")
    rank = np.empty(n, dtype=np.int64); rank[proc] = np.arange(n)
    bas = np.zeros(n)
    seg_of_hru = {int(s): i for i, s in enumerate(net.segId)}
    for a, s in zip(net.area, net.hruSegId):
        bas[seg_of_hru[int(s)]] += a
    tot = bas.copy()
    for i in proc:                                          # upstream first: push the total area downstream
        if down[i] >= 0:
            tot[down[i]] += tot[i]
    ups = [[] for _ in range(n)]
    for i in range(n):
        if down[i] >= 0:
            ups[down[i]].append(i)
    o = Oracle(net, params, opts)
    qi = np.empty((ro.shape[0], n))
    for t in range(ro.shape[0]):
        o.step(ro[t]); qi[t] = o.get(orc.F_BASIN_QI)
    methods = [int(c) for c in opts.route_opt]
    with open(path, "w") as f:
        f.write(f"{n} {net.nHRU} {ro.shape[0]} {len(methods)}\n")
        f.write(" ".join(repr(float(v)) for v in (opts.dt, params.fshape, params.tscale, params.velo, params.diff, params.mann_n, params.wscale)))
        f.write(f" {int(opts.hw_drain_point)} {float(opts.min_length_route)!r}\n")
        f.write(" ".join(str(m) for m in methods) + "\n")
        for i in proc:
            u = sorted(ups[i], key=lambda j: rank[j])
            width = params.wscale * np.sqrt(tot[i])
            f.write(f"{int(net.segId[i])} {int(rank[down[i]]) + 1 if down[i] >= 0 else 0} {float(net.length[i])!r} {max(float(net.slope[i]), 1e-6)!r} "
                    f"{float(bas[i])!r} {float(tot[i] - bas[i])!r} {float(tot[i])!r} {float(width)!r} {len(u)}\n")
            for j in u:
                f.write(f"{int(rank[j]) + 1} {1 if tot[i] > 1e-12 else 0}\n")      # verySmall = 1e-12, public_var.f90
        for t in range(ro.shape[0]):
            f.write(" ".join(repr(float(v)) for v in qi[t][proc]) + "\n")
    q = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))["q"]
    return q[:, :, proc], methods


if __name__ == "__main__":
    export(sys.argv[1], sys.argv[2])
