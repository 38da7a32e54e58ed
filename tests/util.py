"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np

from mizuroute_b200 import synth
from mizuroute_b200.network import RouteOptions, RouteParams

IRF_RTOL = 1.0e-6     # north_star tolerance, IRF (double precision)
KWT_RTOL = 1.0e-4     # north_star tolerance, KWT


def rel_err(a, b, floor=1e-300):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0


def case(kind="random", n=60, seed=5, dt=3600.0, route_opt="012", steps=24, zero_area_frac=0.0, lakes=0, **kw):
    if kind.startswith("tiny:"):                                  # the degenerate networks of tiny_networks()
        net, params, opts, ro = tiny_case(tiny_networks()[kind[5:]], dt=dt, route_opt=route_opt, steps=steps, seed=seed)
        for k, v in kw.items():
            setattr(opts, k, v)
        return net, params, opts, ro
    if kind == "random":
        net = synth.random_tree(n, seed=seed, zero_area_frac=zero_area_frac)
    elif kind == "binary":
        net = synth.binary_tree(n, seed=seed)
    else:
        net = synth.conus_like(n, seed=seed, n_lakes=0)
    opts = RouteOptions(dt=dt, route_opt=route_opt, runoffMin=1e-15, **kw)
    if lakes:
        synth.add_lakes(net, lakes, np.random.default_rng(seed + 100))
        opts.is_lake_sim = True
        opts.LakeInputOption = 1
    ro = synth.runoff_series(net, steps, seed=seed + 1, dt=dt)
    return net, RouteParams(), opts, ro


def star_network(arms=12, depth=3, seed=3, steps=40, dt=3600.0, route_opt="2"):
    """`arms` chains of `depth` interior reaches (each fed by two headwaters) draining into one hub reach: a
    confluence wide enough to overflow the shared-memory-sized KWT scratch (full-capacity retry path)."""
    from mizuroute_b200.network import RiverNetwork
    rng = np.random.default_rng(seed)
    seg, down = [1], [0]
    sid = 2
    for _ in range(arms):
        prev = 1
        for _k in range(depth):
            me = sid; sid += 1
            seg.append(me); down.append(prev)
            for _h in range(2):
                seg.append(sid); down.append(me); sid += 1
            prev = me
    n = len(seg)
    net = RiverNetwork(segId=np.array(seg), downSegId=np.array(down), length=rng.uniform(500, 3000, n), slope=rng.uniform(1e-3, 1e-2, n),
                       hruId=np.arange(1, n + 1), hruSegId=np.array(seg), area=rng.uniform(1e6, 2e7, n))
    opts = RouteOptions(dt=dt, route_opt=route_opt, runoffMin=1e-15)
    ro = np.abs(rng.lognormal(np.log(2e-5), 1.0, size=(steps, n))) + 1e-9
    return net, RouteParams(), opts, ro


def gauge_series(net, K, seed=3, n_gauge=40, record_frac=0.5, scale=(0.3, 2.5), base=None):
    """Synthetic gauge file for the data-assimilation tests: obs [K][nRch] (NaN = no gauge / missing, a few negative),
    has_record [K] int32 (0 = no record at that step).  base [K][nRch]: flows the values are scattered around."""
    rng = np.random.default_rng(seed)
    gauges = rng.choice(net.nRch, min(n_gauge, net.nRch), replace=False)
    has = (rng.random(K) < record_frac).astype(np.int32)
    has[0] = 0; has[min(2, K - 1)] = 1
    obs = np.full((K, net.nRch), np.nan)
    for t in range(K):
        b = base[t, gauges] if base is not None else rng.lognormal(0.0, 1.0, gauges.size)
        v = b * rng.uniform(scale[0], scale[1], gauges.size)
        v[rng.random(gauges.size) < 0.15] = np.nan          # gauge silent at that record
        v[rng.random(gauges.size) < 0.05] = -1.0            # flagged bad
        obs[t, gauges] = v
    return obs, has, gauges


def tiny_networks():
    """Degenerate river networks by hand: one reach; isolated reaches only (every reach an outlet and a headwater); a chain of
    two; a reach without any HRU in a chain of three; a five-arm star."""
    from mizuroute_b200.network import RiverNetwork

    def net(down, hru_seg, seed):
        rng = np.random.default_rng(seed)
        n = len(down)
        return RiverNetwork(segId=np.arange(101, 101 + n), downSegId=np.array([101 + d if d >= 0 else -1 for d in down]),
                            length=rng.uniform(800.0, 6000.0, n), slope=rng.uniform(1e-3, 2e-2, n),
                            hruId=np.arange(1001, 1001 + len(hru_seg)), hruSegId=np.array([101 + s for s in hru_seg]),
                            area=rng.uniform(2e6, 3e7, len(hru_seg)))
    return {
        "one_reach": net([-1], [0], 1),
        "isolated_reaches": net([-1, -1, -1, -1], [0, 1, 2, 3], 2),
        "chain_of_two": net([1, -1], [0, 1], 3),
        "middle_reach_without_hru": net([1, 2, -1], [0, 2], 4),
        "star_of_five": net([5, 5, 5, 5, 5, -1], [0, 1, 2, 3, 4, 5], 5),
    }


def tiny_case(net, dt=3600.0, route_opt="012345", steps=30, seed=0):
    rng = np.random.default_rng(100 + seed)
    opts = RouteOptions(dt=dt, route_opt=route_opt, runoffMin=1e-15)
    season = 1.0 + 0.5 * np.sin(np.arange(steps) / 5.0)
    ro = rng.lognormal(np.log(2e-5), 1.0, (steps, net.nHRU)) * season[:, None]
    return net, RouteParams(), opts, ro
