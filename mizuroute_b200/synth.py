"""Seeded synthetic river networks and runoff forcing (BASELINE.json configs).

The reference ships no data (its Cameo test case is an external download), so every workload is
generated here, reproducibly from a seed:

* ``random_tree``   -- C1 substitute: small random tree, mixed in-degree.
* ``binary_tree``   -- C2: complete binary tree truncated to n reaches (heap numbering).
* ``conus_like``    -- C3/C4/C5: forest of uniform-random binary trees (Shreve's random-topology model,
  height ~ 2*sqrt(pi*n)), heavy-tailed basin sizes with the largest basin ~40 % of the reaches, a few
  percent of confluences widened to in-degree 3/4, optional lakes.

All generators return reaches in shuffled order with non-trivial ids so the host topology code
(id->index join, upstream lists, levels) is exercised the way a real network file would.
"""
from __future__ import annotations

import numpy as np

from .network import RiverNetwork


# ----------------------------------------------------------------------------------------------
# topology generators: return parent[] (downstream index, -1 for an outlet), unshuffled
# ----------------------------------------------------------------------------------------------
def _uniform_binary_tree(n_internal: int, rng: np.random.Generator) -> np.ndarray:
    """Uniform random full binary tree with n_internal confluences (2*n_internal+1 reaches).

    Lukasiewicz word + cycle lemma, fully vectorised: nodes are in preorder; node i is a first
    child iff node i-1 is internal, otherwise its parent is the previous node at the same walk height.
    """
    n = 2 * n_internal + 1
    if n_internal == 0:
        return np.array([-1], dtype=np.int64)
    steps = np.empty(n, dtype=np.int64)
    steps[:n_internal] = 1
    steps[n_internal:] = -1
    rng.shuffle(steps)
    w = np.cumsum(steps)
    k = int(np.argmin(w))                       # first position of the minimum
    steps = np.roll(steps, -(k + 1))            # valid preorder degree word (deg-1)
    W = np.concatenate(([0], np.cumsum(steps)[:-1]))   # W_i = height before visiting node i
    parent = np.full(n, -1, dtype=np.int64)
    idx = np.arange(n)
    first_child = np.zeros(n, dtype=bool)
    first_child[1:] = steps[:-1] == 1
    parent[first_child] = idx[first_child] - 1
    order = np.lexsort((idx, W))
    prev = np.full(n, -1, dtype=np.int64)
    same = W[order[1:]] == W[order[:-1]]
    prev[order[1:][same]] = order[:-1][same]
    second = ~first_child
    second[0] = False
    parent[second] = prev[second]
    return parent


def _contract(parent: np.ndarray, frac: float, rng: np.random.Generator) -> np.ndarray:
    """Remove a random fraction of the confluences; their tributaries join the next confluence
    downstream (creates in-degree 3, 4, ...)."""
    n = parent.shape[0]
    if frac <= 0 or n < 8:
        return parent
    nchild = np.bincount(parent[parent >= 0], minlength=n)
    cand = (nchild > 0) & (parent >= 0)
    removed = cand & (rng.random(n) < frac)
    anc = parent.copy()
    while True:
        hit = (anc >= 0) & removed[np.maximum(anc, 0)]
        if not hit.any():
            break
        anc[hit] = parent[anc[hit]]
    keep = ~removed
    new_index = np.cumsum(keep) - 1
    out = anc[keep]
    out = np.where(out >= 0, new_index[np.maximum(out, 0)], -1)
    return out


def _forest(sizes, frac_wide: float, rng: np.random.Generator) -> np.ndarray:
    parts, off = [], 0
    for s in sizes:
        p = _uniform_binary_tree(max((int(s) - 1) // 2, 0), rng)
        p = _contract(p, frac_wide, rng)
        parts.append(np.where(p >= 0, p + off, -1))
        off += p.shape[0]
    return np.concatenate(parts)


# ----------------------------------------------------------------------------------------------
# attributes
# ----------------------------------------------------------------------------------------------
def _finish(parent: np.ndarray, rng: np.random.Generator, shuffle: bool, meta: dict,
            zero_area_frac: float = 0.0) -> RiverNetwork:
    n = parent.shape[0]
    perm = rng.permutation(n) if shuffle else np.arange(n)      # new position of old node i
    inv = np.empty(n, dtype=np.int64)
    inv[perm] = np.arange(n)                                    # old node at new position p
    segId = (np.arange(n) * 7 + 1001).astype(np.int64)          # id by new position (unique, >0, unsorted vs topology)
    segId = segId[rng.permutation(n)] if shuffle else segId
    down_old = parent[inv]                                      # downstream (old numbering) of the node at position p
    downSegId = np.where(down_old >= 0, segId[perm[np.maximum(down_old, 0)]], -1)
    length = np.clip(np.exp(rng.normal(np.log(2000.0), 0.6, n)), 100.0, 50000.0)
    slope = np.exp(rng.uniform(np.log(1e-4), np.log(1e-1), n))
    area = np.exp(rng.normal(np.log(5.0e6), 0.8, n))
    # one HRU per reach, HRUs stored in their own shuffled order; optionally some reaches get no HRU
    # at all (BASAREA = 0; exercises the goodBas / runoffMin paths, process_remap.f90:413-415)
    hperm = rng.permutation(n) if shuffle else np.arange(n)
    if zero_area_frac > 0:
        hperm = hperm[rng.random(n) >= zero_area_frac]
    hruSegId = segId[hperm]
    hruId = (hruSegId + 50_000_000).astype(np.int64)
    return RiverNetwork(segId=segId, downSegId=downSegId, length=length, slope=slope,
                        hruId=hruId, hruSegId=hruSegId, area=area[hperm], meta=meta)


def random_tree(n: int = 50, seed: int = 1, window: int = 6, shuffle: bool = True,
                zero_area_frac: float = 0.0) -> RiverNetwork:
    """Small random tree/forest: reach i drains to a random reach among the `window` before it."""
    rng = np.random.default_rng(seed)
    parent = np.full(n, -1, dtype=np.int64)
    for i in range(1, n):
        if rng.random() < 0.03:
            continue                                            # another outlet
        parent[i] = rng.integers(max(0, i - window), i)
    return _finish(parent, rng, shuffle, {"kind": "random_tree", "n": n, "seed": seed}, zero_area_frac)


def binary_tree(n: int = 100_000, seed: int = 2, shuffle: bool = True) -> RiverNetwork:
    """Complete binary tree in heap numbering truncated to n reaches (C2: 100 000 reaches, 17 levels)."""
    rng = np.random.default_rng(seed)
    k = np.arange(1, n + 1)
    parent = (k // 2) - 1
    parent[0] = -1
    return _finish(parent.astype(np.int64), rng, shuffle, {"kind": "binary_tree", "n": n, "seed": seed})


def conus_like(n: int = 3_000_000, seed: int = 3, largest_frac: float = 0.4, frac_wide: float = 0.04,
               shuffle: bool = True, n_lakes: int = 0) -> RiverNetwork:
    """CONUS-like forest (C3/C4/C5).  Basin sizes: one basin of largest_frac*n reaches, the rest
    Pareto(alpha=0.9)-distributed between 3 and 8 % of n until n reaches are used up."""
    rng = np.random.default_rng(seed)
    n_gen = int(round(n / (1.0 - 0.49 * frac_wide)))            # contraction removes ~frac_wide/2 of the reaches
    sizes = [int(largest_frac * n_gen)]
    left = n_gen - sizes[0]
    cap = max(int(0.08 * n_gen), 3)
    while left > 0:
        s = int(min(cap, 3.0 * (1.0 + rng.pareto(0.9))))
        s = min(s, left)
        if s % 2 == 0:
            s = max(s - 1, 1)
        sizes.append(s)
        left -= s
    parent = _forest(sizes, frac_wide, rng)
    meta = {"kind": "conus_like", "n_target": n, "seed": seed, "n_basins": len(sizes), "largest_basin": sizes[0]}
    net = _finish(parent, rng, shuffle, meta)
    if n_lakes > 0:
        add_lakes(net, n_lakes, rng)
    return net


def _down_index(net: RiverNetwork) -> np.ndarray:
    n = net.nRch
    order = np.argsort(net.segId, kind="stable")
    sid = net.segId[order]
    pos = np.clip(np.searchsorted(sid, net.downSegId), 0, n - 1)
    hit = (net.downSegId > 0) & (sid[pos] == net.downSegId)
    return np.where(hit, order[pos], -1).astype(np.int64)


def add_lakes(net: RiverNetwork, n_lakes: int, rng: np.random.Generator, frac_endorheic: float = 0.2) -> None:
    """C5: turn mid-network reaches into lakes.  80 % Doll-2003, 20 % endorheic.  The reach below a lake must
    have the lake as its only upstream (kwt_route.f90:551), so for every picked reach c a new "lake outlet"
    reach is spliced in between c and its downstream reach (appended at the end, with its own HRU).  Lakes are
    kept at least one reach apart."""
    n = net.nRch
    down = _down_index(net)
    nup = np.bincount(down[down >= 0], minlength=n)
    cand = np.flatnonzero((nup > 0) & (down >= 0))
    rng.shuffle(cand)
    src = np.flatnonzero(down >= 0)
    o = np.argsort(down[src], kind="stable")
    up_idx = src[o]
    up_ptr = np.concatenate(([0], np.cumsum(nup)))
    blocked = np.zeros(n, dtype=bool)
    picked = []
    for c in cand:
        if len(picked) >= n_lakes:
            break
        if blocked[c]:
            continue
        picked.append(int(c))
        blocked[c] = True
        blocked[down[c]] = True
        blocked[up_idx[up_ptr[c]:up_ptr[c + 1]]] = True
    picked = np.array(picked, dtype=np.int64)
    endo = picked[rng.random(picked.size) < frac_endorheic]      # endorheic lakes are terminal: nothing flows out
    doll = np.setdiff1d(picked, endo)
    m = doll.size
    new_id = (int(net.segId.max()) + 1 + np.arange(m)).astype(np.int32)
    old_down_id = net.downSegId[doll].copy()
    net.downSegId[doll] = new_id                                 # lake -> its outlet reach
    net.downSegId[endo] = -1
    net.segId = np.concatenate([net.segId, new_id])
    net.downSegId = np.concatenate([net.downSegId, old_down_id]).astype(np.int32)
    net.length = np.concatenate([net.length, np.clip(np.exp(rng.normal(np.log(2000.0), 0.6, m)), 100.0, 50000.0)])
    net.slope = np.concatenate([net.slope, np.exp(rng.uniform(np.log(1e-4), np.log(1e-1), m))])
    net.hruSegId = np.concatenate([net.hruSegId, new_id])
    net.hruId = np.concatenate([net.hruId, (new_id.astype(np.int64) + 50_000_000).astype(np.int32)])
    net.area = np.concatenate([net.area, np.exp(rng.normal(np.log(5.0e6), 0.8, m))])
    nn = n + m
    islake = np.zeros(nn, dtype=np.int32)
    islake[picked] = 1
    ltype = np.ones(nn, dtype=np.int32)
    ltype[endo] = 0
    net.islake = islake
    net.lakeModelType = ltype
    net.D03_MaxStorage = np.where(islake == 1, np.exp(rng.normal(np.log(5.0e7), 1.0, nn)), 0.0)
    net.D03_Coefficient = np.where(islake == 1, rng.uniform(0.005, 0.05, nn), 0.0)
    net.D03_Power = np.where(islake == 1, 1.5, 0.0)
    net.D03_S0 = np.zeros(nn)
    net.meta["n_lakes"] = int(islake.sum())


def make_hype_lakes(net: RiverNetwork, rng: np.random.Generator, frac: float = 0.5) -> int:
    """Turn a share of the Doll-2003 lakes into HYPE reservoirs (lakeModelType 3, lake_route.f90:398-438) with plausible
    parameters: ~10 m deep at the emergency spillway, a seasonal primary spillway, both outflow-combination modes."""
    doll = np.flatnonzero((net.islake == 1) & (net.lakeModelType == 1))
    pick = doll[rng.random(doll.size) < frac]
    if pick.size == 0 and doll.size:
        pick = doll[:1]
    n = net.nRch
    net.lakeModelType = net.lakeModelType.copy()
    net.lakeModelType[pick] = 3
    area = np.exp(rng.normal(np.log(2.0e7), 0.7, n))                 # m2
    e_zero = rng.uniform(100.0, 900.0, n)
    net.lake_params = {
        "HYP_A_avg": area, "HYP_E_zero": e_zero, "HYP_E_min": e_zero + rng.uniform(0.5, 2.0, n),
        "HYP_E_lim": e_zero + rng.uniform(4.0, 6.0, n), "HYP_E_emr": e_zero + rng.uniform(8.0, 12.0, n),
        "HYP_Qrate_emr": rng.uniform(20.0, 80.0, n), "HYP_Erate_emr": rng.uniform(1.0, 2.0, n),
        "HYP_Qrate_prim": rng.uniform(2.0, 30.0, n), "HYP_Qrate_amp": rng.uniform(0.0, 1.2, n),
        "HYP_Qrate_phs": rng.integers(0, 365, n).astype(np.float64), "HYP_prim_F": (rng.random(n) < 0.8).astype(np.float64),
        "HYP_Qsim_mode": (rng.random(n) < 0.5).astype(np.float64),
    }
    return int(pick.size)


def make_h06_lakes(net: RiverNetwork, rng: np.random.Generator, frac: float = 0.5, memory: bool = True, mem_years: int = 1) -> int:
    """Turn a share of the Doll-2003 lakes into Hanasaki-2006 reservoirs (lakeModelType 2, lake_route.f90:231-396):
    irrigation and non-irrigation purposes, within-a-year and multi-year storage ratios, optional inflow memory."""
    doll = np.flatnonzero((net.islake == 1) & (net.lakeModelType == 1))
    pick = doll[rng.random(doll.size) < frac]
    if pick.size == 0 and doll.size:
        pick = doll[:1]
    n = net.nRch
    net.lakeModelType = net.lakeModelType.copy()
    net.lakeModelType[pick] = 2
    base = np.exp(rng.normal(np.log(8.0), 0.8, n))                   # mean inflow m3/s
    season = 1.0 + 0.5 * np.sin(2.0 * np.pi * (np.arange(12)[:, None] + rng.uniform(0, 12, n)[None, :]) / 12.0)
    p = dict(net.lake_params)
    months = ["Jan", "Feb", "Mar", "Apr", "May", "Jun", "Jul", "Aug", "Sep", "Oct", "Nov", "Dec"]
    for k, mo in enumerate(months):
        p["H06_I_" + mo] = base * season[k]
        p["H06_D_" + mo] = base * rng.uniform(0.1, 0.9, n) * season[(k + 5) % 12]
    ratio = np.where(rng.random(n) < 0.5, rng.uniform(0.05, 0.4, n), rng.uniform(0.6, 2.0, n))      # c below / above c_compare = 0.5
    p["H06_Smax"] = ratio * base * 365.0 * 86400.0
    p.update({"H06_alpha": np.full(n, 0.85), "H06_envfact": rng.uniform(0.2, 0.9, n), "H06_S_ini": p["H06_Smax"] * 0.8,
              "H06_c1": np.full(n, 0.1), "H06_c2": np.full(n, 0.9), "H06_exponent": np.full(n, 2.0), "H06_denominator": np.full(n, 0.5),
              "H06_c_compare": np.full(n, 0.5), "H06_frac_Sdead": np.full(n, 0.1), "H06_E_rel_ini": rng.uniform(0.6, 1.1, n),
              "H06_purpose": (rng.random(n) < 0.5).astype(np.float64), "H06_I_mem_F": np.full(n, 1.0 if memory else 0.0),
              "H06_D_mem_F": np.zeros(n), "H06_I_mem_L": np.full(n, float(mem_years)), "H06_D_mem_L": np.full(n, float(mem_years))})
    net.lake_params = p
    return int(pick.size)


def runoff_series(net: RiverNetwork, n_steps: int, seed: int = 11, dt: float = 86400.0,
                  mean_mm_s: float = 2.0e-5, sigma: float = 1.0) -> np.ndarray:
    """Strictly positive runoff depth [n_steps, nHRU] in mm/s: per-HRU lognormal level x
    seasonal factor x step-to-step lognormal noise (KWT aborts on zero flow, kwt_route.f90:1365)."""
    rng = np.random.default_rng(seed)
    base = np.exp(rng.normal(np.log(mean_mm_s), sigma, net.nHRU))
    t = np.arange(n_steps) * dt
    season = 1.0 + 0.6 * np.sin(2.0 * np.pi * t / (365.0 * 86400.0))
    storm = np.exp(rng.normal(0.0, 0.5, n_steps))
    out = np.empty((n_steps, net.nHRU), dtype=np.float64)
    for k in range(n_steps):
        out[k] = base * season[k] * storm[k] * np.exp(rng.normal(0.0, 0.3, net.nHRU))
    return out


def network_stats(net: RiverNetwork) -> dict:
    """Realised topology statistics (reported next to the assumed NHDPlus-like targets)."""
    n = net.nRch
    down = _down_index(net)
    nup = np.bincount(down[down >= 0], minlength=n)
    # distance to outlet by pointer doubling
    d = (down >= 0).astype(np.int64)
    a = down.copy()
    while (a >= 0).any():
        has = a >= 0
        ai = np.maximum(a, 0)
        d = d + np.where(has, d[ai], 0)
        a = np.where(has, a[ai], -1)
    conf = nup[nup >= 2]
    return {
        "nRch": n, "n_outlets": int((down < 0).sum()), "headwater_frac": float((nup == 0).mean()),
        "max_indegree": int(nup.max()),
        "conf_deg2": float((conf == 2).mean()) if conf.size else 0.0,
        "conf_deg3": float((conf == 3).mean()) if conf.size else 0.0,
        "conf_deg4p": float((conf >= 4).mean()) if conf.size else 0.0,
        "max_depth": int(d.max()) + 1,
    }
