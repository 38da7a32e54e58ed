"""Spatial decomposition of a river network across GPUs.

The reference splits the network into independent upstream subtrees ("tributary" domains, bin-packed
largest-first onto ranks, domain_decomposition.f90:724-819) plus a serial "mainstem" on rank 0
(:508-519,595-719).  `partition_basins` is the first half of that rule: whole river basins (trees of the
forest) are independent units and are bin-packed largest-first onto the ranks, with no data-path exchange.
"""
from __future__ import annotations

import numpy as np

from .network import RiverNetwork
from .synth import _down_index


def basin_of(net: RiverNetwork) -> np.ndarray:
    """Index of the outlet reach each reach drains to (pointer jumping, O(N log depth))."""
    down = _down_index(net)
    root = np.where(down >= 0, down, np.arange(net.nRch))
    while True:
        nxt = root[root]
        if np.array_equal(nxt, root):
            return root
        root = nxt


def partition_basins(net: RiverNetwork, nparts: int):
    """Largest-first bin packing of whole basins (domain_decomposition.f90:791-809 uses the same greedy rule
    for tributary domains).  Returns a list of sorted reach-index arrays."""
    root = basin_of(net)
    outlets, inv, counts = np.unique(root, return_inverse=True, return_counts=True)
    order = np.argsort(-counts, kind="stable")
    load = np.zeros(nparts, dtype=np.int64)
    owner = np.empty(outlets.size, dtype=np.int64)
    for b in order:
        k = int(np.argmin(load))
        owner[b] = k
        load[k] += counts[b]
    part_of_reach = owner[inv]
    return [np.flatnonzero(part_of_reach == k) for k in range(nparts)]


def subnetwork(net: RiverNetwork, reaches: np.ndarray) -> RiverNetwork:
    """The reaches listed (closed under 'upstream of') and the HRUs that drain into them."""
    keep = np.zeros(net.nRch, dtype=bool)
    keep[reaches] = True
    order = np.argsort(net.segId, kind="stable")
    sid = net.segId[order]
    pos = np.clip(np.searchsorted(sid, net.hruSegId), 0, net.nRch - 1)
    hit = (net.hruSegId > 0) & (sid[pos] == net.hruSegId)
    hru_keep = hit & keep[order[pos]]
    pick = lambda a: None if a is None else a[reaches]
    sub = RiverNetwork(segId=net.segId[reaches], downSegId=net.downSegId[reaches], length=net.length[reaches],
                       slope=net.slope[reaches], hruId=net.hruId[hru_keep], hruSegId=net.hruSegId[hru_keep],
                       area=net.area[hru_keep], width=pick(net.width), man_n=pick(net.man_n), islake=pick(net.islake),
                       lakeModelType=pick(net.lakeModelType), D03_MaxStorage=pick(net.D03_MaxStorage),
                       D03_Coefficient=pick(net.D03_Coefficient), D03_Power=pick(net.D03_Power), D03_S0=pick(net.D03_S0),
                       lake_params={k: np.asarray(v)[reaches] for k, v in (net.lake_params or {}).items()},
                       meta=dict(net.meta))
    sub.meta["hru_index"] = np.flatnonzero(hru_keep)      # columns of the global runoff array this domain reads
    sub.meta["reach_index"] = np.asarray(reaches)
    return sub


# ------------------------------------------------------------------------------------------------
# tributary / mainstem decomposition (domain_decomposition.f90:450-819)
# ------------------------------------------------------------------------------------------------
def upstream_size(net: RiverNetwork) -> np.ndarray:
    """size(allUpSegIndices): number of reaches upstream of each reach INCLUDING itself (O(N), level by level)."""
    down = _down_index(net)
    n = net.nRch
    hops = np.where(down >= 0, 1, 0).astype(np.int64)
    ptr = np.where(down >= 0, down, np.arange(n))
    while True:                                   # pointer doubling: hops to the outlet
        nh, npt = hops + hops[ptr], ptr[ptr]
        if np.array_equal(npt, ptr):
            break
        hops, ptr = nh, npt
    size = np.ones(n, dtype=np.int64)
    order = np.argsort(-hops, kind="stable")
    bounds = np.flatnonzero(np.diff(hops[order])) + 1
    for idx in np.split(order, bounds):           # farthest level first
        idx = idx[down[idx] >= 0]
        np.add.at(size, down[idx], size[idx])
    return size


class Decomposition:
    """Result of `decompose`: which reaches each rank routes as tributaries, the mainstem (routed by rank 0 after
    the hand-off) and the tributary outlets that feed it."""

    def __init__(self, nparts, trib, mainstem, outlets, outlet_owner):
        self.nparts = nparts
        self.trib = trib                  # list[nparts] of sorted global reach indices
        self.mainstem = mainstem          # sorted global reach indices (may be empty)
        self.outlets = outlets            # global reach indices of tributary outlets draining into the mainstem,
        self.outlet_owner = outlet_owner  # ... sorted by (owner rank, reach index), and their owner ranks

    def outlets_of(self, rank):
        return self.outlets[self.outlet_owner == rank]

    def slot_range(self, rank):
        lo = int(np.searchsorted(self.outlet_owner, rank, side="left"))
        hi = int(np.searchsorted(self.outlet_owner, rank, side="right"))
        return lo, hi


def decompose(net: RiverNetwork, nparts: int, mainstem_cost: float = 60.0, mainstem_div: float | None = None) -> Decomposition:
    """The reference's rule: a reach with more than nRch/nparts upstream reaches is MAINSTEM
    (domain_decomposition.f90:508-519); every maximal subtree hanging off the mainstem, and every whole basin
    that has no mainstem, is a TRIBUTARY domain (:640-717); domains go largest-first to the least-loaded rank
    (:791-809).  The mainstem is routed by rank 0, whose load is pre-charged with `mainstem_cost` reach
    equivalents per mainstem reach: a mainstem wavefront holds a handful of reaches and is latency-bound (measured on
    8 B200s, 3 M reaches: 1091 mainstem reaches cost 38 ms per 192-step batch, what ~90 tributary reaches each cost)."""
    n = net.nRch
    down = _down_index(net)
    size = upstream_size(net)
    # mainstem_div > 1 lowers the threshold to nRch / (nparts * mainstem_div): more (and shallower) tributary domains, a longer
    # mainstem.  Every domain is swept stage by stage and a stage costs a fixed latency on the device, so the time of a rank
    # follows the DEPTH of its deepest tributary, not only its number of reaches; the mainstem overlaps the next batch's
    # tributaries on its own stream (multi.py), so depth moved there is hidden until the two are level.  None: environment
    # MR_MAINSTEM_DIV, default 1 = the reference's rule.
    if mainstem_div is None:
        import os
        mainstem_div = float(os.environ.get("MR_MAINSTEM_DIV", "1"))
    max_segs = int(n // (nparts * max(mainstem_div, 1e-9)))
    is_main = size > max_segs if nparts > 1 else np.zeros(n, dtype=bool)
    # root of the tributary domain of every non-mainstem reach: follow downstream until the next reach is
    # mainstem or there is none
    stop = (down < 0) | is_main[np.where(down >= 0, down, 0)]
    root = np.where(stop | is_main, np.arange(n), down)
    while True:
        nxt = root[root]
        if np.array_equal(nxt, root):
            break
        root = nxt
    trib_mask = ~is_main
    roots, inv, counts = np.unique(root[trib_mask], return_inverse=True, return_counts=True)
    load = np.zeros(nparts, dtype=np.float64)
    load[0] = mainstem_cost * float(is_main.sum())
    owner_of_root = np.empty(roots.size, dtype=np.int64)
    for b in np.argsort(-counts, kind="stable"):
        k = int(np.argmin(load))
        owner_of_root[b] = k
        load[k] += counts[b]
    owner = np.full(n, -1, dtype=np.int64)
    owner[trib_mask] = owner_of_root[inv]
    trib = [np.flatnonzero(owner == k) for k in range(nparts)]
    mainstem = np.flatnonzero(is_main)
    is_outlet = trib_mask & (down >= 0) & is_main[np.where(down >= 0, down, 0)]
    out = np.flatnonzero(is_outlet)
    key = np.lexsort((out, owner[out]))
    out = out[key]
    return Decomposition(nparts, trib, mainstem, out, owner[out])


def mainstem_network(net: RiverNetwork, dec: Decomposition) -> RiverNetwork:
    """Mainstem reaches plus the tributary outlets as ghost reaches, in the original index order (so the upstream
    lists keep the reference's UREACHI order); only mainstem HRUs are kept."""
    idx = np.sort(np.concatenate([dec.mainstem, dec.outlets]))
    sub = subnetwork(net, idx)
    ghost = np.isin(net.segId[idx], net.segId[dec.outlets])
    order = np.argsort(net.segId, kind="stable")
    sid = net.segId[order]
    pos = np.clip(np.searchsorted(sid, sub.hruSegId), 0, net.nRch - 1)
    hru_rch = order[pos]                                   # global reach index of each kept HRU
    keep = ~np.isin(hru_rch, dec.outlets)
    sub.meta["hru_index"] = sub.meta["hru_index"][keep]
    sub.hruId, sub.hruSegId, sub.area = sub.hruId[keep], sub.hruSegId[keep], sub.area[keep]
    sub.meta["ghost_mask"] = ghost
    return sub
