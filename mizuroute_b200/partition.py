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
                       meta=dict(net.meta))
    sub.meta["hru_index"] = np.flatnonzero(hru_keep)      # columns of the global runoff array this domain reads
    sub.meta["reach_index"] = np.asarray(reaches)
    return sub
