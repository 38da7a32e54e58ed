// Host-side river-network preparation: what augment_ntopo (process_ntopo.f90:39-266) derives from the
// network-file variables, re-derived in O(N log N) and laid out for the device.
//
//   id -> index join                    network_topo.f90:362-464  (downReachIndex)
//   HRU -> reach lists, weights          network_topo.f90:46-196   (hru2segment)
//   upstream lists                       network_topo.f90:202-311  (up2downSegment)
//   basin / upstream / total areas,
//   goodBas                              network_topo.f90:637-779  (reach_list; goodBas :769-775 is
//                                                                   all-or-nothing per reach)
//   lake inlets                          network_topo.f90:958-985
//
// Device order ("stage order").  The reference sweeps reaches upstream->downstream in Strahler-order /
// branch lists (main_route.f90:356-403).  Here every reach gets  stage = Dmax - (hops to its outlet),
// so that stage(upstream) == stage(downstream) - 1 EXACTLY.  Reaches are stored stage by stage, and
// inside a stage in the order their downstream reaches appear in the next stage, siblings in the
// reference's UREACHI order.  Consequences the kernels rely on:
//   * the upstream reaches of position p are the contiguous positions [up_first[p], up_first[p]+nUps[p]);
//   * a time-skewed wavefront  w = stage + step  touches a contiguous position range;
//   * a producer is exactly one wavefront ahead of its consumer, so 2-deep buffers suffice.
#pragma once
#include <algorithm>
#include <cmath>
#include <cfloat>
#include <cstdint>
#include <string>
#include <vector>

namespace mr {

struct Topology {
    int nRch = 0, nHRU = 0, nStage = 0, maxUps = 0;
    std::vector<int> pos2rch, rch2pos;          // device position <-> caller's reach index
    std::vector<int> stagePtr;                  // [nStage+1] position range of each stage
    std::vector<int> stageOf;                   // [nRch] by position
    std::vector<int> upFirst, nUps, nGood;      // by position
    std::vector<int> downPos;                   // by position, -1 for outlets
    std::vector<int> hruPtr, hruIdx;            // CSR by position; hruIdx in caller's HRU order
    std::vector<double> hruWgt;
    std::vector<double> basArea, upsArea, totArea;   // by position
    std::vector<int> downIndex;                 // by caller's reach index
};

// id -> first index holding that id; ids <= 0 mean "none"
inline void join_ids(const int *keys, int nKeys, const int *ids, int nIds, std::vector<int> &out) {
    std::vector<std::pair<int, int>> tab(nIds);
    for (int i = 0; i < nIds; ++i) tab[i] = {ids[i], i};
    std::sort(tab.begin(), tab.end());
    out.assign(nKeys, -1);
    for (int i = 0; i < nKeys; ++i) {
        if (keys[i] <= 0) continue;
        auto it = std::lower_bound(tab.begin(), tab.end(), std::make_pair(keys[i], -1));
        if (it != tab.end() && it->first == keys[i]) out[i] = it->second;
    }
}

inline int build_topology(int nRch, int nHRU, const int *segId, const int *downSegId, const int *hruSegId,
                          const double *hruArea, Topology &T, std::string &err) {
    T.nRch = nRch; T.nHRU = nHRU;
    join_ids(downSegId, nRch, segId, nRch, T.downIndex);
    const std::vector<int> &down = T.downIndex;

    // upstream lists in the caller's index order (= UREACHI order)
    std::vector<int> uPtr(nRch + 1, 0), uIdx;
    for (int i = 0; i < nRch; ++i) if (down[i] >= 0) uPtr[down[i] + 1]++;
    for (int i = 0; i < nRch; ++i) uPtr[i + 1] += uPtr[i];
    uIdx.resize(uPtr[nRch]);
    {
        std::vector<int> fill(uPtr.begin(), uPtr.end() - 1);
        for (int i = 0; i < nRch; ++i) if (down[i] >= 0) uIdx[fill[down[i]]++] = i;
    }

    // hops to outlet, walking down from the outlets (BFS); unreachable reaches => cycle
    std::vector<int> hops(nRch, -1), frontier, next;
    for (int i = 0; i < nRch; ++i) if (down[i] < 0) { hops[i] = 0; frontier.push_back(i); }
    std::vector<std::vector<int>> byHops;       // reaches per hop count, in stage-internal order
    int visited = 0;
    while (!frontier.empty()) {
        visited += (int)frontier.size();
        next.clear();
        const int h = (int)byHops.size();
        for (int r : frontier)
            for (int m = uPtr[r]; m < uPtr[r + 1]; ++m) { hops[uIdx[m]] = h + 1; next.push_back(uIdx[m]); }
        byHops.push_back(frontier);
        frontier.swap(next);
    }
    if (visited != nRch) { err = "build_topology/river network has a cycle or a dangling downstream id"; return 20; }
    T.nStage = (int)byHops.size();

    T.pos2rch.resize(nRch); T.rch2pos.resize(nRch); T.stagePtr.assign(T.nStage + 1, 0); T.stageOf.resize(nRch);
    int p = 0;
    for (int s = 0; s < T.nStage; ++s) {
        const std::vector<int> &grp = byHops[T.nStage - 1 - s];
        T.stagePtr[s] = p;
        for (int r : grp) { T.pos2rch[p] = r; T.rch2pos[r] = p; T.stageOf[p] = s; ++p; }
    }
    T.stagePtr[T.nStage] = p;

    T.upFirst.assign(nRch, 0); T.nUps.assign(nRch, 0); T.downPos.assign(nRch, -1);
    T.maxUps = 0;
    for (int q = 0; q < nRch; ++q) {
        const int r = T.pos2rch[q];
        const int n = uPtr[r + 1] - uPtr[r];
        T.nUps[q] = n;
        T.upFirst[q] = n ? T.rch2pos[uIdx[uPtr[r]]] : 0;
        T.downPos[q] = down[r] >= 0 ? T.rch2pos[down[r]] : -1;
        T.maxUps = std::max(T.maxUps, n);
        for (int m = 0; m < n; ++m)
            if (T.rch2pos[uIdx[uPtr[r] + m]] != T.upFirst[q] + m) { err = "build_topology/internal: upstream positions not contiguous"; return 60; }
    }

    // HRU lists in the caller's HRU order
    std::vector<int> hruRch;
    join_ids(hruSegId, nHRU, segId, nRch, hruRch);
    T.hruPtr.assign(nRch + 1, 0);
    for (int i = 0; i < nHRU; ++i) if (hruRch[i] >= 0) T.hruPtr[T.rch2pos[hruRch[i]] + 1]++;
    for (int q = 0; q < nRch; ++q) T.hruPtr[q + 1] += T.hruPtr[q];
    T.hruIdx.resize(T.hruPtr[nRch]); T.hruWgt.resize(T.hruPtr[nRch]);
    {
        std::vector<int> fill(T.hruPtr.begin(), T.hruPtr.end() - 1);
        for (int i = 0; i < nHRU; ++i) if (hruRch[i] >= 0) T.hruIdx[fill[T.rch2pos[hruRch[i]]]++] = i;
    }

    // areas accumulate downstream; positions are already upstream-first
    T.basArea.assign(nRch, 0.0); T.upsArea.assign(nRch, 0.0); T.totArea.assign(nRch, 0.0); T.nGood.assign(nRch, 0);
    for (int q = 0; q < nRch; ++q) {
        double ups = 0.0, bas = 0.0;
        for (int m = 0; m < T.nUps[q]; ++m) ups = ups + T.totArea[T.upFirst[q] + m];
        for (int m = T.hruPtr[q]; m < T.hruPtr[q + 1]; ++m) bas += hruArea[T.hruIdx[m]];
        T.basArea[q] = bas; T.upsArea[q] = ups; T.totArea[q] = bas + ups;
        for (int m = T.hruPtr[q]; m < T.hruPtr[q + 1]; ++m) T.hruWgt[m] = hruArea[T.hruIdx[m]] / bas;
        T.nGood[q] = (T.totArea[q] > DBL_MIN) ? T.nUps[q] : 0;
    }
    return 0;
}

}  // namespace mr
