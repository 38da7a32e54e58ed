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
// branch lists (main_route.f90:356-403).  Here
//   * reaches WITHOUT upstream reaches ("headwaters", ~half of a river network) depend on nothing but their
//     own lateral inflow; they are stored first and routed for all steps of a batch in one launch;
//   * every other ("interior") reach gets  stage = Dmax - (hops to its outlet), so that
//     stage(upstream interior reach) == stage(downstream) - 1 EXACTLY.  Interior reaches are stored stage by
//     stage, inside a stage in the order their downstream reaches appear in the next stage.
// Consequences the kernels rely on:
//   * a time-skewed wavefront  w = stage + step  touches a contiguous position range of interior reaches only
//     (warps are not diluted by trivial headwater lanes);
//   * an interior producer is exactly one wavefront ahead of its consumer, so 2-deep particle buffers suffice;
//   * upstream gathers (CSR upPtr/upIdx, kept in the reference's UREACHI order) are monotone in position.
#pragma once
#include <algorithm>
#include <cmath>
#include <cfloat>
#include <cstdint>
#include <cstdlib>
#include <string>
#include <vector>

namespace mr {

struct Topology {
    int nRch = 0, nHRU = 0, nStage = 0, maxUps = 0, nHead = 0;
    std::vector<int> pos2rch, rch2pos;          // device position <-> caller's reach index
    std::vector<int> stagePtr;                  // [nStage+1] position range of each stage
    std::vector<int> stageOf;                   // [nRch] by position
    std::vector<int> upPtr, upIdx, nGood;       // CSR by position (upIdx holds positions), nGood by position
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

// ghostKind (may be null): 0 = ordinary reach; 1 / 2 = ghost copy of a tributary outlet that is routed in another
// domain (the reference appends such reaches to the mainstem slab, mpi_process.f90:593-607,680), 1 if that outlet
// has no contributing upstream reach (a "headwater" there), 2 if it has.  A ghost has no upstream reaches and no
// HRUs here; its total area (ghostTotArea) is the one its owner computed.
inline int build_topology(int nRch, int nHRU, const int *segId, const int *downSegId, const int *hruSegId,
                          const double *hruArea, Topology &T, std::string &err,
                          const int *ghostKind = nullptr, const double *ghostTotArea = nullptr) {
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

    // interior reaches stage by stage (upstream-most stage first), headwaters pulled out in front
    auto isHead = [&](int r) { return uPtr[r + 1] == uPtr[r] && !(ghostKind && ghostKind[r] == 2); };
    std::vector<int> interior;                  // caller indices in device order
    std::vector<int> stageCount(T.nStage, 0);
    interior.reserve(nRch);
    for (int s = 0; s < T.nStage; ++s)
        for (int r : byHops[T.nStage - 1 - s]) if (!isHead(r)) { interior.push_back(r); stageCount[s]++; }
    // Within a stage the order is free.  Reaches of similar KWT work -- same number of (non-headwater) upstream reaches,
    // similar upstream network size -- are grouped, so that the tasks that share a warp (MR_TEAM lanes each) follow
    // similar paths and stay converged: measured on the 3 M-reach network, KWT 512 -> 407 ms per 96 hourly steps.
    // MR_STAGE_SORT=0 keeps the order in which the downstream reaches appear in the next stage; 1/3/4 are other keys.
    const int sortMode = std::getenv("MR_STAGE_SORT") ? std::atoi(std::getenv("MR_STAGE_SORT")) : 2;
    if (sortMode > 0) {
        std::vector<int> usize(nRch, 1), nInt(nRch, 0);
        for (int hh = (int)byHops.size() - 1; hh >= 0; --hh)
            for (int r : byHops[hh]) if (down[r] >= 0) { usize[down[r]] += usize[r]; if (!isHead(r)) nInt[down[r]]++; }
        auto key = [&](int r) {
            int lg = 0; for (int v = usize[r]; v > 1; v >>= 1) ++lg;
            const int ni = nInt[r] > 3 ? 3 : nInt[r], nu = uPtr[r + 1] - uPtr[r] > 7 ? 7 : uPtr[r + 1] - uPtr[r];
            if (sortMode == 2) { int lg2 = 0; for (double v = (double)usize[r]; v > 1.0; v /= 1.41421356) ++lg2; return (ni * 8 + nu) * 128 + lg2; }
            if (sortMode == 3) return lg;
            if (sortMode == 4) return lg * 64 + ni * 8 + nu;
            return ni * 64 + lg;
        };
        size_t k0 = 0;
        for (int s2 = 0; s2 < T.nStage; ++s2) {
            std::stable_sort(interior.begin() + k0, interior.begin() + k0 + stageCount[s2], [&](int a, int b) { return key(a) < key(b); });
            k0 += stageCount[s2];
        }
    }
    std::vector<int> heads;
    heads.reserve(nRch - interior.size());
    for (int r : interior)
        for (int m = uPtr[r]; m < uPtr[r + 1]; ++m) if (isHead(uIdx[m])) heads.push_back(uIdx[m]);
    for (int i = 0; i < nRch; ++i) if (isHead(i) && down[i] < 0) heads.push_back(i);      // isolated reaches
    T.nHead = (int)heads.size();
    if ((size_t)T.nHead + interior.size() != (size_t)nRch) { err = "build_topology/internal: reach partition is inconsistent"; return 60; }

    T.pos2rch.resize(nRch); T.rch2pos.resize(nRch); T.stagePtr.assign(T.nStage + 1, 0); T.stageOf.assign(nRch, -1);
    int p = 0;
    for (int r : heads) { T.pos2rch[p] = r; T.rch2pos[r] = p; ++p; }
    for (int s = 0, k = 0; s < T.nStage; ++s) {
        T.stagePtr[s] = p;
        for (int c = 0; c < stageCount[s]; ++c, ++k) { const int r = interior[k]; T.pos2rch[p] = r; T.rch2pos[r] = p; T.stageOf[p] = s; ++p; }
    }
    T.stagePtr[T.nStage] = p;

    T.upPtr.assign(nRch + 1, 0); T.upIdx.resize(uIdx.size()); T.downPos.assign(nRch, -1);
    T.maxUps = 0;
    for (int q = 0; q < nRch; ++q) {
        const int r = T.pos2rch[q];
        const int n = uPtr[r + 1] - uPtr[r];
        T.upPtr[q + 1] = T.upPtr[q] + n;
        for (int m = 0; m < n; ++m) {
            const int up = T.rch2pos[uIdx[uPtr[r] + m]];
            T.upIdx[T.upPtr[q] + m] = up;
            if (T.stageOf[up] >= 0 && T.stageOf[up] != T.stageOf[q] - 1) { err = "build_topology/internal: stage lag is not one"; return 60; }
        }
        T.downPos[q] = down[r] >= 0 ? T.rch2pos[down[r]] : -1;
        T.maxUps = std::max(T.maxUps, n);
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
        for (int m = T.upPtr[q]; m < T.upPtr[q + 1]; ++m) ups = ups + T.totArea[T.upIdx[m]];
        for (int m = T.hruPtr[q]; m < T.hruPtr[q + 1]; ++m) bas += hruArea[T.hruIdx[m]];
        T.basArea[q] = bas; T.upsArea[q] = ups; T.totArea[q] = bas + ups;
        for (int m = T.hruPtr[q]; m < T.hruPtr[q + 1]; ++m) T.hruWgt[m] = hruArea[T.hruIdx[m]] / bas;
        T.nGood[q] = (T.totArea[q] > DBL_MIN) ? (T.upPtr[q + 1] - T.upPtr[q]) : 0;
        const int gk = ghostKind ? ghostKind[T.pos2rch[q]] : 0;
        if (gk) { T.totArea[q] = ghostTotArea[T.pos2rch[q]]; T.nGood[q] = gk == 2 ? 1 : 0; }
    }
    return 0;
}

}  // namespace mr
