// Data assimilation for the host builds (test infrastructure only): the gauge observations given with *_emul_set_da are
// staged as mr_upload_obs does and turned into Qobs / Qelapsed rows by da_rows (mr_dev.h), the code k_da_rows runs.
#pragma once
#include <vector>
#include "../../mizuroute_b200/csrc/mr_dev.h"
#include "../../mizuroute_b200/csrc/mr_topo.h"

struct DaEmul {
    int blend = 0, trend = 0; const int *hasRecord = nullptr; const double *obs = nullptr; double *qerrOut = nullptr;
    std::vector<double> rows, qobsState, qerr; std::vector<int> el, elState; std::vector<unsigned char> rec;
    // one-shot: consumed by the next run
    bool attach(mr::DevNet &d, const mr::Topology &T, int method, int nSteps) {
        if (!obs) return false;
        const int N = d.nRch;
        rows.assign((size_t)nSteps * N, 0.0); el.assign((size_t)nSteps * N, 0); qobsState.assign(N, 0.0); elState.assign(N, 0); qerr.assign(N, 0.0);
        rec.assign(nSteps, 1);
        for (int t = 0; t < nSteps; ++t) {
            if (hasRecord) rec[t] = hasRecord[t] ? 1 : 0;
            for (int p = 0; p < N; ++p) rows[(size_t)t * N + p] = obs[(size_t)t * N + T.pos2rch[p]];
        }
        for (int p = 0; p < N; ++p) mr::da_rows(rows.data(), el.data(), qobsState.data(), elState.data(), rec.data(), N, p, nSteps);
        d.daQobs = rows.data(); d.daElapsed = el.data(); d.qerr[method] = qerr.data(); d.qBlendPeriod = blend; d.qErrTrend = trend;
        return true;
    }
    void finish(const mr::Topology &T) {
        if (obs && qerrOut) for (size_t r = 0; r < qerr.size(); ++r) qerrOut[r] = qerr[T.rch2pos[r]];
        obs = nullptr; hasRecord = nullptr; qerrOut = nullptr;
    }
};
