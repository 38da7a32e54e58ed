// Single-lane HOST build of the KWT reach step: the thread-per-task path (mizuroute_b200/csrc/mr_kwt_scalar.cuh) first and,
// for the tasks it defers, the warp-cooperative path (mizuroute_b200/csrc/mr_kwt.cuh) -- the order k_route_kwt uses.
//
// TEST INFRASTRUCTURE ONLY.  mr_lanes.h maps a "team" to one lane when compiled without nvcc, so the very same
// source that runs one warp per (reach, step) on the GPU is executed here serially, reach by reach in stage
// order, and compared bit-for-bit with the CPU oracle (tests/test_kwt_emul.py).  This checks the restructured
// algorithm (order-independent merge ranking, linked-list thinning, cached shock crossings) without a GPU.
// It is not a CPU fallback: nothing under mizuroute_b200/ builds, loads or links this file.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../mizuroute_b200/csrc/mr_kwt_scalar.cuh"
#include "../../mizuroute_b200/csrc/mr_topo.h"

using namespace mr;

// mode 0: scalar path first, team path for what it defers (as on the device); 1: team path only
static int g_mode = 0;
static long g_scalar_done = 0, g_heavy_done = 0, g_deferred = 0;
extern "C" void kwt_emul_set_mode(int mode) { g_mode = mode; }
extern "C" long kwt_emul_count(int which) { return which == 0 ? g_scalar_done : (which == 1 ? g_deferred : g_heavy_done); }

extern "C" int kwt_emul_run(int nRch, int nHRU, const int *segId, const int *downSegId, const int *hruSegId, const double *hruArea,
                            const double *length, const double *slope, double mann_n, double wscale, double dt, int nSteps,
                            const double *qr /* [nSteps+1][nRch] BASIN_QR(1) before step 0 and after every step, caller order */,
                            const double *wm_flux /* [nSteps][nRch] REACH_WM_FLUX, caller order, or NULL: the EXT instantiation is used then */,
                            double *q_out /* [nSteps][nRch] REACH_Q, caller order */, int *n_out /* [nRch] live particles */, char *msg) {
    Topology T;
    std::string terr;
    if (build_topology(nRch, nHRU, segId, downSegId, hruSegId, hruArea, T, terr)) { std::snprintf(msg, 256, "%s", terr.c_str()); return 1; }
    const int N = nRch;
    std::vector<double> rlen(N), rslp(N), rwid(N), rman(N, mann_n), kK(N), kAK(N);
    std::vector<int> flags(N, 0);
    for (int p = 0; p < N; ++p) {
        const int r = T.pos2rch[p];
        rlen[p] = length[r]; rslp[p] = std::fmax(slope[r], 1.e-6); rwid[p] = wscale * std::sqrt(T.totArea[p]);
        kK[p] = std::sqrt(rslp[p]) / rman[p]; kAK[p] = (5.0 / 3.0) * std::pow(kK[p], 1.0 / (5.0 / 3.0));   // k_kwt_params
    }
    std::vector<double> qrSer((size_t)(nSteps + 1) * N), qSer((size_t)nSteps * N, 0.0), inflow(N, 0.0), T0s(nSteps), T1s(nSteps);
    for (int t = 0; t <= nSteps; ++t) for (int r = 0; r < N; ++r) qrSer[(size_t)t * N + T.rch2pos[r]] = qr[(size_t)t * N + r];
    std::vector<int> kwN[2], kwNR[2];
    std::vector<double> kwQF[2], kwTI[2], kwTR[2];
    for (int b = 0; b < 2; ++b) {
        kwN[b].assign(N, 0); kwNR[b].assign(N, 0);
        kwQF[b].assign((size_t)KWP * N, 0.0); kwTI[b] = kwQF[b]; kwTR[b] = kwQF[b];
    }
    int err[4] = {0, 0, 0, 0};
    DevNet d{};
    d.nRch = N; d.nHRU = nHRU; d.nStage = T.nStage; d.nHead = T.nHead; d.dt = dt;
    d.stageOf = T.stageOf.data(); d.upPtr = T.upPtr.data(); d.upIdx = T.upIdx.data(); d.nGood = T.nGood.data(); d.flags = flags.data();
    d.rlength = rlen.data(); d.rslope = rslp.data(); d.rwidth = rwid.data(); d.rmann = rman.data(); d.kwK = kK.data(); d.kwAK = kAK.data();
    d.qrSer = qrSer.data(); d.qSer[M_KWT] = qSer.data(); d.inflow[M_KWT] = inflow.data();
    d.T0s = T0s.data(); d.T1s = T1s.data();
    for (int b = 0; b < 2; ++b) { d.kwN[b] = kwN[b].data(); d.kwNR[b] = kwNR[b].data(); d.kwQF[b] = kwQF[b].data(); d.kwTI[b] = kwTI[b].data(); d.kwTR[b] = kwTR[b].data(); }
    d.err = err; d.kwCount = nullptr;
    std::vector<KwsRec> recs(N);
    for (int p = 0; p < N; ++p) recs[p] = kws_make_record(d, p);            // k_kws_records
    d.kwRec = recs.data();
    double t0 = 0.0, t1 = dt;
    for (int t = 0; t < nSteps; ++t) { T0s[t] = t0; T1s[t] = t1; t0 = t1; t1 = t0 + dt; }
    std::vector<double> fS;
    if (wm_flux) {                                          // stage order, as mr_upload_wm
        fS.resize((size_t)nSteps * N);
        for (int t = 0; t < nSteps; ++t) for (int p = 0; p < N; ++p) fS[(size_t)t * N + p] = wm_flux[(size_t)t * N + T.pos2rch[p]];
        d.wmFlux = fS.data();
    }
    static KwtScratch S;
    static KwtScratchSmall Ssmall;
    long retries = 0;
    g_scalar_done = g_heavy_done = g_deferred = 0;
    static KwsWarp<KWS_NL, false> WL;                   // one-lane "warps" of the light and the heavy instantiation
    static KwsWarp<KWS_NH, true> WH;
    for (int t = 0; t < nSteps; ++t) {
        const int b = t & 1;
        for (int p = 0; p < T.nHead; ++p) {                 // k_headwater<M_KWT>
            inflow[p] = 0.0; qSer[(size_t)t * N + p] = qrSer[(size_t)(t + 1) * N + p];
            kwN[b][p] = 1; kwNR[b][p] = 0;
        }
        for (int p = T.nHead; p < N; ++p) {                 // stage order: upstream before downstream
            if (g_mode == 0) {                              // lane-per-task paths (light, then heavy); what they have written before giving up is rewritten below
                int rc = wm_flux ? kws_warp_route<true, KWS_NL, false>(d, WL, p, t, (long long)t, T0s[t], T1s[t], true)
                                 : kws_warp_route<false, KWS_NL, false>(d, WL, p, t, (long long)t, T0s[t], T1s[t], true);
                if (rc == KWS_DONE) { ++g_scalar_done; continue; }
                if (rc == KWS_HEAVY) {
                    rc = wm_flux ? kws_warp_route<true, KWS_NH, true>(d, WH, p, t, (long long)t, T0s[t], T1s[t], true)
                                 : kws_warp_route<false, KWS_NH, true>(d, WH, p, t, (long long)t, T0s[t], T1s[t], true);
                    if (rc == KWS_DONE) { ++g_heavy_done; continue; }
                }
                ++g_deferred;
            }
            // the shared-memory-sized scratch first, the full-capacity one on KWT_RETRY -- as k_route_kwt does
            if (wm_flux) {
                if (kwt_reach_team<KwtScratchSmall, false, true>(d, Ssmall, p, t, (long long)t, T0s[t], T1s[t]) == KWT_RETRY) {
                    ++retries;
                    kwt_reach_team<KwtScratch, false, true>(d, S, p, t, (long long)t, T0s[t], T1s[t]);
                }
            } else if (kwt_reach_team(d, Ssmall, p, t, (long long)t, T0s[t], T1s[t]) == KWT_RETRY) {
                ++retries;
                kwt_reach_team(d, S, p, t, (long long)t, T0s[t], T1s[t]);
            }
            if (err[0]) { std::snprintf(msg, 256, "ierr %d at position %d (reach %d) site %d step %d", err[0], err[1], T.pos2rch[err[1]], err[2], t); return err[0]; }
        }
    }
    for (int t = 0; t < nSteps; ++t) for (int r = 0; r < N; ++r) q_out[(size_t)t * N + r] = qSer[(size_t)t * N + T.rch2pos[r]];
    const int b = (nSteps - 1) & 1;
    for (int r = 0; r < N; ++r) { const int p = T.rch2pos[r]; n_out[r] = kwN[b][p] - (kwNR[b][p] > 0 ? kwNR[b][p] - 1 : 0); }
    std::snprintf(msg, 256, "retries=%ld", retries);
    return 0;
}
