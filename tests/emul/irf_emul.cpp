// Single-threaded HOST build of the accumulation and impulse-response-function reach steps
// (mizuroute_b200/csrc/mr_irf.cuh), with the reach unit hydrographs of mr_uh.h and lake reaches from mr_lake.cuh.
//
// TEST INFRASTRUCTURE ONLY.  Steps the reaches in stage order as the GPU kernels do (ring-buffered QFUTURE_IRF, slot-major
// arrays) and is compared bit for bit with the CPU oracle (tests/test_irf_emul.py).  Not a CPU fallback: nothing under
// mizuroute_b200/ builds or loads this file.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../mizuroute_b200/csrc/mr_irf.cuh"
#include "../../mizuroute_b200/csrc/mr_lake.cuh"
#include "../../mizuroute_b200/csrc/mr_topo.h"
#include "../../mizuroute_b200/csrc/mr_uh.h"
#include "da_emul.h"

using namespace mr;

static DaEmul g_da;
// gauge observations of the NEXT irf_emul_run: obs [nSteps][nRch] caller order, hasRecord [nSteps] or NULL; qerr_out [nRch]
extern "C" void irf_emul_set_da(int blend, int trend, const int *hasRecord, const double *obs, double *qerr_out) {
    g_da.blend = blend; g_da.trend = trend; g_da.hasRecord = hasRecord; g_da.obs = obs; g_da.qerrOut = qerr_out;
}

extern "C" int irf_emul_run(int nRch, int nHRU, const int *segId, const int *downSegId, const int *hruSegId, const double *hruArea,
                            const double *length, const double *slope, const int *islake /* or NULL */, const int *lakeType, const double *maxS,
                            const double *coef, const double *pw, const double *s0, double wscale, double dt, double velo, double diff,
                            int hw_drain_point, double min_length_route, int lakeInputOption, int nSteps,
                            const double *qr /* [nSteps+1][nRch] BASIN_QR(1), caller order */,
                            const double *wm_flux, const double *wm_vol /* [nSteps][nRch] caller order, or NULL */, const double *targ /* [nRch] 0/1 or NULL */,
                            int jumpStart,
                            double *q_sum /* [nSteps][nRch] */, double *q_irf /* [nSteps][nRch] */, double *vol_out, double *wb_out,
                            double *qfut_out /* [nRch][maxtdh] logical order */, int *maxtdh_out, char *msg) {
    Topology T;
    std::string terr;
    if (build_topology(nRch, nHRU, segId, downSegId, hruSegId, hruArea, T, terr)) { std::snprintf(msg, 256, "%s", terr.c_str()); return 1; }
    const int N = nRch;
    std::vector<double> rlen(N), rwid(N), dMaxS(N, 0.0), dCoef(N, 0.0), dPw(N, 0.0), dS0(N, 0.0);
    std::vector<int> flags(N, 0), ltype(N, MR_LAKE_DOLL03), ntdh(N, 1), slot(N, -1);
    std::vector<double> rowMajor((size_t)N * 240);
    int mx = 1, nLake = 0;
    for (int p = 0; p < N; ++p) {
        const int r = T.pos2rch[p];
        rlen[p] = length[r]; rwid[p] = wscale * std::sqrt(T.totArea[p]);
        if (islake && islake[r] == 1) { flags[p] |= FLAG_LAKE; slot[p] = nLake++; ltype[p] = lakeType[r]; dMaxS[p] = maxS[r]; dCoef[p] = coef[r]; dPw[p] = pw[r]; dS0[p] = s0[r]; }
        double *u = &rowMajor[(size_t)p * 240];
        const int n = build_reach_uh(rlen[p], dt, velo, diff, u);                 // as mr_set_network
        if (flags[p] & FLAG_LAKE) { for (int k = 0; k < n; ++k) u[k] = 0.0; u[0] = 1.0; }
        ntdh[p] = n; if (n > mx) mx = n;
    }
    std::vector<double> uh((size_t)mx * N, 0.0), qfut((size_t)mx * N, 0.0);
    for (int p = 0; p < N; ++p) for (int k = 0; k < ntdh[p]; ++k) uh[(size_t)k * N + p] = rowMajor[(size_t)p * 240 + k];
    std::vector<double> qrSer((size_t)(nSteps + 1) * N), qS((size_t)nSteps * N, 0.0), qI((size_t)nSteps * N, 0.0), inflow(N, 0.0), vol0(N, 0.0), vol1(N, 0.0), wb(N, 0.0);
    for (int t = 0; t <= nSteps; ++t) for (int r = 0; r < N; ++r) qrSer[(size_t)t * N + T.rch2pos[r]] = qr[(size_t)t * N + r];
    int err[4] = {0, 0, 0, 0};
    DevNet d{};
    d.nRch = N; d.nHRU = nHRU; d.nStage = T.nStage; d.nHead = T.nHead; d.dt = dt; d.hwDrain = hw_drain_point; d.minLengthRoute = min_length_route;
    d.lakeInputOption = lakeInputOption; d.isLakeSim = islake ? 1 : 0; d.maxtdh = mx;
    d.stageOf = T.stageOf.data(); d.upPtr = T.upPtr.data(); d.upIdx = T.upIdx.data(); d.nGood = T.nGood.data(); d.flags = flags.data();
    d.rlength = rlen.data(); d.rwidth = rwid.data(); d.ntdh = ntdh.data(); d.uh = uh.data(); d.qfutIrf = qfut.data();
    d.lakeType = ltype.data(); d.d03MaxS = dMaxS.data(); d.d03Coef = dCoef.data(); d.d03Pow = dPw.data(); d.d03S0 = dS0.data();
    d.lakeSlot = slot.data(); d.nLake = nLake;
    d.qrSer = qrSer.data(); d.qSer[M_SUM] = qS.data(); d.qSer[M_IRF] = qI.data(); d.inflow[M_IRF] = inflow.data();
    d.vol0[M_IRF] = vol0.data(); d.vol1[M_IRF] = vol1.data(); d.wb[M_IRF] = wb.data(); d.err = err;
    // water management: rows permuted to stage order as mr_upload_wm does; the EXT instantiations are used then
    const bool ext = g_da.attach(d, T, M_IRF, nSteps) || wm_flux || wm_vol;
    std::vector<double> fS, vS; std::vector<unsigned char> tg(N, 0);
    auto stage = [&](const double *src, std::vector<double> &dst) {
        dst.resize((size_t)nSteps * N);
        for (int t = 0; t < nSteps; ++t) for (int p = 0; p < N; ++p) dst[(size_t)t * N + p] = src[(size_t)t * N + T.pos2rch[p]]; };
    if (wm_flux) { stage(wm_flux, fS); d.wmFlux = fS.data(); }
    if (wm_vol && islake) { stage(wm_vol, vS); d.wmVol = vS.data(); }
    if (targ) { for (int p = 0; p < N; ++p) tg[p] = targ[T.pos2rch[p]] != 0.0; d.lakeTargVol = tg.data(); }
    d.volJumpStart = jumpStart;
    for (int t = 0; t < nSteps; ++t)
        for (int p = 0; p < N; ++p) {                        // route_reach<M_SUM>, route_reach<M_IRF>: stage order
            sum_reach(d, p, t);
            if (ext) { if (flags[p] & FLAG_LAKE) lake_reach<M_IRF, true>(d, p, t, (long long)t); else irf_reach<true>(d, p, t, (long long)t); }
            else if (flags[p] & FLAG_LAKE) lake_reach<M_IRF, false>(d, p, t, (long long)t); else irf_reach<false>(d, p, t, (long long)t);
            if (err[0]) { std::snprintf(msg, 256, "ierr %d at position %d site %d step %d", err[0], err[1], err[2], t); return err[0]; }
        }
    for (int t = 0; t < nSteps; ++t) for (int r = 0; r < N; ++r) {
        q_sum[(size_t)t * N + r] = qS[(size_t)t * N + T.rch2pos[r]]; q_irf[(size_t)t * N + r] = qI[(size_t)t * N + T.rch2pos[r]]; }
    *maxtdh_out = mx;
    for (int r = 0; r < N; ++r) {
        const int p = T.rch2pos[r];
        vol_out[r] = vol1[p]; wb_out[r] = wb[p];
        for (int k = 0; k < 240; ++k) qfut_out[(size_t)r * 240 + k] = k < ntdh[p] ? qfut[(size_t)(((long long)nSteps + k) % ntdh[p]) * N + p] : 0.0;   // as mr_get_state
    }
    g_da.finish(T);
    std::snprintf(msg, 256, "ok");
    return 0;
}
