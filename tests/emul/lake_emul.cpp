// Single-threaded HOST build of the lake reach step (mizuroute_b200/csrc/mr_lake.cuh: lake_route for endorheic and
// Doll-2003 lakes with the optional evaporation / precipitation forcing) inside a kinematic-wave (mr_euler.cuh) network.
//
// TEST INFRASTRUCTURE ONLY.  Steps the reaches in stage order exactly as the GPU kernels do and is compared bit for bit
// with the CPU oracle (tests/test_lake_emul.py).  Not a CPU fallback: nothing under mizuroute_b200/ builds or loads this.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../mizuroute_b200/csrc/mr_euler.cuh"
#include "../../mizuroute_b200/csrc/mr_lake.cuh"
#include "../../mizuroute_b200/csrc/mr_topo.h"
#include "../../mizuroute_b200/csrc/mr_calendar.h"

using namespace mr;

extern "C" int lake_emul_run(int nRch, int nHRU, const int *segId, const int *downSegId, const int *hruSegId, const double *hruArea,
                             const double *length, const double *slope, const int *islake, const int *lakeType, const double *maxS,
                             const double *coef, const double *pw, const double *s0, double mann_n, double wscale, double dt,
                             int lakeInputOption, double runoffMin, double tconv, double lconv, int nSteps,
                             const double *qr /* [nSteps+1][nRch] BASIN_QR(1), caller order */,
                             const double *evapo, const double *precip /* [nSteps][nHRU] or NULL */,
                             const double *hyp /* [12][nRch] HYP_* in HypeParams order, caller order, or NULL */,
                             int startY, int startM, int startD, double startSec, int noleap /* startY = 0: no calendar */,
                             double *q_out /* [nSteps][nRch] */, double *vol_out, double *wb_out, double *evap_left /* [nRch], last step */, char *msg) {
    Topology T;
    std::string terr;
    if (build_topology(nRch, nHRU, segId, downSegId, hruSegId, hruArea, T, terr)) { std::snprintf(msg, 256, "%s", terr.c_str()); return 1; }
    constexpr int M = M_KW;
    const int N = nRch, nm = n_molecule(M);
    std::vector<double> rlen(N), rslp(N), rwid(N), rman(N, mann_n), rdep(N, 100000.0), zc(N, 0.0), zf(N, 1000.0), rstor(N);
    std::vector<double> dMaxS(N), dCoef(N), dPw(N), dS0(N);
    std::vector<int> flags(N, 0), ltype(N, MR_LAKE_DOLL03), slot(N, -1), pos;
    for (int p = 0; p < N; ++p) {
        const int r = T.pos2rch[p];
        rlen[p] = length[r]; rslp[p] = std::fmax(slope[r], 1.e-6); rwid[p] = wscale * std::sqrt(T.totArea[p]);
        rstor[p] = rdep[p] * (rwid[p] + zc[p] * rdep[p]) * rlen[p];
        if (islake[r] == 1) { flags[p] |= FLAG_LAKE; slot[p] = (int)pos.size(); pos.push_back(p); }
        ltype[p] = lakeType[r]; dMaxS[p] = maxS[r]; dCoef[p] = coef[r]; dPw[p] = pw[r]; dS0[p] = s0[r];
    }
    const int nLake = (int)pos.size();
    std::vector<double> qrSer((size_t)(nSteps + 1) * N), qSer((size_t)nSteps * N, 0.0), inflow(N, 0.0), vol0(N, 0.0), vol1(N, 0.0), wb(N, 0.0),
        mol((size_t)nm * N, 0.0), flood(N, 0.0), ele(N, 0.0), lakeE((size_t)nSteps * (nLake ? nLake : 1)), lakeP(lakeE.size());
    for (int t = 0; t <= nSteps; ++t) for (int r = 0; r < N; ++r) qrSer[(size_t)t * N + T.rch2pos[r]] = qr[(size_t)t * N + r];
    int err[4] = {0, 0, 0, 0};
    DevNet d{};
    d.nRch = N; d.nHRU = nHRU; d.nStage = T.nStage; d.nHead = T.nHead; d.dt = dt; d.hwDrain = 2; d.minLengthRoute = 0.0;
    d.runoffMin = runoffMin; d.tconv = tconv; d.lconv = lconv; d.lakeInputOption = lakeInputOption; d.isLakeSim = 1;
    d.stageOf = T.stageOf.data(); d.upPtr = T.upPtr.data(); d.upIdx = T.upIdx.data(); d.nGood = T.nGood.data(); d.flags = flags.data();
    d.hruPtr = T.hruPtr.data(); d.hruIdx = T.hruIdx.data(); d.hruWgt = T.hruWgt.data(); d.basArea = T.basArea.data();
    d.rlength = rlen.data(); d.rslope = rslp.data(); d.rwidth = rwid.data(); d.rmann = rman.data();
    d.rdepth = rdep.data(); d.sideSlope = zc.data(); d.fldpSlope = zf.data(); d.rstorage = rstor.data();
    d.lakeType = ltype.data(); d.d03MaxS = dMaxS.data(); d.d03Coef = dCoef.data(); d.d03Pow = dPw.data(); d.d03S0 = dS0.data();
    d.qrSer = qrSer.data(); d.qSer[M] = qSer.data(); d.inflow[M] = inflow.data(); d.vol0[M] = vol0.data(); d.vol1[M] = vol1.data();
    d.wb[M] = wb.data(); d.mol[M] = mol.data(); d.floodVol[M] = flood.data(); d.reachEle[M] = ele.data();
    d.err = err; d.lakeSlot = slot.data(); d.nLake = nLake;
    std::vector<HypeParams> hypBySlot(nLake ? nLake : 1);
    std::vector<int> doy(nSteps, 0);
    if (hyp && nLake) {                                      // as mr_set_network / route_device do
        for (int k = 0; k < HYP_COUNT; ++k) for (int sl = 0; sl < nLake; ++sl) reinterpret_cast<double *>(&hypBySlot[sl])[k] = hyp[(size_t)k * N + T.pos2rch[pos[sl]]];
        d.hyp = hypBySlot.data();
    }
    if (startY) {
        for (int t = 0; t < nSteps; ++t) { int mo, dy; step_calendar(startY, startM, startD, startSec, noleap != 0, dt, t, mo, dy, doy[t]); }
        d.stepDoy = doy.data();
    }
    if (evapo && precip && nLake) {                          // k_lake_forcing
        d.evapo = evapo; d.precip = precip; d.lakeEvap = lakeE.data(); d.lakePrecip = lakeP.data();
        for (int t = 0; t < nSteps; ++t) for (int s = 0; s < nLake; ++s) {
            lakeE[(size_t)t * nLake + s] = lake_basin2reach(d, pos[s], evapo + (size_t)t * nHRU);
            lakeP[(size_t)t * nLake + s] = lake_basin2reach(d, pos[s], precip + (size_t)t * nHRU);
        }
    }
    for (int t = 0; t < nSteps; ++t)
        for (int p = 0; p < N; ++p) {                        // route_reach<M_KW, *>: stage order
            if (flags[p] & FLAG_LAKE) lake_reach<M, true>(d, p, t, (long long)t); else kw_dw_reach<M>(d, p, t);
            if (err[0]) { std::snprintf(msg, 256, "ierr %d at position %d site %d step %d", err[0], err[1], err[2], t); return err[0]; }
        }
    for (int t = 0; t < nSteps; ++t) for (int r = 0; r < N; ++r) q_out[(size_t)t * N + r] = qSer[(size_t)t * N + T.rch2pos[r]];
    for (int r = 0; r < N; ++r) {
        const int p = T.rch2pos[r];
        vol_out[r] = vol1[p]; wb_out[r] = wb[p];
        evap_left[r] = (d.lakeEvap && slot[p] >= 0) ? lakeE[(size_t)(nSteps - 1) * nLake + slot[p]] : 0.0;
    }
    std::snprintf(msg, 256, "ok");
    return 0;
}
