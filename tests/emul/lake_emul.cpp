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
#include "../../mizuroute_b200/csrc/mr_lakeparams.h"
#include <map>

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
    std::vector<HypeParams> hypBySlot;
    std::vector<int> doy(nSteps, 0);
    if (hyp && nLake) {                                      // mr_set_network: the same builder
        static const char *names[HYP_COUNT] = {"HYP_E_emr", "HYP_E_lim", "HYP_E_min", "HYP_E_zero", "HYP_Qrate_emr", "HYP_Erate_emr",
                                               "HYP_Qrate_prim", "HYP_Qrate_amp", "HYP_Qrate_phs", "HYP_prim_F", "HYP_A_avg", "HYP_Qsim_mode"};
        std::map<std::string, std::vector<double>> given;
        for (int k = 0; k < HYP_COUNT; ++k) given[names[k]].assign(hyp + (size_t)k * N, hyp + (size_t)(k + 1) * N);
        const LakeParamLookup par = [&](const std::string &nm) -> const std::vector<double> * { auto it = given.find(nm); return it == given.end() ? nullptr : &it->second; };
        const std::string missing = build_hype_params(nLake, pos.data(), T.pos2rch.data(), par, hypBySlot);
        if (!missing.empty()) { std::snprintf(msg, 256, "missing %s", missing.c_str()); return 20; }
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

// Two Euler methods (kinematic wave + diffusive wave) routed the way route_device orders them when lake state is shared by
// the methods (one stream): per batch of K steps the headwater reaches first (k_headwater, method by method), then wavefront
// by wavefront and, inside a wavefront, method by method.  With Hanasaki-2006 reservoirs (H06Lake built as mr_set_network
// builds it) this must reproduce the oracle, which -- like the reference -- steps method by method inside every time step.
template <int M>
static void emul_reach(DevNet &d, const std::vector<int> &flags, int p, int t, long long tau) {
    if (flags[p] & FLAG_LAKE) lake_reach<M, true>(d, p, t, tau); else kw_dw_reach<M>(d, p, t);
}

extern "C" int lake_emul_run_h06(int nRch, int nHRU, const int *segId, const int *downSegId, const int *hruSegId, const double *hruArea,
                                 const double *length, const double *slope, const int *islake, const int *lakeType, const double *maxS,
                                 const double *coef, const double *pw, const double *s0, double mann_n, double wscale, double dt,
                                 int lakeInputOption, int nSteps, int K,
                                 const double *qr /* [nSteps+1][nRch] BASIN_QR(1), caller order */,
                                 const double *h06 /* [38][nRch]: 10 scalars (H06Lake order), I_Jan..Dec, D_Jan..Dec, purpose, I_mem_F, I_mem_L, 0 */,
                                 int startY, int startM, int startD, double startSec, int noleap,
                                 double *q_out /* [2][nSteps][nRch] */, double *vol_out /* [2][nRch] */, char *msg) {
    Topology T;
    std::string terr;
    if (build_topology(nRch, nHRU, segId, downSegId, hruSegId, hruArea, T, terr)) { std::snprintf(msg, 256, "%s", terr.c_str()); return 1; }
    const int N = nRch, MM[2] = {M_KW, M_DW};
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
    std::vector<H06Lake> lk;
    long long off = 0;
    {                                                        // mr_set_network: the same builder, fed by name
        static const char *mon[12] = {"Jan", "Feb", "Mar", "Apr", "May", "Jun", "Jul", "Aug", "Sep", "Oct", "Nov", "Dec"};
        std::vector<std::string> names = {"H06_Smax", "H06_alpha", "H06_envfact", "H06_c1", "H06_c2", "H06_exponent", "H06_denominator",
                                          "H06_c_compare", "H06_frac_Sdead", "H06_E_rel_ini"};
        for (int k = 0; k < 12; ++k) names.push_back(std::string("H06_I_") + mon[k]);
        for (int k = 0; k < 12; ++k) names.push_back(std::string("H06_D_") + mon[k]);
        names.push_back("H06_purpose"); names.push_back("H06_I_mem_F"); names.push_back("H06_I_mem_L");
        std::map<std::string, std::vector<double>> given;
        for (size_t k = 0; k < names.size(); ++k) given[names[k]].assign(h06 + k * N, h06 + (k + 1) * N);
        const LakeParamLookup par = [&](const std::string &nm) -> const std::vector<double> * { auto it = given.find(nm); return it == given.end() ? nullptr : &it->second; };
        const std::string missing = build_h06_lakes(nLake, pos.data(), T.pos2rch.data(), ltype.data(), dt, par, lk, off);
        if (!missing.empty()) { std::snprintf(msg, 256, "missing %s", missing.c_str()); return 20; }
    }
    std::vector<double> mem((size_t)(off > 0 ? off : 1), 0.0);
    std::vector<double> qrAll((size_t)(nSteps + 1) * N);
    for (int t = 0; t <= nSteps; ++t) for (int r = 0; r < N; ++r) qrAll[(size_t)t * N + T.rch2pos[r]] = qr[(size_t)t * N + r];
    std::vector<double> qrSer((size_t)(K + 1) * N), qSer[2], inflow[2], vol0[2], vol1[2], wb[2], mol[2], flood[2], ele[2];
    int err[4] = {0, 0, 0, 0};
    DevNet d{};
    d.nRch = N; d.nHRU = nHRU; d.nStage = T.nStage; d.nHead = T.nHead; d.dt = dt; d.hwDrain = 2; d.minLengthRoute = 0.0;
    d.lakeInputOption = lakeInputOption; d.isLakeSim = 1; d.noleap = noleap;
    d.stageOf = T.stageOf.data(); d.upPtr = T.upPtr.data(); d.upIdx = T.upIdx.data(); d.nGood = T.nGood.data(); d.flags = flags.data();
    d.hruPtr = T.hruPtr.data(); d.hruIdx = T.hruIdx.data(); d.hruWgt = T.hruWgt.data(); d.basArea = T.basArea.data();
    d.rlength = rlen.data(); d.rslope = rslp.data(); d.rwidth = rwid.data(); d.rmann = rman.data();
    d.rdepth = rdep.data(); d.sideSlope = zc.data(); d.fldpSlope = zf.data(); d.rstorage = rstor.data();
    d.lakeType = ltype.data(); d.d03MaxS = dMaxS.data(); d.d03Coef = dCoef.data(); d.d03Pow = dPw.data(); d.d03S0 = dS0.data();
    d.qrSer = qrSer.data(); d.err = err; d.lakeSlot = slot.data(); d.nLake = nLake; d.h06 = lk.data(); d.h06Mem = mem.data();
    for (int i = 0; i < 2; ++i) {
        const int m = MM[i], nm = n_molecule(m);
        qSer[i].assign((size_t)K * N, 0.0); inflow[i].assign(N, 0.0); vol0[i].assign(N, 0.0); vol1[i].assign(N, 0.0); wb[i].assign(N, 0.0);
        mol[i].assign((size_t)nm * N, 0.0); flood[i].assign(N, 0.0); ele[i].assign(N, 0.0);
        d.qSer[m] = qSer[i].data(); d.inflow[m] = inflow[i].data(); d.vol0[m] = vol0[i].data(); d.vol1[m] = vol1[i].data(); d.wb[m] = wb[i].data();
        d.mol[m] = mol[i].data(); d.floodVol[m] = flood[i].data(); d.reachEle[m] = ele[i].data();
    }
    std::vector<int> doy(K), mon(K), dom(K);
    d.stepDoy = doy.data(); d.stepMonth = mon.data(); d.stepDay = dom.data();
    int lastK = 0;
    for (int s0 = 0; s0 < nSteps; s0 += K) {
        const int kb = nSteps - s0 < K ? nSteps - s0 : K;
        d.lastK = lastK;
        for (int t = 0; t <= kb; ++t) std::memcpy(&qrSer[(size_t)t * N], &qrAll[(size_t)(s0 + t) * N], sizeof(double) * N);
        for (int t = 0; t < kb; ++t) step_calendar(startY, startM, startD, startSec, noleap != 0, dt, s0 + t, mon[t], dom[t], doy[t]);
        for (int i = 0; i < 2; ++i)                          // k_headwater<M>
            for (int p = 0; p < T.nHead; ++p)
                for (int t = 0; t < kb; ++t) { if (MM[i] == M_KW) emul_reach<M_KW>(d, flags, p, t, s0 + t); else emul_reach<M_DW>(d, flags, p, t, s0 + t); }
        for (int w = 0; w < T.nStage + kb; ++w)              // wavefronts: (reach, step) with stage + step == w
            for (int i = 0; i < 2; ++i)
                for (int p = T.nHead; p < N; ++p) {
                    const int t = w - T.stageOf[p];
                    if (t < 0 || t >= kb) continue;
                    if (MM[i] == M_KW) emul_reach<M_KW>(d, flags, p, t, s0 + t); else emul_reach<M_DW>(d, flags, p, t, s0 + t);
                    if (err[0]) { std::snprintf(msg, 256, "ierr %d at position %d site %d", err[0], err[1], err[2]); return err[0]; }
                }
        for (int i = 0; i < 2; ++i)
            for (int t = 0; t < kb; ++t) for (int r = 0; r < N; ++r)
                q_out[((size_t)i * nSteps + s0 + t) * N + r] = qSer[i][(size_t)t * N + T.rch2pos[r]];
        lastK = kb;
    }
    for (int i = 0; i < 2; ++i) for (int r = 0; r < N; ++r) vol_out[(size_t)i * N + r] = vol1[i][T.rch2pos[r]];
    std::snprintf(msg, 256, "ok");
    return 0;
}
