// Single-threaded HOST build of the Euler routing schemes (mizuroute_b200/csrc/mr_euler.cuh: kinematic wave,
// Muskingum-Cunge, diffusive wave).
//
// TEST INFRASTRUCTURE ONLY.  The very source the GPU runs one thread per (reach, step) is executed here reach by reach
// in stage order and compared bit for bit with the CPU oracle (tests/test_euler_emul.py): that pins the restructured
// Thomas solve (coefficients by row range, molecule stored node-major, low-storage reduction applied while
// back-substituting) without a GPU.  Not a CPU fallback: nothing under mizuroute_b200/ builds, loads or links this file.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../mizuroute_b200/csrc/mr_euler.cuh"
#include "../../mizuroute_b200/csrc/mr_topo.h"
#include "da_emul.h"

using namespace mr;

static DaEmul g_da;
// gauge observations of the NEXT euler_emul_run: obs [nSteps][nRch] caller order, hasRecord [nSteps] or NULL; qerr_out [nRch]
extern "C" void euler_emul_set_da(int blend, int trend, const int *hasRecord, const double *obs, double *qerr_out) {
    g_da.blend = blend; g_da.trend = trend; g_da.hasRecord = hasRecord; g_da.obs = obs; g_da.qerrOut = qerr_out;
}

template <int M>
static void run_method(DevNet &d, const Topology &T, int nSteps) {
    const bool ext = d.wmFlux != nullptr || d.daQobs != nullptr;   // water management / data assimilation: the EXT instantiations, as route_device picks them
    for (int t = 0; t < nSteps; ++t)
        for (int p = 0; p < d.nRch; ++p) {                  // stage order: upstream before downstream
            if (ext) { if constexpr (M == M_MC) mc_reach<true>(d, p, t); else kw_dw_reach<M, true>(d, p, t); }
            else { if constexpr (M == M_MC) mc_reach<false>(d, p, t); else kw_dw_reach<M, false>(d, p, t); }
        }
}

extern "C" int euler_emul_run(int method, int nRch, int nHRU, const int *segId, const int *downSegId, const int *hruSegId, const double *hruArea,
                              const double *length, const double *slope, double mann_n, double wscale, double dt, int hw_drain_point,
                              double min_length_route, int floodplain, int nSteps,
                              const double *qr /* [nSteps+1][nRch] BASIN_QR(1) before step 0 and after every step, caller order */,
                              const double *wm_flux /* [nSteps][nRch] caller order, or NULL */,
                              double *q_out /* [nSteps][nRch] REACH_Q */, double *vol_out /* [nRch] REACH_VOL(1) */,
                              double *mol_out /* [nRch][n_molecule] */, char *msg) {
    Topology T;
    std::string terr;
    if (build_topology(nRch, nHRU, segId, downSegId, hruSegId, hruArea, T, terr)) { std::snprintf(msg, 256, "%s", terr.c_str()); return 1; }
    if (method < M_KW || method > M_DW) { std::snprintf(msg, 256, "method must be 3, 4 or 5"); return 1; }
    const int N = nRch, nm = n_molecule(method);
    std::vector<double> rlen(N), rslp(N), rwid(N), rman(N, mann_n), rdep(N), zc(N, 0.0), zf(N, 1000.0), rstor(N);
    for (int p = 0; p < N; ++p) {
        const int r = T.pos2rch[p];
        rlen[p] = length[r]; rslp[p] = std::fmax(slope[r], 1.e-6); rwid[p] = wscale * std::sqrt(T.totArea[p]);
        rdep[p] = floodplain ? (double)0.000045f * std::sqrt(T.totArea[p]) : 100000.0;                  // as mr_set_network
        rstor[p] = rdep[p] * (rwid[p] + zc[p] * rdep[p]) * rlen[p];
    }
    std::vector<double> qrSer((size_t)(nSteps + 1) * N), qSer((size_t)nSteps * N, 0.0), inflow(N, 0.0), vol0(N, 0.0), vol1(N, 0.0), wb(N, 0.0),
        mol((size_t)nm * N, 0.0), flood(N, 0.0), ele(N, 0.0);
    for (int t = 0; t <= nSteps; ++t) for (int r = 0; r < N; ++r) qrSer[(size_t)t * N + T.rch2pos[r]] = qr[(size_t)t * N + r];
    DevNet d{};
    d.nRch = N; d.nHRU = nHRU; d.nStage = T.nStage; d.nHead = T.nHead; d.dt = dt; d.hwDrain = hw_drain_point; d.minLengthRoute = min_length_route;
    d.stageOf = T.stageOf.data(); d.upPtr = T.upPtr.data(); d.upIdx = T.upIdx.data(); d.nGood = T.nGood.data();
    d.rlength = rlen.data(); d.rslope = rslp.data(); d.rwidth = rwid.data(); d.rmann = rman.data();
    d.rdepth = rdep.data(); d.sideSlope = zc.data(); d.fldpSlope = zf.data(); d.rstorage = rstor.data();
    d.qrSer = qrSer.data(); d.qSer[method] = qSer.data(); d.inflow[method] = inflow.data(); d.vol0[method] = vol0.data(); d.vol1[method] = vol1.data();
    d.wb[method] = wb.data(); d.mol[method] = mol.data(); d.floodVol[method] = flood.data(); d.reachEle[method] = ele.data();
    std::vector<double> fS;
    if (wm_flux) {                                          // stage order, as mr_upload_wm
        fS.resize((size_t)nSteps * N);
        for (int t = 0; t < nSteps; ++t) for (int p = 0; p < N; ++p) fS[(size_t)t * N + p] = wm_flux[(size_t)t * N + T.pos2rch[p]];
        d.wmFlux = fS.data();
    }
    g_da.attach(d, T, method, nSteps);
    if (method == M_KW) run_method<M_KW>(d, T, nSteps); else if (method == M_MC) run_method<M_MC>(d, T, nSteps); else run_method<M_DW>(d, T, nSteps);
    for (int t = 0; t < nSteps; ++t) for (int r = 0; r < N; ++r) q_out[(size_t)t * N + r] = qSer[(size_t)t * N + T.rch2pos[r]];
    for (int r = 0; r < N; ++r) {
        const int p = T.rch2pos[r];
        vol_out[r] = vol1[p];
        for (int k = 0; k < nm; ++k) mol_out[(size_t)r * nm + k] = mol[(size_t)k * N + p];
    }
    g_da.finish(T);
    std::snprintf(msg, 256, "ok");
    return 0;
}
