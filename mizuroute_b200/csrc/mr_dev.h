// Device image of one routing domain (DevNet), constants and error plumbing shared by the kernels
// (mr_kernels.cuh), the warp-cooperative KWT code (mr_kwt.cuh) and its single-lane host build (tests/emul).
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include "../../include/mizuroute_b200.h"
#include "mr_lanes.h"

namespace mr {

constexpr int KWS = MR_KW_SLOTS;      // particle slots per reach per buffer (restart schema)
constexpr int KWP = 24;               // pitch of a reach's particle row in HBM: 192 B = 6 aligned 32-B sectors
constexpr int WCAP = 160;             // per-warp particle scratch (own + merged upstream), see kwt_reach
constexpr int MAXSER = 32;            // series merged at one confluence (basins + non-headwater reaches)
constexpr int POOL = 208;             // staged upstream series points (<= WCAP + MAXSER + MAXSER/2)
constexpr int WCAP_S = 64;            // the same for the shared-memory fast path (tasks that need more are re-run
constexpr int POOL_S = 88;            // with the full-capacity scratch in the global arena)
constexpr int NKIN = MR_MAXQPAR + 2;  // kinwav work arrays (1-based, <= 19 particles routed)

enum { FLAG_LAKE = 1, FLAG_LAKE_UP = 2, FLAG_GHOST = 4 };
// Everything static the lane-per-task KWT code (mr_kwt_scalar.cuh) needs to know about an interior reach, in one record of
// 96 B (written once per network, k_kws_records): one load level instead of the chain upPtr -> upIdx -> nGood / R_WIDTH.
constexpr int KWS_BMAX = 3;          // upstream reaches (basin series; at most as many wave series) of that code
struct KwsRec {
    int stage, cls, nGood, nUps;     // cls != 0: not a plain reach (lake, lake outlet, ghost, more than KWS_BMAX upstream reaches)
    int U[KWS_BMAX], isr;            // upstream positions; bit s: upstream s has contributing area (a wave series)
    double W, scfB, aK, XMX;         // R_WIDTH, 1 / R_WIDTH, ALFA * K**(1/ALFA), RLENGTH
    double scf[KWS_BMAX];            // R_WIDTH(upstream s) / R_WIDTH
    double pad_;
};
// HYPE reservoir parameters of one lake (dataTypes.f90:202-213; integers / logicals as 0/1 doubles), in this order
struct HypeParams { double E_emr, E_lim, E_min, E_zero, Qrate_emr, Erate_emr, Qrate_prim, Qrate_amp, Qrate_phs, prim_F, A_avg, Qsim_mode; };
constexpr int HYP_COUNT = 12;
// Hanasaki-2006 reservoir of one lake (dataTypes.f90:215-254): parameters, the monthly mean inflows / release coefficient the
// model itself rewrites, and its inflow memory QPASTUP_IRF(12, L31) kept as one ring per month (newest value at `head`).
// All of it is per REACH in the reference (RPARAM / RCHFLX), i.e. shared by the routing methods of a step.
struct H06Lake {
    double Smax, alpha, envfact, c1, c2, exponent, denominator, c_compare, frac_Sdead, E_rel_ini;
    double I[12], D[12];
    int purpose, memF, L31, L30, LF, LFnoleap, filled;      // LF / LFnoleap: February row length in the standard / noleap calendar
    int head[12];
    long long memOff;              // offset of this lake's [12][L31] block in DevNet::h06Mem
};
enum { M_SUM = 0, M_IRF = 1, M_KWT = 2, M_KW = 3, M_MC = 4, M_DW = 5, N_METHODS = 6 };   // = digits of <route_opt>, public_var.f90:74-80
// computational molecules of the Euler schemes (init_model_data.f90:386-393): KW 20, MC 2, DW 20 nodes per reach
constexpr int n_molecule(int m) { return m == M_KW || m == M_DW ? 20 : (m == M_MC ? 2 : 0); }
template <int M> constexpr int NMOL = (M == M_KW || M == M_DW) ? 20 : (M == M_MC ? 2 : 0);      // the same, usable in device code

struct DevNet {
    int nRch, nHRU, nStage, ntdhBas, maxtdh;
    // options
    double dt, runoffMin, tconv, lconv, minLengthRoute;
    int doesBasinRoute, hwDrain, isLakeSim, lakeInputOption;
    // topology / parameters
    int nHead;
    const int *stageOf, *upPtr, *upIdx, *nGood, *hruPtr, *hruIdx, *flags, *ntdh, *lakeType;
    const double *hruWgt, *basArea, *rlength, *rslope, *rwidth, *rmann;
    const double *uh, *fracFuture;
    const double *kwK, *kwAK;       // KWT: sqrt(R_SLOPE)/R_MAN_N and ALFA*K**(1/ALFA) per reach (kwt_route.f90:1283-1296)
    const KwsRec *kwRec;            // KWT: static per-reach record of the lane-per-task code
    const double *d03MaxS, *d03Coef, *d03Pow, *d03S0;
    // lake forcing (optional): HRU-level evaporation / precipitation of the batch [K][nHRU] and their reach-level values for
    // the lake reaches [kmax][nLake]; lakeEvap == nullptr = no lake forcing (exact zeros in lake_route)
    const int *lakeSlot; int nLake;
    // HYPE reservoirs (lakeModelType 3): parameters by lake slot [nLake] and the day of year of every step of the batch [kmax]
    // (nullptr = no simulation start datetime was given)
    const struct HypeParams *hyp;
    struct H06Lake *h06;            // Hanasaki-2006 reservoirs by lake slot (nullptr = none in this domain)
    double *h06Mem;
    const int *stepMonth, *stepDay; // month / day of month of every step of the batch [kmax], with stepDoy
    // water management of the batch (mr_upload_wm; nullptr = off): abstraction (+) / injection (-) and target lake volume per
    // reach and step [kmax][nRch] in stage order, -9999 = none; lakes that follow the target volume; volume jump start
    const double *wmFlux, *wmVol; const unsigned char *lakeTargVol; int volJumpStart;
    int noleap;                     // calendar given with mr_set_sim_start
    int lastK;                      // steps of the previous batch (its last REACH_Q row is still in qSer)
    const int *stepDoy;
    const double *evapo, *precip;
    double *lakeEvap, *lakePrecip;
    const double *rdepth, *sideSlope, *fldpSlope, *rstorage;   // Euler schemes: bankfull depth, side / floodplain slopes, bankfull storage
    // forcing and per-step times
    const double *runoff, *T0s, *T1s;
    // state and fluxes
    double *qfutBas, *qrSer, *basinQI;
    double *qSer[N_METHODS], *vol0[N_METHODS], *vol1[N_METHODS], *inflow[N_METHODS], *wb[N_METHODS];
    double *mol[N_METHODS];         // molecule%Q of KW / MC / DW, node-major [n_molecule][nRch]
    double *floodVol[N_METHODS], *reachEle[N_METHODS];
    double *qfutIrf;
    int *kwN[2], *kwNR[2];
    double *kwQF[2], *kwTI[2], *kwTR[2];
    // multi-domain hand-off: per-step records of exported outlets / imported ghosts, [slot][kmax][recLen]
    const int *expSlot, *impSlot;   // by position, -1 = none (nullptr = feature off)
    double *expBuf; const double *impBuf;
    int recLen, kmax, nRoutes, routeSlot[N_METHODS];
    void *kwArena; unsigned long long *kwArenaMask;   // full-capacity KWT scratch: 64 slots per SM + their busy bits
    int *err;                       // [0] code (0 = ok) [1] position [2] site
    unsigned *kwCount;              // optional per-reach count of particles read+written (nullptr = off)
    unsigned long long *kwProf;     // optional [8][2] cycles / tasks per task class (development profile, nullptr = off)
    // (last, so that the parameter offsets the other kernels read stay where they were)
    // data assimilation by direct insertion (mr_set_da / mr_upload_obs; daQobs == nullptr = off): RCHFLX%Qobs and %Qelapsed as
    // every step of the batch sees them [kmax][nRch] in stage order (written by k_da_rows), and ROUTE(:)%Qerror per method
    const double *daQobs; const int *daElapsed; double *qerr[N_METHODS]; int qBlendPeriod, qErrTrend;
};

// site ids for error messages (decoded in mr_lib.cu)
enum {
    E_NEG_RUNOFF = 1, E_LAKE_UPS = 2, E_NEG_FLOW = 3, E_SCRATCH = 4, E_STUCK = 5, E_TIME_ORDER = 6, E_BRACKET = 7,
    E_QD_BOUNDS = 8, E_ZERO_FLOW = 9, E_TEXIT2 = 10, E_RUPDATE = 11, E_NO_NONROUTED = 12, E_INTERP = 13,
    E_LAKE_TYPE = 14, E_TOO_MANY_UPS = 15, E_THIN = 16, E_NO_ROUTED_UP = 17, E_LAKE_PARAM = 18, E_NO_CALENDAR = 19
};

MR_DEV_NOINLINE void raise(int *err, int code, int p, int site) {
#if defined(__CUDACC__)
    if (atomicCAS(&err[0], 0, code) == 0) { err[1] = p; err[2] = site; }
#else
    if (err[0] == 0) { err[0] = code; err[1] = p; err[2] = site; }
#endif
}

// Water abstraction (+) / injection (-) of a river reach (irf_route.f90:114-142 = kwe_route.f90:118-146 = mc_route.f90:118-146 =
// dfw_route.f90:122-150): taken from the storage first, then from the upstream inflow, then from the lateral flow.  Returns
// REACH_WM_FLUX_actual (it starts as the demand -- the missing value included, as in the reference).
MR_DEV double wm_cascade(double want, double dt, double &v1, double &qup, double &qlat) {
    double actual = want;
    if (want == -9999.0) return actual;                // realMissing: no water management at this reach
    double Qabs = want;
    if (Qabs > 0) {
        if (v1 / dt > Qabs) {
            v1 = v1 - Qabs * dt;
        } else {
            Qabs = Qabs - v1 / dt;
            v1 = 0.0;
            if (qup > Qabs) {
                qup = qup - Qabs;
            } else {
                Qabs = Qabs - qup;
                qup = 0.0;
                if (qlat > Qabs) {
                    qlat = qlat - Qabs;
                } else {
                    Qabs = Qabs - qlat;
                    qlat = 0.0;
                    actual = want - Qabs;
                }
            }
        }
    } else {
        qlat = qlat - Qabs;
    }
    return actual;
}

// Gauge observations of a batch -> what direct_insertion sees at every step (main_route.f90:125-148).  obs [K][N] holds the
// gauge values of the steps (NaN / negative = none at that reach) and is overwritten by RCHFLX%Qobs; el [K][N] receives
// RCHFLX%Qelapsed; hasRecord[t] = 0: the gauge file has no record at step t (every reach's Qelapsed goes up by one).  The
// running Qobs / Qelapsed of the reach (qobsState, elState) carry over to the next batch.  One thread per reach.
MR_DEV void da_rows(double *obs, int *el, double *qobsState, int *elState, const unsigned char *hasRecord, int N, int p, int K) {
    double qobs = qobsState[p];
    int e = elState[p];
    for (int t = 0; t < K; ++t) {
        if (hasRecord[t]) {
            const double v = obs[(size_t)t * N + p];
            if (!((v != v) || (v < 0))) { qobs = v; e = 0; }
        } else e = e + 1;
        obs[(size_t)t * N + p] = qobs; el[(size_t)t * N + p] = e;
    }
    qobsState[p] = qobs; elState[p] = e;
}

// direct_insertion (data_assimilation.f90:23-97) of method M at (reach p, step t): returns the corrected REACH_Q.  The trend
// model (1 constant, 2 linear, 3 logistic, 4 exponential) is validated by mr_set_da.
MR_DEV_NOINLINE double direct_insertion(const DevNet &d, int M, int p, int t, double q) {
    const size_t i = (size_t)t * d.nRch + p;
    const double qobs = d.daQobs[i];
    const int el = d.daElapsed[i], blend = d.qBlendPeriod;
    double qerror = d.qerr[M][p], qcorrect = 0.0;
    if (qobs > 0.0) qerror = q - qobs;
    if (el > blend) qerror = 0.0;
    if (el <= blend) {
        if (d.qErrTrend == 1) qcorrect = qerror;
        else if (d.qErrTrend == 2) qcorrect = qerror * (1.0 - (double)el / (double)blend);
        else if (d.qErrTrend == 3) {
            const double x0 = 0.25, y0 = (double)0.90f;            // single-precision literals, data_assimilation.f90:76
            const double k = log(1.0 / y0 - 1.0) / (blend / 2.0 - blend * x0);
            qcorrect = qerror / (1.0 + exp(-k * (1.0 * el - blend / 2.0)));
        } else if (qerror != 0.0) {
            const double k = log(0.1 / fabs(qerror)) / (1.0 * blend);
            qcorrect = qerror * exp(k * el);
        }
    }
    d.qerr[M][p] = qerror;
    return fmax(q - qcorrect, 0.0);
}

// one out-of-line copy of pow(): its inlined body is ~250 instructions per call site
#if defined(MR_TEST_POW_NOISE) && !defined(__CUDACC__)
// Host test builds only (tests/test_kwt_conditioning.py, tests/test_euler_emul.py): pow() perturbed by -1/0/+1 ulp, pseudo-randomly, to measure how
// a case reacts to the last-ulp differences between two correct pow() implementations (libm vs the device).
inline double mr_pow(double x, double y) {
    const double r = pow(x, y);
    unsigned long long u; memcpy(&u, &x, 8);
    u = (u * 0x9E3779B97F4A7C15ull + (unsigned long long)(MR_TEST_POW_NOISE)) >> 61;
    return u < 3 ? nextafter(r, 1e300) : (u < 6 ? nextafter(r, -1e300) : r);
}
#else
MR_DEV_NOINLINE double mr_pow(double x, double y) { return pow(x, y); }
#endif

// x**p2 with p2 = (ALFA-1)/ALFA as the reference evaluates it in double precision (kwt_route.f90:1296; 0.4 + 2.22e-17): the one
// pow() of the KWT inner loop, once per wave particle and step.  x**(2/5) = fifth root of x*x: the root of the mantissa
// (scaled to [0.5, 16)) starts from a single-precision estimate and takes two Newton steps in double precision, the second on
// a residual carried in two doubles that also holds what x*x lost to rounding; the 2.22e-17 the double exponent lies above 2/5 enters the same final
// correction as the factor 1 + 2.22e-17 ln(x).  The result is within 0.51 ulp of the exact power -- tighter than the 2 ulp of the CUDA pow() it
// replaces, at a sixth of its instructions.  Outside [2**-255, 2**257) -- zero, negative, infinite, NaN included -- pow() answers.
// Host builds of the device code (tests/emul) keep calling mr_pow(): they are compared with the oracle's libm bit for bit,
// which no second implementation of pow can promise; tests/test_fastpow.py measures this routine against extended precision.
inline
#if defined(__CUDACC__)
__host__ __device__
#endif
double mr_pow04_fast(double x) {
    unsigned long long ux; memcpy(&ux, &x, 8);
    const int hx = (int)(ux >> 32);
    if (hx < 0x30000000 || hx >= 0x50000000) return pow(x, ((5.0 / 3.0) - 1.0) / (5.0 / 3.0));
    const double y = x * x;
    unsigned long long uy; memcpy(&uy, &y, 8);
    const int e = (int)((uy >> 52) & 0x7ff) - 1022;            // y = m * 2**e, m in [0.5, 1)
    const int k = (e + 1280) / 5 - 256, j = e - 5 * k;         // e = 5k + j, j in 0..4
    const unsigned long long um = (uy & 0x000fffffffffffffull) | ((unsigned long long)(1022 + j) << 52);
    double mj; memcpy(&mj, &um, 8);                            // m * 2**j in [0.5, 16)
#if defined(__CUDA_ARCH__)
    const float lg = __log2f((float)mj);
    const float g = exp2f(0.2f * lg);
    const float g2 = g * g;
    const double c = (double)__frcp_rn(5.0f * g2 * g2);        // ~ 1 / (5 r**4)
#else
    const float lg = log2f((float)mj);
    const float g = exp2f(0.2f * lg);
    const float g2 = g * g;
    const double c = (double)(1.0f / (5.0f * g2 * g2));
#endif
    double r = (double)g;
    {                                                          // Newton step on r**5 = mj, derivative from the estimate
        const double r2 = r * r, r4 = r2 * r2, r5 = r4 * r;
        r = fma(mj - r5, c, r);
    }
    {                                                          // the same with r**5 in two doubles (r5 + t5), one final rounding
        const double r2 = r * r, e2 = fma(r, r, -r2);
        const double r4 = r2 * r2, e4 = fma(r2, r2, -r4);
        const double t4 = fma(2.0 * r2, e2, e4);
        const double r5 = r4 * r, e5 = fma(r4, r, -r5);
        const double t5 = fma(t4, r, e5);
        // what x*x lost to rounding, scaled like mj (exact: a power of two)
        const unsigned long long ui = (unsigned long long)(1023 - 5 * k) << 52;
        double inv; memcpy(&inv, &ui, 8);
        const double ey = fma(x, x, -y) * inv;
        // p2 - 2/5 = 0.4 * 2**-54; ln(x) = ln(2)/2 * log2(x*x)
        const double dl = 2.2204460492503131e-17 * (0.34657359027997264 * ((double)(5 * k) + (double)lg));
        r = r + fma(((mj - r5) - t5) + ey, c, r * dl);
    }
    const unsigned long long us = (unsigned long long)(1023 + k) << 52;
    double sc; memcpy(&sc, &us, 8);
    return r * sc;
}
#if defined(__CUDACC__)
MR_DEV double mr_pow04(double x) { return mr_pow04_fast(x); }
#else
inline double mr_pow04(double x) { return mr_pow(x, ((5.0 / 3.0) - 1.0) / (5.0 / 3.0)); }
#endif

}  // namespace mr
