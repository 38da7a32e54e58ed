// sm_100a kernels of the reach-routing path.  Compiled with --fmad=false: every a*b+c below is two
// IEEE roundings, as in the reference built without FMA contraction, so SUM / IRF / hillslope-UH
// results are bit-identical to a scalar CPU evaluation in the same operation order.
//
// Layout (all arrays in "stage order", see mr_topo.h; N = nRch):
//   per-reach scalars      x[p]
//   windows / particles    x[k*N + p]            (slot-major: thread-per-reach accesses are coalesced)
//   per-step series        x[t*N + p]
//
// Kernels
//   k_basin       K1 basin2reach (process_remap.f90:372-420) fused with K2 hillslope UH
//                 (basinUH.f90:94-176) for all steps of a batch; the UH window is a ring (no shift copy),
//                 staged in shared memory when it fits so a batch reads/writes it once
//   k_route<M>    one time-skewed wavefront of route_network (main_route.f90:356-403):
//                 M=0 accum_inst_runoff (accum_runoff.f90:60-75), M=1 irf_rch+conv_upsbas_qr
//                 (irf_route.f90:82-150,235-262), M=2 kwt_rch and callees (kwt_route.f90:36-1622);
//                 lake reaches branch to lake_route (lake_route.f90:87-229)
#pragma once
#include <cfloat>
#include <cstdint>
#include "../../include/mizuroute_b200.h"

namespace mr {

constexpr int KWS = MR_KW_SLOTS;      // particle slots per reach per buffer
constexpr int WCAP = 160;             // per-thread particle scratch (own + merged upstream), see kwt_reach
constexpr int MAXSER = 32;            // series merged at one confluence (basins + non-headwater reaches)
constexpr int NKIN = MR_MAXQPAR + 2;  // kinwav work arrays (1-based, <= 19 particles routed)

enum { FLAG_LAKE = 1, FLAG_LAKE_UP = 2, FLAG_GHOST = 4 };
enum { M_SUM = 0, M_IRF = 1, M_KWT = 2 };

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
    const double *d03MaxS, *d03Coef, *d03Pow, *d03S0;
    // forcing and per-step times
    const double *runoff, *T0s, *T1s;
    // state and fluxes
    double *qfutBas, *qrSer, *basinQI;
    double *qSer[3], *vol0[3], *vol1[3], *inflow[3], *wb[3];
    double *qfutIrf;
    int *kwN[2], *kwNR[2];
    double *kwQF[2], *kwTI[2], *kwTR[2];
    int *err;                       // [0] code (0 = ok) [1] position [2] site
    unsigned *kwCount;              // optional per-reach count of particles read+written (nullptr = off)
};

// site ids for error messages (decoded in mr_lib.cu)
enum {
    E_NEG_RUNOFF = 1, E_LAKE_UPS = 2, E_NEG_FLOW = 3, E_SCRATCH = 4, E_STUCK = 5, E_TIME_ORDER = 6, E_BRACKET = 7,
    E_QD_BOUNDS = 8, E_ZERO_FLOW = 9, E_TEXIT2 = 10, E_RUPDATE = 11, E_NO_NONROUTED = 12, E_INTERP = 13,
    E_LAKE_TYPE = 14, E_TOO_MANY_UPS = 15, E_THIN = 16, E_NO_ROUTED_UP = 17
};

__device__ __forceinline__ void raise(int *err, int code, int p, int site) {
    if (atomicCAS(&err[0], 0, code) == 0) { err[1] = p; err[2] = site; }
}

// ------------------------------------------------------------------------------------------------
// K1 + K2
// ------------------------------------------------------------------------------------------------
// One thread per reach, all K steps of the batch, in chunks of BASIN_TC steps:
//   phase 1  basin2reach for the chunk's steps -> s_rr[t][thread]            (process_remap.f90:386-416)
//   phase 2  the UH window (a ring: logical slot k of step tau lives at physical slot (tau+k) mod nb) is
//            walked slot-group by slot-group; each physical slot is loaded ONCE per chunk, accumulates its
//            chunk's contributions  uh[k]*rr[t]  in a register in step order (the same additions, in the
//            same order, that irf_conv basinUH.f90:165-176 applies to that slot), emits BASIN_QR(1) when it
//            becomes logical slot 0, and is stored once.  HBM traffic per reach-step: 16*nb/BASIN_TC bytes
//            instead of the 16*nb of a per-step sweep.
constexpr int BASIN_TC = 64;     // steps per chunk (shared memory: BASIN_TC * blockDim * 8 B)
constexpr int BASIN_G = 8;       // slots in flight per thread (independent accumulation chains)
constexpr int BASIN_TPB = 128;

__global__ void __launch_bounds__(BASIN_TPB) k_basin(DevNet d, int K, long long tau0) {
    extern __shared__ double s_dyn[];
    double (*s_rr)[BASIN_TPB] = reinterpret_cast<double (*)[BASIN_TPB]>(s_dyn);   // [BASIN_TC][BASIN_TPB] reach runoff
    double *s_uh = s_dyn + BASIN_TC * BASIN_TPB;                                  // [nb] hillslope UH
    const int N = d.nRch, nb = d.ntdhBas;
    const int p = blockIdx.x * BASIN_TPB + threadIdx.x;
    for (int k = threadIdx.x; k < nb; k += BASIN_TPB) s_uh[k] = d.fracFuture[k];
    const bool live = p < N && !(p < N && (d.flags[p] & FLAG_GHOST));
    int h0 = 0, h1 = 0; double area = 0.0; bool lake = false;
    if (live) { h0 = d.hruPtr[p]; h1 = d.hruPtr[p + 1]; area = d.basArea[p]; lake = (d.flags[p] & FLAG_LAKE) != 0; }
    double rr = 0.0;
    for (int c0 = 0; c0 < K; c0 += BASIN_TC) {
        const int nc = (K - c0 < BASIN_TC) ? K - c0 : BASIN_TC;
        __syncthreads();                          // s_uh ready / previous chunk's s_rr consumed
        if (live) {
            for (int t = 0; t < nc; ++t) {
                if (h1 > h0) {
                    double r = 0.0;
                    for (int m = h0; m < h1; ++m) {
                        const double ro = d.runoff[(size_t)(c0 + t) * d.nHRU + d.hruIdx[m]];
                        if (ro < -1.e-3) raise(d.err, 20, p, E_NEG_RUNOFF);     // negRunoffTol, public_var.f90:31
                        r = r + d.hruWgt[m] * ro * d.tconv * d.lconv;
                    }
                    if (r < d.runoffMin) r = d.runoffMin;
                    rr = r * area;
                } else {
                    rr = d.runoffMin;
                }
                s_rr[t][threadIdx.x] = rr;
                if (d.doesBasinRoute != 1) d.qrSer[(size_t)(c0 + t + 1) * N + p] = rr;   // main_route.f90:223-226
            }
        }
        if (d.doesBasinRoute != 1 || !live) continue;
        // phase 2.  Step t of the chunk has ring head (tau0+c0+t) mod nb; physical slot s is logical
        // k = (s - head) mod nb at that step, i.e. k decreases by one per step and wraps from 0 to nb-1.
        const int head0 = (int)((tau0 + c0) % nb);
        for (int s0 = 0; s0 < nb; s0 += BASIN_G) {
            double v[BASIN_G]; int k[BASIN_G];
#pragma unroll
            for (int g = 0; g < BASIN_G; ++g) {
                const int s = s0 + g;
                v[g] = (s < nb) ? d.qfutBas[(size_t)s * N + p] : 0.0;
                int kk = s - head0; if (kk < 0) kk += nb;
                k[g] = kk;
            }
            for (int t = 0; t < nc; ++t) {
                const double x = s_rr[t][threadIdx.x];
#pragma unroll
                for (int g = 0; g < BASIN_G; ++g) {
                    if (s0 + g < nb) {
                        const double u = lake ? (k[g] == 0 ? 1.0 : 0.0) : s_uh[k[g]];      // basinUH.f90:113-116
                        v[g] = v[g] + u * x;
                        if (k[g] == 0) {          // this slot is BASIN_QR(1) of step t; it re-enters as slot nb-1 = 0
                            d.qrSer[(size_t)(c0 + t + 1) * N + p] = v[g];
                            v[g] = 0.0;
                            k[g] = nb;
                        }
                        k[g] -= 1;
                    }
                }
            }
#pragma unroll
            for (int g = 0; g < BASIN_G; ++g) if (s0 + g < nb) d.qfutBas[(size_t)(s0 + g) * N + p] = v[g];
        }
    }
    if (live) d.basinQI[p] = rr;
}

// carry BASIN_QR(1) of the previous batch into row 0 of the series
__global__ void k_carry_qr(double *qrSer, int N, int Kprev) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < N && Kprev > 0) qrSer[p] = qrSer[(size_t)Kprev * N + p];
}

// ------------------------------------------------------------------------------------------------
// water balance, water_balance.f90:67-87 (no water management, no precipitation/evaporation forcing)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double reach_wb(double v1, double v0, double qup, double qlat, double q, double dt) {
    const double dVol = v1 - v0;
    const double Qin = qup * dt, Qlateral = qlat * dt, precip = 0.0, evapo = 0.0;
    const double Qout = -1.0 * q * dt;
    const double Qtake = -1.0 * 0.0 * dt;
    return dVol - (Qin + Qlateral + precip + Qtake + Qout + evapo);
}

// lake_route.f90:87-229,466-470 for the endorheic and Doll-2003 types
template <int M>
__device__ void lake_reach(const DevNet &d, int p, int t, long long tau) {
    const int N = d.nRch;
    double *Qs = d.qSer[M] + (size_t)t * N;
    const int u0 = d.upPtr[p], u1 = d.upPtr[p + 1];
    const double dt = d.dt;
    double qup = 0.0;
    for (int m = u0; m < u1; ++m) qup = qup + Qs[d.upIdx[m]];
    const int type = d.lakeType[p];
    double v1 = d.vol1[M][p];
    if (tau == 0) {                                    // iTime==1 cold start, lake_route.f90:139-157
        if (type == MR_LAKE_ENDORHEIC) v1 = d.d03S0[p];
        else if (type == MR_LAKE_DOLL03) v1 = d.d03MaxS[p];
        else { raise(d.err, 20, p, E_LAKE_TYPE); return; }
    }
    const double v0 = v1;
    const double qr1 = d.qrSer[(size_t)(t + 1) * N + p];
    v1 = v1 + qup * dt;
    if (d.lakeInputOption == 1 || d.lakeInputOption == 2) v1 = v1 + qr1 * dt;
    if (d.lakeInputOption == 0 || d.lakeInputOption == 2) {
        v1 = v1 + 0.0 * dt;                            // basinprecip = basinevapo = 0 (no such forcing on this path)
        if (v1 > 0.0 * dt) v1 = v1 - 0.0 * dt; else v1 = 0.0;
    }
    double q;
    if (type == MR_LAKE_ENDORHEIC) {
        q = 0.0;
    } else if (type == MR_LAKE_DOLL03) {
        const double s0 = d.d03S0[p];
        if ((v1 - s0) > 0) q = d.d03Coef[p] * (v1 - s0) * pow((v1 - s0) / (d.d03MaxS[p] - s0), d.d03Pow[p]);
        else q = 0;
        q = q / 86400.0;
        q = fmin(q, v1 / dt);
        v1 = v1 - q * dt;
    } else { raise(d.err, 20, p, E_LAKE_TYPE); return; }
    Qs[p] = q;
    d.vol0[M][p] = v0; d.vol1[M][p] = v1;
    d.wb[M][p] = reach_wb(v1, v0, qup, qr1, q, dt);
}

// ------------------------------------------------------------------------------------------------
// kwt_route.f90 pieces
// ------------------------------------------------------------------------------------------------
// interp_rch (kwt_route.f90:1444-1622) for a single output interval; TOLD/QOLD are 0-based here
__device__ int kwt_time_average(const double *TOLD, const double *QOLD, int NOLD, double T0, double T1, double &QNEW) {
    if (TOLD[0] > T0 || TOLD[NOLD - 1] < T1) return 1;
    int IBEG = 0, IEND = 0;
    for (int i = 1; i < NOLD; ++i) if (T0 <= TOLD[i]) { IBEG = i; break; }
    for (int i = 0; i < NOLD; ++i) if (T1 <= TOLD[i]) { IEND = i; break; }
    if (T1 < TOLD[IBEG]) {
        const double SLOPE = (QOLD[IBEG] - QOLD[IBEG - 1]) / (TOLD[IBEG] - TOLD[IBEG - 1]);
        const double QEST0 = SLOPE * (T0 - TOLD[IBEG - 1]) + QOLD[IBEG - 1];
        const double QEST1 = SLOPE * (T1 - TOLD[IBEG - 1]) + QOLD[IBEG - 1];
        QNEW = 0.5 * (QEST0 + QEST1);
        return 0;
    }
    double AREAB = 0.0, AREAE = 0.0, AREAM = 0.0;
    if (T0 < TOLD[IBEG]) {
        const double SLOPE = (QOLD[IBEG] - QOLD[IBEG - 1]) / (TOLD[IBEG] - TOLD[IBEG - 1]);
        const double QEST0 = SLOPE * (T0 - TOLD[IBEG - 1]) + QOLD[IBEG - 1];
        AREAB = (TOLD[IBEG] - T0) * 0.5 * (QEST0 + QOLD[IBEG]);
    }
    if (T1 < TOLD[IEND]) {
        const double SLOPE = (QOLD[IEND] - QOLD[IEND - 1]) / (TOLD[IEND] - TOLD[IEND - 1]);
        const double QEST1 = SLOPE * (T1 - TOLD[IEND - 1]) + QOLD[IEND - 1];
        AREAE = (T1 - TOLD[IEND - 1]) * 0.5 * (QOLD[IEND - 1] + QEST1);
    }
    if (IBEG < IEND) {
        for (int IMID = IBEG + 1; IMID <= IEND; ++IMID)
            if (IMID < IEND || (IMID == IEND && T1 == TOLD[IEND] && T0 < TOLD[IEND - 1]))
                AREAM = AREAM + (TOLD[IMID] - TOLD[IMID - 1]) * 0.5 * (QOLD[IMID - 1] + QOLD[IMID]);
    }
    QNEW = (AREAB + AREAE + AREAM) / (T1 - T0);
    return 0;
}

__device__ __forceinline__ double thin_err(const double *Q, const double *T, int a, int m, int b) {
    // |INTERP(T(m), Q(a), Q(b), T(a), T(b)) - Q(m)|, kwt_route.f90:1054,1062,1114-1121
    return fabs((Q[a] + ((Q[b] - Q[a]) / (T[b] - T[a])) * (T[m] - T[a])) - Q[m]);
}

// remove_rch (kwt_route.f90:999-1123): greedy removal of the particle whose linear interpolation error is
// smallest until MAXQPAR remain.  The reference re-packs index arrays each pass; a doubly linked list of
// survivors visits them in the same order, so "first minimum" picks the same particle.
__device__ __noinline__ int kwt_thin(double *Q, double *T, double *X, int &n) {
    unsigned char prv[WCAP], nxt[WCAP];
    double ERR[WCAP];
    const int last = n - 1;
    for (int i = 0; i < n; ++i) { prv[i] = (unsigned char)(i - 1); nxt[i] = (unsigned char)(i + 1); ERR[i] = DBL_MAX; }
    for (int i = 1; i < last; ++i) ERR[i] = thin_err(Q, T, i - 1, i, i + 1);
    int count = n;
    while (count - 1 >= MR_MAXQPAR) {
        int sel = 0; double emin = ERR[0];
        for (int i = nxt[0]; i <= last; i = nxt[i]) if (ERR[i] < emin) { emin = ERR[i]; sel = i; }
        if (sel == 0 || sel == last) return 1;
        const int a = prv[sel], b = nxt[sel];
        if (a > 0) ERR[a] = thin_err(Q, T, prv[a], a, b);
        if (b < last) ERR[b] = thin_err(Q, T, a, b, nxt[b]);
        nxt[a] = (unsigned char)b; prv[b] = (unsigned char)a;
        --count;
    }
    int k = 0;
    for (int i = 0; i <= last; i = nxt[i]) { Q[k] = Q[i]; T[k] = T[i]; X[k] = X[i]; ++k; if (i == last) break; }
    n = k;
    return 0;
}

// kinwav_rch (kwt_route.f90:1130-1439).  Qj/Te/Tx point at element 1 of the reach arrays (the particles to
// route); on return elements 0..NQ2-1 hold flow, entry time and exit time, routed[] the FROUTE flags.
__device__ int kwt_kinwav(const DevNet &d, int p, double T_START, double T_END,
                          double *Qj, double *Te, double *Tx, unsigned &routed, int NQ1, int &NQ2) {
    signed char IX[NKIN], MF[NKIN];
    double T0[NKIN], T1[NKIN], Q0[NKIN], Q1[NKIN], Q2[NKIN], WC[NKIN];
    NQ2 = 0;
    if (NQ1 == 0) return 0;
    const double ALFA = 5.0 / 3.0;
    const double K = sqrt(d.rslope[p]) / d.rmann[p];
    const double XMX = d.rlength[p];
    const double p1 = 1.0 / ALFA, p2 = (ALFA - 1.0) / ALFA;
    int NN = NQ1;
    const int NI = NQ1;
    const double aK = ALFA * pow(K, p1);
    for (int i = 1; i <= NI; ++i) {
        MF[i] = (signed char)i; IX[i] = (signed char)i;
        Q0[i] = Q1[i] = Q2[i] = Qj[i - 1];
        T0[i] = T1[i] = Te[i - 1];
        WC[i] = aK * pow(Q1[i], p2);
    }
    if (NN > 1) {                                     // breaking waves, kwt_route.f90:1301-1349
        // The reference rescans every adjacent pair after each merge; a pair's crossing point changes only when
        // one of its two particles was merged, so the crossing points (and 1/WC) are cached and only the two
        // pairs around a merge are recomputed -- same operands, same divisions.  +inf marks "no crossing".
        double IWC[NKIN], XX[NKIN];
        const double NOX = __longlong_as_double(0x7ff0000000000000LL);
        for (int i = 1; i <= NN; ++i) IWC[i] = 1.0 / WC[i];
        auto cross = [&](int IW) -> double {
            const int JW = IW - 1;
            if (WC[IW] == 0.0 || WC[JW] == 0.0) return NOX;
            const double WDIFF = IWC[JW] - IWC[IW];
            if (WDIFF == 0.0) return NOX;
            if (WC[IW] == WC[JW]) return NOX;
            return (T1[IW] - T1[JW]) / WDIFF;
        };
        for (int IW = 2; IW <= NN; ++IW) XX[IW] = cross(IW);
        double X = 0.0;
        for (;;) {
            double XB = XMX; int IXB = 0;
            for (int IW = 2; IW <= NN; ++IW) {
                const double XXB = XX[IW];
                if (XXB < X || XXB > XB) continue;
                XB = XXB; IXB = IW;
            }
            if (XB == XMX) break;
            NN = NN - 1;
            const int JXB = IXB - 1;
            Q2[JXB] = fmax(Q2[JXB], Q2[IXB]);
            Q1[JXB] = fmin(Q1[JXB], Q1[IXB]);
            const double A2 = pow(Q2[JXB] / K, p1);
            const double A1 = pow(Q1[JXB] / K, p1);
            const double CM = (Q2[JXB] - Q1[JXB]) / (A2 - A1);
            T1[JXB] = T1[JXB] + XB / WC[JXB] - XB / CM;
            WC[JXB] = CM; IWC[JXB] = 1.0 / CM;
            for (int i = IX[IXB]; i <= NI; ++i) MF[i] = (signed char)(MF[i] - 1);
            for (int i = IXB; i <= NN; ++i) { IX[i] = IX[i + 1]; T1[i] = T1[i + 1]; WC[i] = WC[i + 1]; IWC[i] = IWC[i + 1]; Q1[i] = Q1[i + 1]; Q2[i] = Q2[i + 1]; XX[i] = XX[i + 1]; }
            if (JXB >= 2) XX[JXB] = cross(JXB);
            if (IXB <= NN) XX[IXB] = cross(IXB);
            X = XB;
        }
    }
    int ICOUNT = 0, bad = 0;
    auto rupdate = [&](double QNEW, double TOLD, double TNEW) {          // kwt_route.f90:1409-1437
        ICOUNT = ICOUNT + 1;
        if (ICOUNT > NQ1) { bad = 1; ICOUNT = NQ1; return; }
        Qj[ICOUNT - 1] = QNEW; Te[ICOUNT - 1] = TOLD; Tx[ICOUNT - 1] = TNEW;
        if (ICOUNT > 1) { if (Tx[ICOUNT - 1] <= Tx[ICOUNT - 2]) Tx[ICOUNT - 1] = Tx[ICOUNT - 2] + 1.0; }
        if (ICOUNT == 1 && Tx[0] <= T_START) Tx[0] = T_START + 1.0;
        if (Tx[ICOUNT - 1] < T_END) routed |= 1u << (ICOUNT - 1);
    };
    if (WC[1] < DBL_MIN) return 20;                                      // zero flow, kwt_route.f90:1365-1368
    double TEXIT = fmin(XMX / WC[1] + T1[1], DBL_MAX);
    for (int IR = 1; IR <= NN; ++IR) {
        double TNEXT = DBL_MAX;                                          // exit time of the next particle (computed once)
        if (IR < NN) {
            if (WC[IR + 1] < DBL_MIN) return 20;
            TNEXT = fmin(XMX / WC[IR + 1] + T1[IR + 1], DBL_MAX);
        }
        if (Q1[IR] != Q2[IR]) {
            if (TEXIT < T_END) {
                const double TEXIT2 = fmin(TEXIT + 1.0, TEXIT + 0.5 * (fmin(TNEXT, T_END) - TEXIT));
                if (TEXIT2 == TEXIT) return 30;
                rupdate(Q1[IR], T1[IR], TEXIT);
                rupdate(Q2[IR], T1[IR], TEXIT2);
            } else {
                for (int JR = 1; JR <= NI; ++JR) if (MF[JR] == IR) rupdate(Q0[JR], T0[JR], TEXIT);
            }
        } else {
            rupdate(Q1[IR], T1[IR], TEXIT);
        }
        TEXIT = TNEXT;
    }
    if (bad) return 60;
    NQ2 = ICOUNT;
    return 0;
}

// qexmul_rch (kwt_route.f90:619-993): merge the upstream basin series and routed-particle series into one
// particle stream QD/TD (written at Qo/To).  Upstream particle arrays are read in place from the buffer the
// upstream reaches wrote this step; nothing upstream is modified (the reference's strip, :840-844, is applied
// by the owner when it reads its own state back, see kwt_reach).
//
// The reference re-brackets every series at every emitted time (:930-957).  Times are emitted in ascending
// order and an un-exhausted series' current particle is never earlier than the emitted time, so the bracket
// of series s is always [itim-1, itim] and changes only when s itself advances: each upstream particle is
// loaded once and each segment slope computed once -- the same divisions on the same operands.
__device__ int kwt_merge_upstream(const DevNet &d, int p, int t, int b, double T0, double T1,
                                  double *Qo, double *To, int room, int &ND, int &nRead) {
    const int N = d.nRch;
    const int u0 = d.upPtr[p], NUPB = d.upPtr[p + 1] - u0;
    const double W = d.rwidth[p];
    const double *qr0 = d.qrSer + (size_t)t * N, *qr1 = d.qrSer + (size_t)(t + 1) * N;
    ND = 0;
    int NUPR = 0;
    for (int i = 0; i < NUPB; ++i) if (d.nGood[d.upIdx[u0 + i]] > 0) ++NUPR;
    const int NUPS = NUPB + NUPR;
    if (NUPS == 1) {                                   // single headwater upstream, kwt_route.f90:743-759
        Qo[0] = qr1[d.upIdx[u0]] / W; To[0] = T1; ND = 1;
        return 0;
    }
    if (NUPS > MAXSER) return -E_TOO_MANY_UPS;
    // per-series cursor: bracket [begin,end] = particles [itim-1, itim]
    double qb[MAXSER], tb[MAXSER], qe[MAXSER], te[MAXSER], slope[MAXSER], scfac[MAXSER];
    int upos[MAXSER]; short slen[MAXSER], nrt[MAXSER], itim[MAXSER];
    unsigned done = 0;
    const double *QF = d.kwQF[b], *TR = d.kwTR[b];
    int IMAX = NUPB, r = NUPB;
    for (int i = 0; i < NUPB; ++i) {
        const int U = d.upIdx[u0 + i];
        upos[i] = U; slen[i] = 2; nrt[i] = 2; itim[i] = 1;
        qb[i] = qr0[U]; tb[i] = T0; qe[i] = qr1[U]; te[i] = T1;
        slope[i] = (qe[i] - qb[i]) / (te[i] - tb[i]);
        scfac[i] = 1.0 / W;
        if (d.nGood[U] > 0) {
            const int NS = d.kwN[b][U], NR = d.kwNR[b][U];
            if (NS < 2 || NR < 1) return -E_NO_ROUTED_UP;
            upos[r] = U; slen[r] = (short)((NR + 1 < NS) ? NR + 1 : NS); nrt[r] = (short)NR; itim[r] = 1;
            qb[r] = QF[U]; tb[r] = TR[U]; qe[r] = QF[(size_t)N + U]; te[r] = TR[(size_t)N + U];
            slope[r] = (qe[r] - qb[r]) / (te[r] - tb[r]);
            scfac[r] = d.rwidth[U] / W;
            IMAX += NR - 1;
            nRead += slen[r];
            ++r;
        }
    }
    if (IMAX > room) return -E_SCRATCH;
    const unsigned all = (NUPS == 32) ? 0xffffffffu : ((1u << NUPS) - 1u);
    int IPRT = 0, jOld = -1, iOld = -1;
    double TIME_OLD = -DBL_MAX;
    for (;;) {
        int J = -1; double CT = DBL_MAX;               // MINLOC over CTIME: first minimum; exhausted series hold huge
        for (int s = 0; s < NUPS; ++s) {
            const double c = ((done >> s) & 1u) ? DBL_MAX : te[s];
            if (J < 0 || c < CT) { J = s; CT = c; }
        }
        if (J == jOld && itim[J] == iOld) return -E_STUCK;
        jOld = J; iOld = itim[J];
        if (!((done >> J) & 1u)) {
            if (!(itim[J] < nrt[J])) {                 // next particle not routed yet
                done |= 1u << J;
            } else {
                if (CT < TIME_OLD) return -E_TIME_ORDER;
                if (CT != TIME_OLD) {
                    double Q_AGG = 0.0;
                    for (int s = 0; s < NUPS; ++s) {
                        double SFLOW;
                        if (s == J) {
                            SFLOW = qe[s] * scfac[s];
                        } else {
                            if (te[s] < CT || tb[s] > CT) return -E_BRACKET;
                            const double PREDV = qb[s] + slope[s] * (CT - tb[s]);
                            SFLOW = PREDV * scfac[s];
                        }
                        Q_AGG = Q_AGG + SFLOW;
                    }
                    IPRT = IPRT + 1;
                    if (IPRT > IMAX) return -E_QD_BOUNDS;
                    Qo[IPRT - 1] = Q_AGG; To[IPRT - 1] = CT;
                    TIME_OLD = CT;
                }
                if (itim[J] == slen[J] - 1) {
                    done |= 1u << J;
                } else {                               // advance the cursor of series J (reach series only: basins have 2 points)
                    const int k = ++itim[J];
                    qb[J] = qe[J]; tb[J] = te[J];
                    qe[J] = QF[(size_t)k * N + upos[J]]; te[J] = TR[(size_t)k * N + upos[J]];
                    slope[J] = (qe[J] - qb[J]) / (te[J] - tb[J]);
                }
            }
        }
        if (done == all) break;
    }
    ND = IPRT;
    return 0;
}

// kwt_rch (kwt_route.f90:36-346) for interior reach p at batch step t (absolute step tau).
// State buffers: a reach writes its complete post-step particle array KWAVE(0:NQ2+1) and the number of routed
// entries NR into buffer tau&1.  What the reference removes afterwards -- the downstream reach strips
// KWAVE(0:NR-2) (:840-844), outlets and lake inlets strip themselves (:325-344) -- always leaves
// KWAVE(NR-1:), so the owner simply starts reading at NR-1 next step, and the consumer (exactly one wavefront
// behind, reading the same buffer) sees the unstripped array.
__device__ void kwt_reach(const DevNet &d, int p, int t, long long tau, double T0, double T1) {
    const int N = d.nRch;
    const int b = (int)(tau & 1), bp = b ^ 1;
    double *Qs = d.qSer[M_KWT] + (size_t)t * N;
    const double qr1 = d.qrSer[(size_t)(t + 1) * N + p];
    const int nGood = d.nGood[p];
    if (nGood == 0) {                                  // headwater, kwt_route.f90:181-205
        d.inflow[M_KWT][p] = 0.0;
        Qs[p] = qr1;
        d.kwN[b][p] = 1; d.kwNR[b][p] = 0;
        d.kwQF[b][p] = -9999.0; d.kwTI[b][p] = -9999.0; d.kwTR[b][p] = -9999.0;
        return;
    }
    double Q[WCAP], TE[WCAP], TX[WCAP];
    const int u0 = d.upPtr[p];
    const double W = d.rwidth[p];

    // getusq_rch, kwt_route.f90:461-613
    const int nPrev = d.kwN[bp][p], nrPrev = d.kwNR[bp][p];
    const int first = nrPrev > 0 ? nrPrev - 1 : 0;
    const int nOwn = nPrev > 0 ? nPrev - first : 1;
    for (int i = 0; i < nOwn && nPrev > 0; ++i) {
        Q[i] = d.kwQF[bp][(size_t)(first + i) * N + p];
        TE[i] = d.kwTI[bp][(size_t)(first + i) * N + p];
        TX[i] = d.kwTR[bp][(size_t)(first + i) * N + p];
    }
    int ND = 0, ND_read = 0;
    if (d.flags[p] & FLAG_LAKE_UP) {                   // lake outlet reach, kwt_route.f90:540-559
        if (d.upPtr[p + 1] - u0 > 1) { raise(d.err, 10, p, E_LAKE_UPS); return; }
        Q[nOwn] = Qs[d.upIdx[u0]] / W; TE[nOwn] = T1; ND = 1;
    } else {
        const int e = kwt_merge_upstream(d, p, t, b, T0, T1, Q + nOwn, TE + nOwn, WCAP - nOwn, ND, ND_read);
        if (e) { const int site = -e; raise(d.err, site == E_TIME_ORDER ? 30 : (site == E_BRACKET ? 40 : (site == E_STUCK ? 20 : 60)), p, site); return; }
    }
    if (nPrev == 0) {                                  // cold start, kwt_route.f90:587-596
        Q[0] = Q[nOwn]; TE[0] = T0 - (T1 - T0); TX[0] = T0;
    }
    for (int i = 0; i < ND; ++i) TX[nOwn + i] = -9999.0;
    int n = nOwn + ND;
    for (int i = 0; i < n; ++i) if (Q[i] < 0.0) { raise(d.err, 20, p, E_NEG_FLOW); return; }

    double qup = 0.0;                                  // kwt_route.f90:168-174
    for (int m = 0; m < nGood; ++m) qup = qup + Qs[d.upIdx[u0 + m]];
    d.inflow[M_KWT][p] = qup;

    if (n > MR_MAXQPAR) { if (kwt_thin(Q, TE, TX, n)) { raise(d.err, 60, p, E_THIN); return; } }

    const int NQ1 = n - 1;
    unsigned routed = 0;
    int NQ2;
    const int ek = kwt_kinwav(d, p, T0, T1, Q + 1, TE + 1, TX + 1, routed, NQ1, NQ2);
    if (ek) { raise(d.err, ek, p, ek == 20 ? E_ZERO_FLOW : (ek == 30 ? E_TEXIT2 : E_RUPDATE)); return; }
    const int NR = __popc(routed);                     // count(FROUTE)-1 (FROUTE(0) is always true)
    if (NR + 1 > NQ2) { raise(d.err, 21, p, E_NO_NONROUTED); return; }

    double QNEW;
    if (kwt_time_average(TX, Q, NR + 2, T0, T1, QNEW)) { raise(d.err, 40, p, E_INTERP); return; }
    Qs[p] = QNEW * W + qr1;                            // kwt_route.f90:273

    // end-of-step point, kwt_route.f90:288-292
    const double Q_END = Q[NR] + ((Q[NR + 1] - Q[NR]) / (TX[NR + 1] - TX[NR])) * (T1 - TX[NR]);
    const double TIMEI = TE[NR] + ((TE[NR + 1] - TE[NR]) / (TX[NR + 1] - TX[NR])) * (T1 - TX[NR]);

    // KWAVE(0:NQ2+1) = routed(0:NR) | end-of-step point | non-routed(NR+1:NQ2), kwt_route.f90:299-311
    double *oQ = d.kwQF[b] + p, *oI = d.kwTI[b] + p, *oR = d.kwTR[b] + p;
    for (int i = 0; i <= NR; ++i) { oQ[(size_t)i * N] = Q[i]; oI[(size_t)i * N] = TE[i]; oR[(size_t)i * N] = TX[i]; }
    oQ[(size_t)(NR + 1) * N] = Q_END; oI[(size_t)(NR + 1) * N] = TIMEI; oR[(size_t)(NR + 1) * N] = T1;
    for (int i = NR + 1; i <= NQ2; ++i) { oQ[(size_t)(i + 1) * N] = Q[i]; oI[(size_t)(i + 1) * N] = TE[i]; oR[(size_t)(i + 1) * N] = TX[i]; }
    d.kwN[b][p] = NQ2 + 2;
    d.kwNR[b][p] = NR + 2;
    if (d.kwCount) d.kwCount[p] += (unsigned)(nOwn + ND_read + NQ2 + 2);
}

// ------------------------------------------------------------------------------------------------
// per-reach bodies of the three methods (route_network loop body, main_route.f90:372-390)
// ------------------------------------------------------------------------------------------------
template <int M, bool HEAD>
__device__ __forceinline__ void route_reach(const DevNet &d, int p, int t, long long tau) {
    const int N = d.nRch;
    const int flags = d.flags[p];
    if (flags & FLAG_GHOST) return;
    if (M != M_SUM && (flags & FLAG_LAKE)) { lake_reach<M>(d, p, t, tau); return; }
    if (M == M_KWT && HEAD) {                          // no upstream reach => count(goodBas)=0, kwt_route.f90:181-205
        const int b = (int)(tau & 1);
        d.inflow[M_KWT][p] = 0.0;
        d.qSer[M_KWT][(size_t)t * N + p] = d.qrSer[(size_t)(t + 1) * N + p];
        d.kwN[b][p] = 1; d.kwNR[b][p] = 0;
        d.kwQF[b][p] = -9999.0; d.kwTI[b][p] = -9999.0; d.kwTR[b][p] = -9999.0;
        return;
    }
    if (M == M_SUM) {                                  // accum_runoff.f90:60-75
        double *Qs = d.qSer[M_SUM] + (size_t)t * N;
        const int u0 = d.upPtr[p], u1 = d.upPtr[p + 1];
        double q = d.qrSer[(size_t)(t + 1) * N + p];
        if (u1 > u0) {
            double qup = 0.0;
            for (int m = u0; m < u1; ++m) qup = qup + Qs[d.upIdx[m]];
            q = q + qup;
        }
        Qs[p] = q;
    } else if (M == M_IRF) {                           // irf_route.f90:82-150,235-262
        double *Qs = d.qSer[M_IRF] + (size_t)t * N;
        const int nUps = d.nGood[p], u0 = d.upPtr[p];
        const double qr1 = d.qrSer[(size_t)(t + 1) * N + p], dt = d.dt;
        double v1 = d.vol1[M_IRF][p], v0 = v1;
        double qup = 0.0, qlat = 0.0;
        if (nUps > 0) {
            for (int m = 0; m < nUps; ++m) qup = qup + Qs[d.upIdx[u0 + m]];
            qlat = qr1;
        } else if (d.hwDrain == 1) { qup = qup + qr1; qlat = 0.0; }
        else if (d.hwDrain == 2) { qlat = qr1; }
        d.inflow[M_IRF][p] = qup;
        const int nt = d.ntdh[p];
        double *qf = d.qfutIrf + p;
        const double *uh = d.uh + p;
        double q;
        if (d.rlength[p] > d.minLengthRoute) {
            const int head = (int)(tau % nt);
            int s = head;
            for (int k = 0; k < nt; ++k) {
                qf[(size_t)s * N] = qf[(size_t)s * N] + uh[(size_t)k * N] * qup;
                if (++s == nt) s = 0;
            }
            double q1 = qf[(size_t)head * N];
            q1 = fmin((fmax(0.0, v1) / dt + qup) * 0.999, q1);
            v1 = v1 - (q1 - qup) * dt;
            q = q1 + qlat;
            qf[(size_t)head * N] = 0.0;
        } else {                                       // pass-through, irf_route.f90:255-262
            const int nxt = (int)((tau + 1) % nt);
            for (int k = 0; k < nt; ++k) qf[(size_t)k * N] = 0.0;
            qf[(size_t)nxt * N] = qup;                 // logical slot 0 as seen by the next step / by mr_get_state
            q = qup + qlat;
            v0 = 0.0; v1 = 0.0;
        }
        Qs[p] = q;
        d.vol0[M_IRF][p] = v0; d.vol1[M_IRF][p] = v1;
        d.wb[M_IRF][p] = reach_wb(v1, v0, qup, qlat, q, dt);
    } else {
        kwt_reach(d, p, t, tau, d.T0s[t], d.T1s[t]);
    }
}

// headwater reaches (positions [0, nHead)): no upstream dependency, so one thread routes all K steps
template <int M>
__global__ void __launch_bounds__(256) k_headwater(DevNet d, int K, long long tau0) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= d.nHead) return;
    for (int t = 0; t < K; ++t) route_reach<M, true>(d, p, t, tau0 + t);
}

// one wavefront of interior reaches: positions [lo,hi) hold stages w-K+1..w; the reach at stage s does step t = w - s
template <int M>
__global__ void __launch_bounds__(M == M_KWT ? 128 : 256) k_route(DevNet d, int lo, int hi, int w, long long tau0) {
    const int p = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= hi) return;
    const int t = w - d.stageOf[p];
    route_reach<M, false>(d, p, t, tau0 + t);
}

// ------------------------------------------------------------------------------------------------
// order conversion: stage order <-> caller's reach order
// ------------------------------------------------------------------------------------------------
// out[row][rch] = in[row][pos(rch)]; rows = methods x steps
__global__ void k_unpermute_rows(const double *in, double *out, const int *rch2pos, int N, int rows) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= N) return;
    const int p = rch2pos[r];
    for (int k = blockIdx.y; k < rows; k += gridDim.y) out[(size_t)k * N + r] = in[(size_t)k * N + p];
}

}  // namespace mr
