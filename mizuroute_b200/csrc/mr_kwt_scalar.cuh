// Thread-per-task kinematic-wave-tracking reach step: kwt_rch and callees (kwt_route.f90:36-1622) for the common case.
//
// Four out of five (reach, step) tasks of a network hold at most MR_MAXQPAR particles before routing (no thinning,
// remove_rch), are fed by at most KWS_BMAX upstream reaches, and neither break a wave (kinwav_rch's shock merge) nor need
// rUpdate's exit-time fix-ups.  For those the whole reach step is a short sequential program -- the reference's own loop
// structure -- and ONE LANE runs it: the k-way time merge of qexmul_rch walks the upstream wave rows with a cursor per
// series, the merged particles go to a per-thread column of shared memory (element i of thread x at [i * CS + x]: no bank
// conflict whatever the index), and kinwav_rch / interp_rch / the new wave are evaluated in one streaming pass that keeps
// only the previous particle in registers.  A warp thus carries 32 tasks instead of two.
//
// The 32 lanes of a warp must stay converged to be worth anything: every loop whose trip count depends on the task runs
// for the warp's maximum (MR_WARP_MAX / MR_WARP_ANY, mr_lanes.h) behind a warp-wide sync, a lane that is done or has given
// up idles through the remaining trips, and there is a single exit.  All 32 lanes call the function (active = false: no task).
//
// Everything else -- thinning, shocks, lakes, ghosts, water management, exported outlets, wide confluences, any anomaly the
// reference reports as an error -- ends in KWS_DEFER before a result the consumers read has been written, and the task is
// then routed by the team code (kwt_reach_team, mr_kwt.cuh), which also owns every error message.  What the scalar path has
// written by then (part of its own row in the step's particle buffer, REACH_INFLOW) is rewritten by the team code.
// Both paths evaluate the same floating-point operations on the same operands in the same order, so which path routes a
// task does not change a bit of the result (tests/test_kwt_emul.py runs both on the host against the oracle).
#pragma once
#include "mr_kwt.cuh"

namespace mr {

constexpr int KWS_NL = MR_MAXQPAR;   // particles a task may hold (own + merged); more = thinning = team code
constexpr int KWS_BMAX = 3;          // upstream reaches (basin series; at most as many wave series)
enum { KWS_DONE = 0, KWS_DEFER = 1 };

// cQ / cT: this thread's columns of KWS_NL doubles each, stride CS
template <bool EXT, int CS>
MR_DEV int kwt_reach_scalar(const DevNet &d, double *cQ, double *cT, int p, int t, long long tau, double T0, double T1, bool active = true) {
    const int N = d.nRch;
    const int b = (int)(tau & 1), bp = b ^ 1;
    bool live = active;                                // still routing its task on this path
    bool defer = false;
    if (live) { if (d.flags[p] & (FLAG_GHOST | FLAG_LAKE | FLAG_LAKE_UP)) { defer = true; live = false; } }
    if (live && d.expSlot) { if (d.expSlot[p] >= 0) { defer = true; live = false; } }
    if (EXT) { if (live && d.wmFlux) { if (d.wmFlux[(size_t)t * N + p] != -9999.0) { defer = true; live = false; } } }
    double *Qs = d.qSer[M_KWT] + (size_t)t * N;
    const double *qr0row = d.qrSer + (size_t)t * N, *qr1row = d.qrSer + (size_t)(t + 1) * N;
    const size_t row = (size_t)(live ? p : 0) * KWP;
    double qr1 = 0.0, W = 1.0;
    int nGood = 0, u0 = 0, NUPB = 0, nPrev = 0, nrPrev = 0;
    if (live) {
        qr1 = qr1row[p];
        nGood = d.nGood[p];
        if (nGood == 0) {                              // no contributing area upstream, kwt_route.f90:181-205
            d.inflow[M_KWT][p] = 0.0;
            Qs[p] = qr1;
            d.kwN[b][p] = 1; d.kwNR[b][p] = 0;
            d.kwQF[b][row] = -9999.0; d.kwTI[b][row] = -9999.0; d.kwTR[b][row] = -9999.0;
            live = false;
        }
    }
    if (live) {
        u0 = d.upPtr[p]; NUPB = d.upPtr[p + 1] - u0;
        if (NUPB > KWS_BMAX) { defer = true; live = false; NUPB = 0; }
    }
    if (live) { W = d.rwidth[p]; nPrev = d.kwN[bp][p]; nrPrev = d.kwNR[bp][p]; }
    const int first = nrPrev > 0 ? nrPrev - 1 : 0;
    int nOwn = nPrev > 0 ? nPrev - first : 1;

    // ---- the upstream reaches: basin series (T0, QR0), (T1, QR1) of each, wave series of those with contributing area
    int U[KWS_BMAX], cur[KWS_BMAX], nc[KWS_BMAX], cmax[KWS_BMAX];
    double bq0[KWS_BMAX], bq1[KWS_BMAX], scf[KWS_BMAX], tcur[KWS_BMAX];
    bool isr[KWS_BMAX];
    int sumNc = 0, nRead = 0, NUPR = 0;
    double qup = 0.0;                                  // REACH_INFLOW, kwt_route.f90:168-174
#pragma unroll
    for (int i = 0; i < KWS_BMAX; ++i) {
        U[i] = 0; cur[i] = 1; nc[i] = 0; cmax[i] = 0; bq0[i] = 0.0; bq1[i] = 0.0; scf[i] = 0.0; tcur[i] = 0.0; isr[i] = false;
        if (i < NUPB) {
            const int u = d.upIdx[u0 + i];
            U[i] = u;
            bq0[i] = qr0row[u]; bq1[i] = qr1row[u];
            if (i < nGood) qup = qup + Qs[u];
            if (d.nGood[u] > 0) {
                const int NS = d.kwN[b][u], NR = d.kwNR[b][u];
                if (NS < 2 || NR < 1) defer = true;    // "upstream wave has no routed element": the team code reports it
                else {
                    const int sl = NR + 1 < NS ? NR + 1 : NS;
                    int c = sl - 1 < NR - 1 ? sl - 1 : NR - 1;
                    if (c < 0) c = 0;
                    isr[i] = true; nc[i] = c; cmax[i] = NR < sl - 1 ? NR : sl - 1;
                    scf[i] = d.rwidth[u] / W;
                    sumNc += c; nRead += sl; ++NUPR;
                }
            }
        }
    }
    // at most one particle per wave candidate before T1 plus the one all series share at T1
    if (live && (defer || nOwn + sumNc + 1 > KWS_NL)) { defer = true; live = false; }
    if (!live) { NUPB = 0; nOwn = 0; }
#pragma unroll
    for (int i = 0; i < KWS_BMAX; ++i) if (!live) isr[i] = false;

    // ---- own wave KWAVE(NR-1:) of the previous step, kwt_route.f90:461-613
    double TX0 = 0.0;
    {
        const int nCopy = (live && nPrev > 0) ? nOwn : 0;
        const double *oq = d.kwQF[bp] + row + first, *oi = d.kwTI[bp] + row + first;
        const int m = MR_WARP_MAX(nCopy);
        MR_NOUNROLL
        for (int i = 0; i < m; ++i) if (i < nCopy) { cQ[i * CS] = oq[i]; cT[i * CS] = oi[i]; }
        if (nCopy) TX0 = d.kwTR[bp][row + first];
    }
    if (live) d.inflow[M_KWT][p] = qup;

    // ---- qexmul_rch (kwt_route.f90:619-993): candidates in (time, series) order, one particle per distinct time
    int ND = 0;
    const double scfB = 1.0 / W;
    bool merging = false;
    if (live) {
        if (NUPB == 1 && NUPR == 0) {                  // single headwater upstream, kwt_route.f90:743-759
            cQ[nOwn * CS] = bq1[0] / W; cT[nOwn * CS] = T1;
            ND = 1;
        } else merging = true;
    }
    {
        bool odd = false;                              // an ordering the reference treats as an error
        double bsl[KWS_BMAX];                          // SLOPE of the basin series: the same operands at every emission
#pragma unroll
        for (int s = 0; s < KWS_BMAX; ++s) bsl[s] = (merging && s < NUPB) ? (bq1[s] - bq0[s]) / (T1 - T0) : 0.0;
        // the flow of every series at time CT; the series that supplies CT (the first basin at T1: Ji < 0; the wave of
        // upstream Ji at its point k otherwise) contributes its own point, the others are interpolated in their bracket
        auto emit = [&](double CT, int Ji, int k) {
            double Q_AGG = 0.0;
#pragma unroll
            for (int s = 0; s < KWS_BMAX; ++s) {
                if (s < NUPB) {
                    double SFLOW;
                    if (Ji < 0 && s == 0) SFLOW = bq1[0] * scfB;
                    else { const double PREDV = bq0[s] + bsl[s] * (CT - T0); SFLOW = PREDV * scfB; }
                    Q_AGG = Q_AGG + SFLOW;
                }
            }
#pragma unroll
            for (int s = 0; s < KWS_BMAX; ++s) {
                if (isr[s]) {
                    const double *QF = d.kwQF[b] + (size_t)U[s] * KWP, *TR = d.kwTR[b] + (size_t)U[s] * KWP;
                    double SFLOW;
                    if (s == Ji) SFLOW = QF[k] * scf[s];
                    else {
                        const int cu = cur[s] > cmax[s] ? cmax[s] : cur[s];
                        const double tb = TR[cu - 1], te = TR[cu], qb = QF[cu - 1], qe = QF[cu];
                        if (te < CT || tb > CT) odd = true;
                        const double SLOPE = (qe - qb) / (te - tb);
                        const double PREDV = qb + SLOPE * (CT - tb);
                        SFLOW = PREDV * scf[s];
                    }
                    Q_AGG = Q_AGG + SFLOW;
                }
            }
            cQ[(nOwn + ND) * CS] = Q_AGG; cT[(nOwn + ND) * CS] = CT;
            ++ND;
        };
#pragma unroll
        for (int s = 0; s < KWS_BMAX; ++s) if (merging && isr[s] && nc[s] >= 1) tcur[s] = d.kwTR[b][(size_t)U[s] * KWP + 1];
        bool basinsDone = false, any = false;
        double lastT = 0.0;
        MR_NOUNROLL
        while (MR_WARP_ANY(merging)) {
            MR_WARP_SYNC();
            if (merging) {
                int best = -1, bk = 1; double bt = 0.0;
#pragma unroll
                for (int s = 0; s < KWS_BMAX; ++s) if (isr[s] && cur[s] <= nc[s]) { if (best < 0 || tcur[s] < bt) { best = s; bt = tcur[s]; bk = cur[s]; } }
                if (!basinsDone && (best < 0 || T1 <= bt)) {   // the basin series (lower series numbers) come first at T1
                    if (any && T1 < lastT) odd = true;
                    else if (!any || T1 != lastT) { emit(T1, -1, 1); lastT = T1; any = true; }
                    basinsDone = true;
                } else if (best < 0) {
                    merging = false;
                } else {
                    if (any && bt < lastT) odd = true;
                    else if (!any || bt != lastT) { emit(bt, best, bk); lastT = bt; any = true; }
#pragma unroll
                    for (int s = 0; s < KWS_BMAX; ++s) {
                        if (s == best) {
                            cur[s] = cur[s] + 1;
                            if (cur[s] <= nc[s]) { const double tn = d.kwTR[b][(size_t)U[s] * KWP + cur[s]]; if (tn < tcur[s]) odd = true; tcur[s] = tn; }
                        }
                    }
                }
                if (odd) { merging = false; defer = true; live = false; }
            }
        }
    }
    if (live && nPrev == 0) {                          // cold start, kwt_route.f90:587-596
        cQ[0] = cQ[nOwn * CS]; cT[0] = T0 - (T1 - T0); TX0 = T0;
    }
    const int n = live ? nOwn + ND : 0;
    if (live && TX0 > T0) { defer = true; live = false; }                          // interp_rch "bad bounds"

    // ---- kinwav_rch (kwt_route.f90:1130-1439) without wave breaking, interp_rch (:1444-1622) and the new wave
    // KWAVE(0:NQ2+1) = routed(0:NR) | end-of-step point | non-routed (:299-311), in one pass over the particles
    const double aK = live ? d.kwAK[p] : 1.0, XMX = live ? d.rlength[p] : 1.0;
    double *oQ = d.kwQF[b] + row, *oI = d.kwTI[b] + row, *oR = d.kwTR[b] + row;
    double qp = 0.0, ep = 0.0, xp = TX0;               // previous point: flow, entry time, exit time
    double wcp = 0.0, iwcp = 0.0;
    if (live) {
        qp = cQ[0]; ep = cT[0];
        if (qp < 0.0) { defer = true; live = false; }  // "negative flow extracted from upstream reach"
        else { oQ[0] = qp; oI[0] = ep; oR[0] = xp; }
    }
    int NR = -1;                                       // number of routed particles once the first non-routed one is met
    bool begFound = false; int IBEG = 0;
    double AREAB = 0.0, AREAM = 0.0, QNEW = 0.0;
    const int nMax = MR_WARP_MAX(live ? n : 0);
    MR_NOUNROLL
    for (int i = 1; i < nMax; ++i) {
        MR_WARP_SYNC();
        if (live && i < n) {
            const double q = cQ[i * CS], e = cT[i * CS];
            const double wc = aK * mr_pow04(q);
            const double iwc = 1.0 / wc;
            bool giveUp = q < 0.0 || wc < DBL_MIN;     // "negative flow", "zero flow"
            if (i >= 2 && wc != 0.0 && wcp != 0.0) {   // would particles i-1 and i cross inside the reach? (:1308-1319)
                const double WDIFF = iwcp - iwc;
                if (WDIFF != 0.0 && wc != wcp) {
                    const double XXB = (e - ep) / WDIFF;
                    if (!(XXB < 0.0 || XXB > XMX) && XXB != XMX) giveUp = true;     // a shock: merged by the team code
                }
            }
            const double x = fmin(XMX / wc + e, DBL_MAX);
            if (i == 1 ? x <= T0 : x <= xp) giveUp = true;                          // rUpdate would move the exit time
            const bool routed = x < T1;
            if (!routed && !(x >= T1)) giveUp = true;
            if (NR < 0 && !routed && !begFound && !(T0 <= x)) giveUp = true;
            if (giveUp) { defer = true; live = false; }
            else {
                if (NR < 0) {
                    // interp_rch over the points (T_EXIT, Q)(0 : NR+1), this one included
                    if (!begFound) {
                        if (T0 <= x) {
                            begFound = true; IBEG = i;
                            if (T1 < x) {
                                const double SLOPE = (q - qp) / (x - xp);
                                const double QEST0 = SLOPE * (T0 - xp) + qp;
                                const double QEST1 = SLOPE * (T1 - xp) + qp;
                                QNEW = 0.5 * (QEST0 + QEST1);
                            } else if (T0 < x) {
                                const double SLOPE = (q - qp) / (x - xp);
                                const double QEST0 = SLOPE * (T0 - xp) + qp;
                                AREAB = (x - T0) * 0.5 * (QEST0 + q);
                            }
                        }
                    } else if (routed) {
                        AREAM = AREAM + (x - xp) * 0.5 * (qp + q);
                    }
                    if (!routed) {
                        NR = i - 1;
                        if (!(IBEG == i && T1 < x)) {
                            double AREAE = 0.0;
                            if (T1 < x) {
                                const double SLOPE = (q - qp) / (x - xp);
                                const double QEST1 = SLOPE * (T1 - xp) + qp;
                                AREAE = (T1 - xp) * 0.5 * (qp + QEST1);
                            }
                            if (IBEG < i) { if (T1 == x && T0 < xp) AREAM = AREAM + (x - xp) * 0.5 * (qp + q); }
                            QNEW = (AREAB + AREAE + AREAM) / (T1 - T0);
                        }
                        // end-of-step point, kwt_route.f90:288-292
                        const double Q_END = qp + ((q - qp) / (x - xp)) * (T1 - xp);
                        const double TIMEI = ep + ((e - ep) / (x - xp)) * (T1 - xp);
                        oQ[i] = Q_END; oI[i] = TIMEI; oR[i] = T1;
                    }
                }
                const int j = NR < 0 ? i : i + 1;
                oQ[j] = q; oI[j] = e; oR[j] = x;
                qp = q; ep = e; xp = x; wcp = wc; iwcp = iwc;
            }
        }
    }
    if (live && NR < 0) { defer = true; live = false; }                            // "no non-routed particle left"
    if (live) {
        Qs[p] = QNEW * W + qr1;                        // kwt_route.f90:273
        d.kwN[b][p] = n + 1;                           // NQ2 + 2
        d.kwNR[b][p] = NR + 2;
        if (d.kwCount) d.kwCount[p] += (unsigned)(nOwn + nRead + n + 1);
    }
    return defer ? KWS_DEFER : KWS_DONE;
}

}  // namespace mr
