// Team-cooperative kinematic-wave-tracking reach step: kwt_rch and callees, kwt_route.f90:36-1622.
//
// One team (MR_TEAM = 16 lanes, half a warp, on the device; see mr_lanes.h) routes one (reach, step); the two teams of
// a warp re-converge at the phase boundaries of their tasks (kwt_reach_team).  The wave particles of the reach live
// in the team's shared-memory scratch, one particle per lane:
//   getusq_rch/qexmul_rch  every candidate exit time of the upstream series is ranked and evaluated by its own
//                          lane (the reference's sequential k-way MINLOC merge, :895-976, visits candidates in
//                          (time, series) order; rank, duplicate flag and interpolation brackets of a candidate
//                          are functions of that order alone, so they are computed independently);
//   remove_rch             lane-parallel error evaluation, team argmin ("first minimum", :1081) per removal;
//   kinwav_rch             lane-parallel celerity (pow), crossing points and exit times; team argmin per shock
//                          merge; exit times are final in parallel unless rUpdate's +1 s fix-up applies, in which case
//                          (and for merged particles) the order-dependent emission runs on lane 0 as the reference's
//                          serial loop; interp_rch's trapezoids are evaluated by the lanes and added in order by lane 0.
// Every floating-point value is produced by the same operations on the same operands as the serial
// restatement, so results do not depend on the lane count.
//
// HBM layout of the wave state (per buffer b, see kwt_reach_team): kwQF/kwTI/kwTR[b][p*KWP + k], kwN/kwNR[b][p].
#pragma once
#include <cstring>
#include "mr_dev.h"

namespace mr {

// Per-team scratch.  Two sizes: the small one lives in shared memory and covers all but the widest confluences;
// a task that does not fit (kwt_reach_team returns KWT_RETRY before it has changed any state) is re-run with the
// full-capacity scratch, which lives in a global-memory arena (kwt_task, mr_kernels.cuh).
template <int WCAP_, int POOL_>
struct KwtScratchT {
    static constexpr int CAP = WCAP_, PCAP = POOL_;
    double Q[WCAP_], TE[WCAP_];        // Q_JRCH, TENTRY (0:n-1)
    double TX[KWP];                    // T_EXIT: only element 0 is an input; 1..NQ2 are written by kinwav
    union {
        struct {                       // qexmul_rch: staged upstream series
            double sq[POOL_], st[POOL_]; // flow and time of every staged point
            double scf[MAXSER];        // UWIDTH / R_WIDTH(JRCH)
            int upos[MAXSER];
            short soff[MAXSER], slen[MAXSER], ncand[MAXSER], cmax[MAXSER], cbase[MAXSER];
            unsigned char flag[WCAP_];
        } m;
        struct {                       // remove_rch
            double ERR[WCAP_];
            unsigned char prv[WCAP_], nxt[WCAP_];
        } th;
        struct {                       // kinwav_rch
            double T0[NKIN], T1[NKIN], Q0[NKIN], Q1[NKIN], Q2[NKIN], WC[NKIN], IWC[NKIN], XX[NKIN], TEX[NKIN];
            signed char IX[NKIN], MF[NKIN];
        } k;
    } u;
};
using KwtScratch = KwtScratchT<WCAP, POOL>;             // full capacity
using KwtScratchSmall = KwtScratchT<WCAP_S, POOL_S>;    // shared-memory fast path
constexpr int KWT_RETRY = 1;

// interp_rch (kwt_route.f90:1444-1622) for a single output interval (TOLD/QOLD 0-based), as a team: the trapezoids
// of the middle part are evaluated by the lanes and added by lane 0 in the
// reference's order.  W is scratch of at least NOLD doubles.  QNEW is valid on lane 0.
MR_DEV int kwt_time_average_team(const double *TOLD, const double *QOLD, int NOLD, double T0, double T1, double *W, double &QNEW) {
    const int lane = MR_LANE;
    if (TOLD[0] > T0 || TOLD[NOLD - 1] < T1) return 1;
    int ib = 0x7fffffff, ie = 0x7fffffff;
    MR_NOUNROLL
    for (int i = lane; i < NOLD; i += MR_NL) {
        const double tt = TOLD[i];
        if (i >= 1 && T0 <= tt && i < ib) ib = i;
        if (T1 <= tt && i < ie) ie = i;
    }
    int IBEG = team_min(ib), IEND = team_min(ie);
    if (IBEG == 0x7fffffff) IBEG = 0;
    if (IEND == 0x7fffffff) IEND = 0;
    MR_NOUNROLL
    for (int IMID = IBEG + 1 + lane; IMID <= IEND; IMID += MR_NL)
        W[IMID] = (TOLD[IMID] - TOLD[IMID - 1]) * 0.5 * (QOLD[IMID - 1] + QOLD[IMID]);
    // the slopes of the first and of the last segment (one division each) on two lanes
    const int L1 = MR_NL > 1 ? 1 : 0;
    double slB, slE;
    {
        const int I = (lane == L1 && MR_NL > 1) ? IEND : IBEG;
        double sl = 0.0;
        if (I >= 1) sl = (QOLD[I] - QOLD[I - 1]) / (TOLD[I] - TOLD[I - 1]);
        slB = team_bcast(sl, 0);
        if (MR_NL > 1) slE = team_bcast(sl, L1);
        else slE = IEND >= 1 ? (QOLD[IEND] - QOLD[IEND - 1]) / (TOLD[IEND] - TOLD[IEND - 1]) : 0.0;
    }
    MR_SYNC();
    if (lane == 0) {
        if (T1 < TOLD[IBEG]) {
            const double SLOPE = slB;
            const double QEST0 = SLOPE * (T0 - TOLD[IBEG - 1]) + QOLD[IBEG - 1];
            const double QEST1 = SLOPE * (T1 - TOLD[IBEG - 1]) + QOLD[IBEG - 1];
            QNEW = 0.5 * (QEST0 + QEST1);
        } else {
            double AREAB = 0.0, AREAE = 0.0, AREAM = 0.0;
            if (T0 < TOLD[IBEG]) {
                const double SLOPE = slB;
                const double QEST0 = SLOPE * (T0 - TOLD[IBEG - 1]) + QOLD[IBEG - 1];
                AREAB = (TOLD[IBEG] - T0) * 0.5 * (QEST0 + QOLD[IBEG]);
            }
            if (T1 < TOLD[IEND]) {
                const double SLOPE = slE;
                const double QEST1 = SLOPE * (T1 - TOLD[IEND - 1]) + QOLD[IEND - 1];
                AREAE = (T1 - TOLD[IEND - 1]) * 0.5 * (QOLD[IEND - 1] + QEST1);
            }
            if (IBEG < IEND) {
                MR_NOUNROLL
                for (int IMID = IBEG + 1; IMID < IEND; ++IMID) AREAM = AREAM + W[IMID];
                if (T1 == TOLD[IEND] && T0 < TOLD[IEND - 1]) AREAM = AREAM + W[IEND];
            }
            QNEW = (AREAB + AREAE + AREAM) / (T1 - T0);
        }
    }
    return 0;
}


MR_DEV_NOINLINE double thin_err(const double *Q, const double *T, int a, int m, int b) {
    // |INTERP(T(m), Q(a), Q(b), T(a), T(b)) - Q(m)|, kwt_route.f90:1054,1062,1114-1121
    return fabs((Q[a] + ((Q[b] - Q[a]) / (T[b] - T[a])) * (T[m] - T[a])) - Q[m]);
}

// ------------------------------------------------------------------------------------------------
// remove_rch (kwt_route.f90:999-1123): greedy removal of the particle whose linear-interpolation error is
// smallest until MAXQPAR remain.  The reference re-packs index arrays each pass; a doubly linked list of
// survivors visits them in the same order, so "first minimum" picks the same particle.  T_EXIT needs no
// compaction: only element 0 (never removed) is read afterwards.
// ------------------------------------------------------------------------------------------------
template <class SC>
MR_DEV int kwt_thin_team(SC &S, int &n) {
    const int lane = MR_LANE;
    double *Q = S.Q, *T = S.TE, *ERR = S.u.th.ERR;
    unsigned char *prv = S.u.th.prv, *nxt = S.u.th.nxt;
    const int last = n - 1;
    MR_NOUNROLL
    for (int i = lane; i < n; i += MR_NL) {
        prv[i] = (unsigned char)(i - 1); nxt[i] = (unsigned char)(i + 1);
        ERR[i] = (i > 0 && i < last) ? thin_err(Q, T, i - 1, i, i + 1) : DBL_MAX;
    }
    MR_SYNC();
    int count = n;
    MR_NOUNROLL
    while (count - 1 >= MR_MAXQPAR) {
        // removed particles hold ERR = DBL_MAX like the two ends, so a strict "<" never selects them
        double emin = DBL_MAX; int sel = 0x7fffffff;
        MR_NOUNROLL
        for (int i = lane; i < n; i += MR_NL) if (ERR[i] < emin) { emin = ERR[i]; sel = i; }
        team_argmin_first(emin, sel);
        if (sel <= 0 || sel >= last) return 1;
        const int a = prv[sel], b = nxt[sel];
        MR_SYNC();
        // the errors of the two neighbours change: lane 0 re-evaluates the left one, lane 1 the right one, in one pass
        if (MR_NL > 1) {
            if (lane < 2) {
                const bool left = lane == 0;
                const int m = left ? a : b;
                const bool need = left ? a > 0 : b < last;
                if (need) ERR[m] = thin_err(Q, T, left ? prv[a] : a, m, left ? b : nxt[b]);
                if (left) ERR[sel] = DBL_MAX;
            }
        } else {
            if (a > 0) ERR[a] = thin_err(Q, T, prv[a], a, b);
            ERR[sel] = DBL_MAX;
            if (b < last) ERR[b] = thin_err(Q, T, a, b, nxt[b]);
        }
        MR_SYNC();
        if (lane == 0) { nxt[a] = (unsigned char)b; prv[b] = (unsigned char)a; nxt[sel] = 255; }
        MR_SYNC();
        --count;
    }
    // compact the survivors (positions only move left)
    int pos = 0;
    MR_NOUNROLL
    for (int base = 0; base < n; base += MR_NL) {
        const int i = base + lane;
        const bool keep = i < n && nxt[i] != 255;
        double qv = 0.0, tv = 0.0;
        if (keep) { qv = Q[i]; tv = T[i]; }
        MR_SYNC();
        int tot;
        const int r = team_rank(keep, tot);
        if (keep) { Q[pos + r] = qv; T[pos + r] = tv; }
        pos += tot;
        MR_SYNC();
    }
    n = pos;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// kinwav_rch (kwt_route.f90:1130-1439) on particles S.Q/S.TE[1..NQ1]; on return elements 1..NQ2 of
// S.Q/S.TE/S.TX hold flow, entry time and exit time, `routed` the FROUTE flags (bit i-1 = particle i).
// Returns the reference's ierr (20 zero flow, 30 TEXIT==TEXIT2, 60 rUpdate bounds), identical on all lanes.
// ------------------------------------------------------------------------------------------------
template <class SC>
MR_DEV int kwt_kinwav_team(const DevNet &d, SC &S, int p, double T_START, double T_END, int NQ1, int &NQ2, unsigned &routed) {
    const int lane = MR_LANE;
    double *T0 = S.u.k.T0, *T1 = S.u.k.T1, *Q0 = S.u.k.Q0, *Q1 = S.u.k.Q1, *Q2 = S.u.k.Q2, *WC = S.u.k.WC, *IWC = S.u.k.IWC,
           *XX = S.u.k.XX, *TEX = S.u.k.TEX;
    signed char *IX = S.u.k.IX, *MF = S.u.k.MF;
    NQ2 = 0; routed = 0;
    if (NQ1 == 0) return 0;
    const double ALFA = 5.0 / 3.0;
    const double K = d.kwK[p];                         // sqrt(R_SLOPE)/R_MAN_N          (k_kwt_params, once per network)
    const double aK = d.kwAK[p];                       // ALFA*K**(1/ALFA)
    const double XMX = d.rlength[p];
    const double p1 = 1.0 / ALFA;                      // (the celerity exponent (ALFA-1)/ALFA is inside mr_pow04)
    int NN = NQ1;
    const int NI = NQ1;
    MR_NOUNROLL
    for (int i = 1 + lane; i <= NI; i += MR_NL) {
        MF[i] = (signed char)i; IX[i] = (signed char)i;
        const double q = S.Q[i], te = S.TE[i];
        Q0[i] = q; Q1[i] = q; Q2[i] = q;
        T0[i] = te; T1[i] = te;
        const double wc = aK * mr_pow04(q);
        WC[i] = wc; IWC[i] = 1.0 / wc;
    }
    MR_SYNC();
    const double NOX = DBL_MAX * 2.0;                  // +inf: "no crossing"
    // crossing point of particles IW-1 and IW (kwt_route.f90:1308-1319); 1/WC is cached per particle
    auto cross = [&](int IW) -> double {
        const int JW = IW - 1;
        if (WC[IW] == 0.0 || WC[JW] == 0.0) return NOX;
        const double WDIFF = IWC[JW] - IWC[IW];
        if (WDIFF == 0.0) return NOX;
        if (WC[IW] == WC[JW]) return NOX;
        return (T1[IW] - T1[JW]) / WDIFF;
    };
    if (NN > 1) {                                      // breaking waves, kwt_route.f90:1301-1349
        MR_NOUNROLL
        for (int IW = 2 + lane; IW <= NN; IW += MR_NL) XX[IW] = cross(IW);
        MR_SYNC();
        double X = 0.0;
        MR_NOUNROLL
        for (;;) {
            // serial scan: XB = XMX; for IW: if (XXB < X || XXB > XB) skip; else XB = XXB, IXB = IW  => minimum, last on ties
            double XB = XMX; int IXB = 0;
            MR_NOUNROLL
            for (int IW = 2 + lane; IW <= NN; IW += MR_NL) {
                const double XXB = XX[IW];
                if (XXB < X || XXB > XB) continue;
                XB = XXB; IXB = IW;
            }
            team_argmin_last(XB, IXB);
            if (XB == XMX) break;
            NN = NN - 1;
            const int JXB = IXB - 1;
            const double q2n = fmax(Q2[JXB], Q2[IXB]), q1n = fmin(Q1[JXB], Q1[IXB]);
            // the two mr_pow() of a merge run on two lanes
            double A2, A1;
            if (MR_NL > 1) {
                const double a = mr_pow(((lane & 1) ? q1n : q2n) / K, p1);
                A2 = team_bcast(a, 0); A1 = team_bcast(a, 1);
            } else {
                A2 = mr_pow(q2n / K, p1); A1 = mr_pow(q1n / K, p1);
            }
            const double CM = (q2n - q1n) / (A2 - A1);
            const double t1n = T1[JXB] + XB / WC[JXB] - XB / CM;
            const int ixb0 = IX[IXB];
            // drop particle IXB: shift IXB+1.. down by one, a chunk of MR_NL entries at a time (read, sync, write)
            MR_NOUNROLL
            for (int j = ixb0 + lane; j <= NI; j += MR_NL) MF[j] = (signed char)(MF[j] - 1);
            MR_NOUNROLL
            for (int base = IXB; base <= NN; base += MR_NL) {
                const int i = base + lane;
                double sT1 = 0.0, sWC = 0.0, sIWC = 0.0, sQ1 = 0.0, sQ2 = 0.0, sXX = 0.0; signed char sIX = 0;
                if (i <= NN) { sIX = IX[i + 1]; sT1 = T1[i + 1]; sWC = WC[i + 1]; sIWC = IWC[i + 1]; sQ1 = Q1[i + 1]; sQ2 = Q2[i + 1]; sXX = XX[i + 1]; }
                MR_SYNC();
                if (i <= NN) { IX[i] = sIX; T1[i] = sT1; WC[i] = sWC; IWC[i] = sIWC; Q1[i] = sQ1; Q2[i] = sQ2; XX[i] = sXX; }
                MR_SYNC();
            }
            MR_SYNC();
            if (lane == 0) { Q2[JXB] = q2n; Q1[JXB] = q1n; T1[JXB] = t1n; WC[JXB] = CM; IWC[JXB] = 1.0 / CM; }
            MR_SYNC();
            if (lane == 0 && JXB >= 2) XX[JXB] = cross(JXB);
            if (lane == (MR_NL > 1 ? 1 : 0) && IXB <= NN) XX[IXB] = cross(IXB);
            MR_SYNC();
            X = XB;
        }
    }
    // exit times of the (merged) particles, kwt_route.f90:1363-1370
    bool zero = false;
    MR_NOUNROLL
    for (int IR = 1 + lane; IR <= NN; IR += MR_NL) {
        if (WC[IR] < DBL_MIN) zero = true;
        TEX[IR] = fmin(XMX / WC[IR] + T1[IR], DBL_MAX);
    }
    if (team_any(zero)) return 20;                     // zero flow, kwt_route.f90:1365-1368
    MR_SYNC();
    int ierr = 0, ICOUNT = 0;
    if (NN == NI) {
        // No particle merged: particle IR is emitted as entry IR with its own Q and TENTRY (already in place), so
        // only T_EXIT is new.  rUpdate's fix-ups (:1431,1434) fire only where the raw exit times are not strictly
        // increasing (or the first is not after T_START); when no lane sees that, the raw times are final.
        bool viol = false;
        MR_NOUNROLL
        for (int IR = 1 + lane; IR <= NN; IR += MR_NL) viol = viol || (IR == 1 ? TEX[1] <= T_START : TEX[IR] <= TEX[IR - 1]);
        if (!team_any(viol)) {
            unsigned rt = 0;
            MR_NOUNROLL
            for (int IR = 1 + lane; IR <= NN; IR += MR_NL) {
                const double tx = TEX[IR];
                S.TX[IR] = tx;
                if (tx < T_END) rt |= 1u << (IR - 1);
            }
            routed = team_or(rt);
            NQ2 = NN;
            MR_SYNC();
            return 0;
        }
    }
    if (lane == 0) {                                   // order-dependent emission + rUpdate, kwt_route.f90:1371-1437
        int bad = 0;
        auto rupdate = [&](double QNEW, double TOLD, double TNEW) {
            ICOUNT = ICOUNT + 1;
            if (ICOUNT > NQ1) { bad = 1; ICOUNT = NQ1; return; }
            S.Q[ICOUNT] = QNEW; S.TE[ICOUNT] = TOLD; S.TX[ICOUNT] = TNEW;
            if (ICOUNT > 1) { if (S.TX[ICOUNT] <= S.TX[ICOUNT - 1]) S.TX[ICOUNT] = S.TX[ICOUNT - 1] + 1.0; }
            if (ICOUNT == 1 && S.TX[1] <= T_START) S.TX[1] = T_START + 1.0;
            if (S.TX[ICOUNT] < T_END) routed |= 1u << (ICOUNT - 1);
        };
        MR_NOUNROLL
        for (int IR = 1; IR <= NN && !ierr; ++IR) {
            const double TEXIT = TEX[IR];
            const double TNEXT = IR < NN ? TEX[IR + 1] : DBL_MAX;
            if (Q1[IR] != Q2[IR]) {
                if (TEXIT < T_END) {
                    const double TEXIT2 = fmin(TEXIT + 1.0, TEXIT + 0.5 * (fmin(TNEXT, T_END) - TEXIT));
                    if (TEXIT2 == TEXIT) { ierr = 30; break; }
                    rupdate(Q1[IR], T1[IR], TEXIT);
                    rupdate(Q2[IR], T1[IR], TEXIT2);
                } else {
                    MR_NOUNROLL
                    for (int JR = 1; JR <= NI; ++JR) if (MF[JR] == IR) rupdate(Q0[JR], T0[JR], TEXIT);
                }
            } else {
                rupdate(Q1[IR], T1[IR], TEXIT);
            }
        }
        if (!ierr && bad) ierr = 60;
    }
    ierr = team_bcast(ierr, 0);
    NQ2 = team_bcast(ICOUNT, 0);
    routed = team_bcast(routed, 0);
    MR_SYNC();
    return ierr;
}

// ------------------------------------------------------------------------------------------------
// qexmul_rch (kwt_route.f90:619-993): merge the upstream basin series and routed-particle series into one
// particle stream QD/TD, written to S.Q/S.TE[nOwn ..].  Upstream particle rows are read in place from the buffer
// the upstream reaches wrote this step; nothing upstream is modified (the reference's strip, :840-844, is
// applied by the owner when it reads its own state back, see kwt_reach_team).
//
// Series s < NUPB: basin of upstream i, points (T0, QR0), (T1, QR1), width 1.  Series NUPB..: the wave of every
// non-headwater upstream, points k = 0..slen-1.  A series' CANDIDATES are its points k = 1..ncand (routed, and
// not beyond the series end); the sequential merge processes all candidates in (time, series) order, emits one
// particle per distinct time, and interpolates every other series in the bracket [cursor-1, cursor], where
// cursor = 1 + (number of that series' candidates already processed), capped at cmax.
// Returns 0 or -site (identical on all lanes).
// ------------------------------------------------------------------------------------------------
template <class SC>
MR_DEV int kwt_merge_team(const DevNet &d, SC &S, int p, int t, int b, double T0, double T1, int nOwn, int &ND, int &nRead) {
    const int lane = MR_LANE;
    const int N = d.nRch;
    const int u0 = d.upPtr[p], NUPB = d.upPtr[p + 1] - u0;
    const double W = d.rwidth[p];
    const double *qr0 = d.qrSer + (size_t)t * N, *qr1 = d.qrSer + (size_t)(t + 1) * N;
    ND = 0;
    if (2 * NUPB > MAXSER) return -E_TOO_MANY_UPS;
    // series descriptors; upstream i is handled by lane i.  REACH_INFLOW (kwt_route.f90:168-174) is summed here too,
    // while the upstream discharges are in flight with everything else.
    int NUPR = 0, poolN = 2 * NUPB, M = NUPB;
    bool bad = false;
    const double *Qs = d.qSer[M_KWT] + (size_t)t * N;
    const int nGood = d.nGood[p];
    MR_NOUNROLL
    for (int base = 0; base < NUPB; base += MR_NL) {
        const int i = base + lane;
        int U = 0, NS = 0, NR = 0;
        bool isr = false;
        if (i < NUPB) {
            U = d.upIdx[u0 + i];
            const bool reach = d.nGood[U] > 0;
            const double q0 = qr0[U], q1 = qr1[U];
            if (i < nGood) S.TX[1 + i] = Qs[U];                      // parked in T_EXIT(1:), which kinwav fills only later
            if (reach) { isr = true; NS = d.kwN[b][U]; NR = d.kwNR[b][U]; }
            S.u.m.soff[i] = (short)(2 * i); S.u.m.slen[i] = 2; S.u.m.ncand[i] = 1; S.u.m.cmax[i] = 1; S.u.m.cbase[i] = (short)i;
            S.u.m.scf[i] = 1.0 / W;
            S.u.m.sq[2 * i] = q0; S.u.m.st[2 * i] = T0; S.u.m.sq[2 * i + 1] = q1; S.u.m.st[2 * i + 1] = T1;
        }
        if (isr && (NS < 2 || NR < 1)) bad = true;
        const int sl = isr ? (NR + 1 < NS ? NR + 1 : NS) : 0;
        int nc = isr ? (sl - 1 < NR - 1 ? sl - 1 : NR - 1) : 0;
        if (nc < 0) nc = 0;
        // one scan for three prefix sums: series rank (5 bits), pool offset and candidate base (10 bits each, <= 16*22)
        int tot;
        const int ex = team_excl_scan((isr ? 1 : 0) | (sl << 5) | (nc << 15), tot);
        const int r = NUPB + NUPR + (ex & 31), off = poolN + ((ex >> 5) & 1023), cb = M + (ex >> 15);
        NUPR += tot & 31; poolN += (tot >> 5) & 1023; M += tot >> 15;
        if (isr) {
            S.u.m.soff[r] = (short)off; S.u.m.slen[r] = (short)sl; S.u.m.ncand[r] = (short)nc;
            S.u.m.cmax[r] = (short)(NR < sl - 1 ? NR : sl - 1); S.u.m.cbase[r] = (short)cb;
            S.u.m.scf[r] = d.rwidth[U] / W; S.u.m.upos[r] = U;
        }
    }
    MR_SYNC();
    if (lane == 0) {
        double qup = 0.0;
        MR_NOUNROLL
        for (int m = 0; m < nGood; ++m) qup = qup + S.TX[1 + m];
        d.inflow[M_KWT][p] = qup;
    }
    const int NUPS = NUPB + NUPR;
    if (NUPS == 1) {                                   // single headwater upstream, kwt_route.f90:743-759
        if (lane == 0) { S.Q[nOwn] = S.u.m.sq[1] / W; S.TE[nOwn] = T1; }
        ND = 1;
        MR_SYNC();
        return 0;
    }
    if (NUPR == 0) {
        // only headwater basins upstream: the merge emits the single time T1 (series 0 supplies it, the other
        // basins are interpolated at their end point, :930-957)
        MR_NOUNROLL
        for (int s = lane; s < NUPB; s += MR_NL) {         // SFLOW of every basin series, one per lane (the upper half of scf is free here)
            const double qb = S.u.m.sq[2 * s], qe = S.u.m.sq[2 * s + 1];
            double SFLOW;
            if (s == 0) SFLOW = qe * S.u.m.scf[0];
            else { const double SLOPE = (qe - qb) / (T1 - T0); SFLOW = (qb + SLOPE * (T1 - T0)) * S.u.m.scf[s]; }
            S.u.m.scf[MAXSER / 2 + s] = SFLOW;
        }
        MR_SYNC();
        if (lane == 0) {
            double Q_AGG = 0.0;
            MR_NOUNROLL
            for (int s = 0; s < NUPB; ++s) Q_AGG = Q_AGG + S.u.m.scf[MAXSER / 2 + s];
            S.Q[nOwn] = Q_AGG; S.TE[nOwn] = T1;
        }
        ND = 1;
        MR_SYNC();
        return 0;
    }
    if (team_any(bad)) return -E_NO_ROUTED_UP;
    if (M > SC::CAP - nOwn || poolN > SC::PCAP) return -E_SCRATCH;
    MR_SYNC();
    // stage the upstream waves
    MR_NOUNROLL
    for (int s = NUPB; s < NUPS; ++s) {
        const int U = S.u.m.upos[s], o = S.u.m.soff[s], sl = S.u.m.slen[s];
        const double *QF = d.kwQF[b] + (size_t)U * KWP, *TR = d.kwTR[b] + (size_t)U * KWP;
        MR_NOUNROLL
        for (int k = lane; k < sl; k += MR_NL) { S.u.m.sq[o + k] = QF[k]; S.u.m.st[o + k] = TR[k]; }
        nRead += sl;
    }
    MR_SYNC();
    // one candidate per lane
    bool ebrk = false, eord = false;
    MR_NOUNROLL
    for (int base = 0; base < M; base += MR_NL) {
        const int c = base + lane;
        if (c < M) {
            int J = 0;
            MR_NOUNROLL
            while (J < NUPS - 1 && c >= S.u.m.cbase[J] + S.u.m.ncand[J]) ++J;
            const int k = c - S.u.m.cbase[J] + 1;
            const int oJ = S.u.m.soff[J];
            const double CT = S.u.m.st[oJ + k];
            if (k >= 2 && CT < S.u.m.st[oJ + k - 1]) eord = true;
            int ord = 0; bool dup = false, brk = false;
            double Q_AGG = 0.0;
            MR_NOUNROLL
            for (int s = 0; s < NUPS; ++s) {
                const int o = S.u.m.soff[s];
                double SFLOW;
                if (s == J) {
                    ord += k - 1;
                    SFLOW = S.u.m.sq[o + k] * S.u.m.scf[s];
                } else {
                    const int nc = S.u.m.ncand[s];
                    // candidates of series s processed before this one: times < CT, or == CT in an earlier series
                    // (times ascend within a series: linear scan for short series, bisection for long ones)
                    int cnt = 0;
                    if (nc <= 6) {
                        MR_NOUNROLL
                        for (int kk = 1; kk <= nc; ++kk) {
                            const double tt = S.u.m.st[o + kk];
                            if (tt < CT || (tt == CT && s < J)) ++cnt; else break;
                        }
                    } else {
                        int hi = nc;
                        MR_NOUNROLL
                        while (cnt < hi) {
                            const int mid = (cnt + hi + 1) >> 1;
                            const double tt = S.u.m.st[o + mid];
                            if (tt < CT || (tt == CT && s < J)) cnt = mid; else hi = mid - 1;
                        }
                    }
                    if (s < J && cnt >= 1 && S.u.m.st[o + cnt] == CT) dup = true;
                    ord += cnt;
                    int cur = 1 + cnt;
                    if (cur > S.u.m.cmax[s]) cur = S.u.m.cmax[s];
                    const double tb = S.u.m.st[o + cur - 1], te = S.u.m.st[o + cur];
                    const double qb = S.u.m.sq[o + cur - 1], qe = S.u.m.sq[o + cur];
                    if (te < CT || tb > CT) brk = true;
                    const double SLOPE = (qe - qb) / (te - tb);
                    const double PREDV = qb + SLOPE * (CT - tb);
                    SFLOW = PREDV * S.u.m.scf[s];
                }
                Q_AGG = Q_AGG + SFLOW;
            }
            if (brk && !dup) ebrk = true;
            S.Q[nOwn + ord] = Q_AGG; S.TE[nOwn + ord] = CT; S.u.m.flag[ord] = dup ? 0 : 1;
        }
    }
    if (team_any(eord)) return -E_TIME_ORDER;
    if (team_any(ebrk)) return -E_BRACKET;
    MR_SYNC();
    // drop the duplicates (kwt_route.f90:926)
    int pos = 0;
    MR_NOUNROLL
    for (int base = 0; base < M; base += MR_NL) {
        const int i = base + lane;
        const bool keep = i < M && S.u.m.flag[i] != 0;
        double qv = 0.0, tv = 0.0;
        if (keep) { qv = S.Q[nOwn + i]; tv = S.TE[nOwn + i]; }
        MR_SYNC();
        int tot;
        const int r = team_rank(keep, tot);
        if (keep) { S.Q[nOwn + pos + r] = qv; S.TE[nOwn + pos + r] = tv; }
        pos += tot;
        MR_SYNC();
    }
    ND = pos;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// kwt_rch (kwt_route.f90:36-346) for interior reach p at batch step t (absolute step tau).
// State buffers: a reach writes its complete post-step particle array KWAVE(0:NQ2+1) and the number of routed
// entries NR into buffer tau&1.  What the reference removes afterwards -- the downstream reach strips
// KWAVE(0:NR-2) (:840-844), outlets and lake inlets strip themselves (:325-344) -- always leaves
// KWAVE(NR-1:), so the owner simply starts reading at NR-1 next step, and the consumer (exactly one wavefront
// behind, reading the same buffer) sees the unstripped array.
// ------------------------------------------------------------------------------------------------
// WS = true: every lane of the WARP calls this together (teams of MR_TEAM < 32 lanes, one task each; a team without a
// task passes active = false) and the phases of the task -- gather + merge | thinning | routing | averaging + store --
// end in a full-warp sync, so the teams of a warp, which diverge inside a phase whenever their tasks differ,
// re-converge at every phase boundary and share the instruction stream wherever their paths agree.
// EXT = true: the instantiation for batches with water management (extract_from_rch between thinning and routing).
template <class SC, bool WS = false, bool EXT = false>
MR_DEV int kwt_reach_team(const DevNet &d, SC &S, int p, int t, long long tau, double T0, double T1, int *nPre = nullptr, bool active = true) {
    const int lane = MR_LANE;
    const int N = d.nRch;
    const int b = (int)(tau & 1), bp = b ^ 1;
    double *Qs = d.qSer[M_KWT] + (size_t)t * N;
    int rc = 0, nOwn = 0, ND = 0, ND_read = 0, n = 0, NQ2 = 0, NR = 0;
    double qr1 = 0.0, W = 1.0;
    unsigned routed = 0;
    bool live = active;

    // ---- phase 1: getusq_rch (own wave + merged upstream waves), kwt_route.f90:461-613.  false = task finished or failed.
    auto phase_in = [&]() -> bool {
        qr1 = d.qrSer[(size_t)(t + 1) * N + p];
        const int nGood = d.nGood[p];
        if (nGood == 0) {                              // no contributing area upstream, kwt_route.f90:181-205
            if (lane == 0) {
                d.inflow[M_KWT][p] = 0.0;
                Qs[p] = qr1;
                d.kwN[b][p] = 1; d.kwNR[b][p] = 0;
                const size_t row = (size_t)p * KWP;
                d.kwQF[b][row] = -9999.0; d.kwTI[b][row] = -9999.0; d.kwTR[b][row] = -9999.0;
            }
            return false;
        }
        const int u0 = d.upPtr[p];
        W = d.rwidth[p];
        const int nPrev = d.kwN[bp][p], nrPrev = d.kwNR[bp][p];
        const int first = nrPrev > 0 ? nrPrev - 1 : 0;
        nOwn = nPrev > 0 ? nPrev - first : 1;
        if (nPrev > 0) {
            const size_t row = (size_t)p * KWP + first;
            MR_NOUNROLL
            for (int i = lane; i < nOwn; i += MR_NL) { S.Q[i] = d.kwQF[bp][row + i]; S.TE[i] = d.kwTI[bp][row + i]; }
            if (lane == 0) S.TX[0] = d.kwTR[bp][row];
        }
        if (d.flags[p] & FLAG_LAKE_UP) {               // lake outlet reach, kwt_route.f90:540-559
            if (d.upPtr[p + 1] - u0 > 1) { if (lane == 0) raise(d.err, 10, p, E_LAKE_UPS); return false; }
            if (lane == 0) {
                S.Q[nOwn] = Qs[d.upIdx[u0]] / W; S.TE[nOwn] = T1;
                double qup = 0.0;                      // kwt_route.f90:168-174
                MR_NOUNROLL
                for (int m = 0; m < nGood; ++m) qup = qup + Qs[d.upIdx[u0 + m]];
                d.inflow[M_KWT][p] = qup;
            }
            ND = 1;
        } else {
            const int e = kwt_merge_team(d, S, p, t, b, T0, T1, nOwn, ND, ND_read);
            if (e == -E_SCRATCH && SC::CAP < WCAP) { rc = KWT_RETRY; return false; }   // nothing modified yet: re-run with the full scratch
            if (e) {
                const int site = -e;
                if (lane == 0) raise(d.err, site == E_TIME_ORDER ? 30 : (site == E_BRACKET ? 40 : (site == E_STUCK ? 20 : 60)), p, site);
                return false;
            }
        }
        MR_SYNC();
        if (nPrev == 0 && lane == 0) {                 // cold start, kwt_route.f90:587-596
            S.Q[0] = S.Q[nOwn]; S.TE[0] = T0 - (T1 - T0); S.TX[0] = T0;
        }
        MR_SYNC();
        n = nOwn + ND;
        if (nPre) *nPre = n;
        bool neg = false;
        MR_NOUNROLL
        for (int i = lane; i < n; i += MR_NL) if (S.Q[i] < 0.0) neg = true;
        if (team_any(neg)) { if (lane == 0) raise(d.err, 20, p, E_NEG_FLOW); return false; }
        return true;
    };
    if (live) live = phase_in();
    if (WS) MR_WSYNC();

    // ---- phase 2: remove_rch
    if (live && n > MR_MAXQPAR) { if (kwt_thin_team(S, n)) { if (lane == 0) raise(d.err, 60, p, E_THIN); live = false; } }
    if (WS) MR_WSYNC();

    // ---- phase 2b: extract_from_rch (kwt_route.f90:351-455), water management only.  The waves (all but particle 0) are
    // scaled by the share of the step's mean flow -- time average of the (TENTRY, Q) series, the same interp_rch -- that is
    // added (Qtake > 0: this routine reads the sign the other way round than the other schemes) or removed (Qtake < 0,
    // smaller than what is there); otherwise everything becomes MINFLOW (0).  The exit times it sets are redone by kinwav_rch.
    if (EXT) {
        if (live && d.wmFlux) {
            const double Qtake = d.wmFlux[(size_t)t * N + p];
            if (Qtake != -9999.0) {
                double Qavg = 0.0;
                if (kwt_time_average_team(S.TE, S.Q, n, T0, T1, S.u.k.XX, Qavg)) { if (lane == 0) raise(d.err, 40, p, E_INTERP); live = false; }
                else {
                    const double totQ = team_bcast(Qavg, 0) * W;      // the average is valid on lane 0
                    MR_SYNC();
                    if (Qtake > 0.0) { const double Qfrac = Qtake / totQ; MR_NOUNROLL for (int i = 1 + lane; i < n; i += MR_NL) S.Q[i] = S.Q[i] * (1.0 + Qfrac); }
                    else if (Qtake < 0.0 && fabs(Qtake) < totQ) { const double Qfrac = fabs(Qtake) / totQ; MR_NOUNROLL for (int i = 1 + lane; i < n; i += MR_NL) S.Q[i] = S.Q[i] * (1.0 - Qfrac); }
                    else { MR_NOUNROLL for (int i = lane; i < n; i += MR_NL) S.Q[i] = 0.0; }
                    MR_SYNC();
                }
            }
        }
        if (WS) MR_WSYNC();
    }

    // ---- phase 3: kinwav_rch
    if (live) {
        const int ek = kwt_kinwav_team(d, S, p, T0, T1, n - 1, NQ2, routed);
        if (ek) { if (lane == 0) raise(d.err, ek, p, ek == 20 ? E_ZERO_FLOW : (ek == 30 ? E_TEXIT2 : E_RUPDATE)); live = false; }
        else {
            NR = mr_popc(routed);                      // count(FROUTE)-1 (FROUTE(0) is always true)
            if (NR + 1 > NQ2) { if (lane == 0) raise(d.err, 21, p, E_NO_NONROUTED); live = false; }
        }
    }
    if (WS) MR_WSYNC();

    // ---- phase 4: interp_rch, end-of-step point, new wave
    if (live) {
        double QNEW = 0.0;
        if (kwt_time_average_team(S.TX, S.Q, NR + 2, T0, T1, S.u.k.XX, QNEW)) { if (lane == 0) raise(d.err, 40, p, E_INTERP); return rc; }
        // end-of-step point, kwt_route.f90:288-292: flow on lane 0, entry time on lane 1 (one division each)
        double Q_END, TIMEI;
        {
            const int L1 = MR_NL > 1 ? 1 : 0;
            const double *A = (lane == L1 && MR_NL > 1) ? S.TE : S.Q;
            const double v = A[NR] + ((A[NR + 1] - A[NR]) / (S.TX[NR + 1] - S.TX[NR])) * (T1 - S.TX[NR]);
            Q_END = team_bcast(v, 0);
            if (MR_NL > 1) TIMEI = team_bcast(v, L1);
            else TIMEI = S.TE[NR] + ((S.TE[NR + 1] - S.TE[NR]) / (S.TX[NR + 1] - S.TX[NR])) * (T1 - S.TX[NR]);
        }
        if (lane == 0) Qs[p] = QNEW * W + qr1;         // kwt_route.f90:273

        // KWAVE(0:NQ2+1) = routed(0:NR) | end-of-step point | non-routed(NR+1:NQ2), kwt_route.f90:299-311
        const size_t row = (size_t)p * KWP;
        double *oQ = d.kwQF[b] + row, *oI = d.kwTI[b] + row, *oR = d.kwTR[b] + row;
        MR_NOUNROLL
        for (int i = lane; i <= NQ2; i += MR_NL) {
            const int j = i <= NR ? i : i + 1;
            oQ[j] = S.Q[i]; oI[j] = S.TE[i]; oR[j] = S.TX[i];
        }
        if (d.expSlot) {                               // tributary outlet: leave this step's wave for the mainstem domain
            const int slot = d.expSlot[p];
            if (slot >= 0) {
                double *rec = d.expBuf + ((size_t)slot * d.kmax + t) * d.recLen + d.nRoutes + 1;
                MR_NOUNROLL
                for (int i = lane; i <= NQ2; i += MR_NL) {
                    const int j = i <= NR ? i : i + 1;
                    rec[2 + j] = S.Q[i]; rec[2 + KWP + j] = S.TX[i];
                }
                if (lane == 0) { rec[0] = (double)(NQ2 + 2); rec[1] = (double)(NR + 2); rec[2 + NR + 1] = Q_END; rec[2 + KWP + NR + 1] = T1; }
            }
        }
        if (lane == 0) {
            oQ[NR + 1] = Q_END; oI[NR + 1] = TIMEI; oR[NR + 1] = T1;
            d.kwN[b][p] = NQ2 + 2;
            d.kwNR[b][p] = NR + 2;
            if (d.kwCount) d.kwCount[p] += (unsigned)(nOwn + ND_read + NQ2 + 2);
        }
    }
    return rc;
}

}  // namespace mr
