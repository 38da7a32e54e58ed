// Warp-of-tasks kinematic-wave-tracking reach step: kwt_rch and callees (kwt_route.f90:36-1622) with one (reach, step) task
// per LANE, 32 tasks per warp.
//
// A task is a short sequential program (the reference's own loop structure) around a lot of independent per-particle
// arithmetic -- a pow(), three divisions per particle, a weighted interpolation per merged particle.  The warp therefore
// alternates between two kinds of phases over the particle columns it keeps in shared memory (element i of task x at
// [i][x]: no bank conflict whatever the index):
//   per task   (lane = task)  the decisions: which upstream series supplies the next particle of qexmul_rch's k-way time
//                             merge and where the other series' cursors stand (times and integers only); thinning
//                             (remove_rch) when the task holds more than MR_MAXQPAR particles; kinwav_rch's exit-time checks,
//                             routed count, interp_rch's time average and the end-of-step point;
//   per particle (pool)       the arithmetic: the particles of all 32 tasks form one pool that is dealt to the lanes 32 at
//                             a time, whatever task they belong to -- merged flows (the interpolations of qexmul_rch),
//                             thinning errors, celerity / exit time / crossing point of kinwav_rch, and the stores of the
//                             new wave.  Neighbouring particles of a task sit on neighbouring lanes (shuffles).
// Every loop whose trip count depends on the task runs for the warp's maximum (MR_WARP_MAX / MR_WARP_ANY, mr_lanes.h) behind
// a warp-wide sync, and there is a single exit: the 32 lanes stay converged.  All 32 lanes call (active = false: no task).
//
// NL = particles a task may hold before routing.  The light instantiation (NL = MR_MAXQPAR, no thinning) takes four tasks
// out of five; the heavy one (NL = 44, THIN) the tasks that must thin.  What is left -- wave breaking (kinwav_rch's shock
// merge), rUpdate's exit-time fix-ups, lakes, ghosts, water management, exported outlets, confluences of more than KWS_BMAX
// reaches, and every anomaly the reference reports as an error -- ends in KWS_TEAM before a result the consumers read has
// been written, and is routed by the team code (kwt_reach_team, mr_kwt.cuh), which also owns every error message.  What
// this path has written by then (part of its own row in the step's particle buffer, REACH_INFLOW) is rewritten there.
// All paths evaluate the same floating-point operations on the same operands in the same order, so which one routes a
// task does not change a bit of the result (tests/test_kwt_emul.py runs them on the host against the oracle).
#pragma once
#include "mr_kwt.cuh"

namespace mr {

constexpr int KWS_NL = MR_MAXQPAR;   // light instantiation: no thinning
#ifndef KWS_NH_N
#define KWS_NH_N 44
#endif
#ifndef KWS_NK_N
#define KWS_NK_N 17
#endif
constexpr int KWS_NK = KWS_NK_N;     // particles the serial kinwav_rch of the heavy instantiation can route (work arrays: see there)
constexpr int KWS_NH = KWS_NH_N;     // heavy instantiation: thinning from up to 44 particles (static shared memory: 48 KB)
constexpr int KWS_WCS = KWS_NH - (MR_MAXQPAR + KWS_NK + 1);      // slots per column left for the celerities of the serial kinwav_rch
static_assert(KWS_WCS >= 1 && 3 * KWS_WCS >= KWS_NK + 1, "the work arrays of the serial kinwav_rch do not fit the heavy columns");
enum { KWS_DONE = 0, KWS_HEAVY = 1, KWS_TEAM = 2 };

#if defined(__CUDACC__)
#ifndef KWS_WPB_N
#define KWS_WPB_N 4
#endif
constexpr int KWS_WNL = 32;
constexpr int KWS_WPB = KWS_WPB_N;     // warps of the block that routes one set of 32 tasks (see kws_warp_route)
#define KWS_LANE ((int)(threadIdx.x & 31u))
#define KWS_WARP ((int)(threadIdx.x >> 5))
#define KWS_BSYNC() __syncthreads()
MR_DEV double kws_up(double v) { return __shfl_up_sync(0xffffffffu, v, 1); }
MR_DEV double kws_last(double v) { return __shfl_sync(0xffffffffu, v, 31); }
MR_DEV int kws_incl_scan(int v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, v, o); if (KWS_LANE >= o) v += y; }
    return v;
}
MR_DEV void kws_count(int *c) { atomicAdd(c, 1); }
MR_DEV void kws_flag(int *f, int v) { atomicMax(f, v); }
MR_DEV double kws_ll2d(long long v) { return __longlong_as_double(v); }
MR_DEV long long kws_d2ll(double v) { return __double_as_longlong(v); }
#else
constexpr int KWS_WNL = 1;
constexpr int KWS_WPB = 1;
#define KWS_LANE 0
#define KWS_WARP 0
#define KWS_BSYNC() ((void)0)
inline double kws_up(double v) { return v; }
inline double kws_last(double v) { return v; }
inline int kws_incl_scan(int v) { return v; }
inline void kws_count(int *c) { ++*c; }
inline void kws_flag(int *f, int v) { if (v > *f) *f = v; }
inline double kws_ll2d(long long v) { double r; memcpy(&r, &v, 8); return r; }
inline long long kws_d2ll(double v) { long long r; memcpy(&r, &v, 8); return r; }
#endif

// the static record of reach p (k_kws_records on the device, the host emulation directly)
MR_DEV KwsRec kws_make_record(const DevNet &d, int p) {
    KwsRec r;
    r.stage = d.stageOf[p]; r.nGood = d.nGood[p];
    const int u0 = d.upPtr[p];
    r.nUps = d.upPtr[p + 1] - u0;
    r.cls = ((d.flags[p] & (FLAG_GHOST | FLAG_LAKE | FLAG_LAKE_UP)) || r.nUps > KWS_BMAX) ? 1 : 0;
    r.W = d.rwidth[p]; r.scfB = 1.0 / r.W; r.aK = d.kwAK[p]; r.XMX = d.rlength[p];
    r.isr = 0; r.pad_ = 0.0;
    for (int s = 0; s < KWS_BMAX; ++s) {
        r.U[s] = 0; r.scf[s] = 0.0;
        if (s < r.nUps && !r.cls) {
            const int u = d.upIdx[u0 + s];
            r.U[s] = u;
            if (d.nGood[u] > 0) { r.isr |= 1 << s; r.scf[s] = d.rwidth[u] / r.W; }
        }
    }
    return r;
}

// shared memory of one warp (= KWS_WNL tasks)
template <int NL, bool THIN>
struct KwsWarp {
    double Q[NL][KWS_WNL], T[NL][KWS_WNL], X[NL][KWS_WNL];      // per particle: flow, entry time, exit time (thinning: error)
    // per task: what the pool phases need to know about the task of a particle
    double aK[KWS_WNL], XMX[KWS_WNL], scfB[KWS_WNL], bq10[KWS_WNL], T0[KWS_WNL], T1[KWS_WNL];
    double bq0[KWS_BMAX][KWS_WNL], bsl[KWS_BMAX][KWS_WNL], scf[KWS_BMAX][KWS_WNL];
    int U[KWS_BMAX][KWS_WNL];
    int p[KWS_WNL], b[KWS_WNL], meta[KWS_WNL], n[KWS_WNL], nOwn[KWS_WNL], NR[KWS_WNL], flag[KWS_WNL];
    int off[KWS_BMAX + 1][KWS_WNL], ncm[KWS_BMAX][KWS_WNL];     // wave series: first candidate slot; candidates | last bracket end << 8
    int pre[KWS_WNL + 1];                                        // pool offsets of the tasks
    unsigned char map[KWS_WNL * NL];                             // task of every pool item
    double pe[THIN ? KWS_WPB : 1][KWS_WNL]; int ps[THIN ? KWS_WPB : 1][KWS_WNL], rem[KWS_WNL];   // thinning: minima of the warps' segments; removals to do
    unsigned char prv[THIN ? NL : 1][KWS_WNL], nxt[THIN ? NL : 1][KWS_WNL];      // thinning: survivors as a doubly linked list
};

// (re)build the pool from the per-lane item counts: offsets of the tasks and the task of every item; returns the pool size
template <int NL, bool THIN>
MR_DEV int kws_pool(KwsWarp<NL, THIN> &S, int count) {
    const int lane = KWS_LANE;
    const int inc = kws_incl_scan(count);
    MR_WARP_SYNC();                                    // the previous pool is no longer read
    if (lane == 0) S.pre[0] = 0;
    S.pre[lane + 1] = inc;
    MR_NOUNROLL
    for (int c = inc - count; c < inc; ++c) S.map[c] = (unsigned char)lane;
    MR_WARP_SYNC();
    return S.pre[KWS_WNL];
}

template <bool EXT, int NL, bool THIN>
MR_DEV int kws_warp_route(const DevNet &d, KwsWarp<NL, THIN> &S, int p, int t, long long tau, double T0, double T1, bool active) {
    const int lane = KWS_LANE;
    const bool w0 = KWS_WARP == 0;                     // the per-task phases are warp 0's; the other warps of the block join for the pool phases
    const int N = d.nRch;
    const int b = (int)(tau & 1), bp = b ^ 1;
    bool live = active && w0;                          // still routing its task on this path
#if defined(__CUDACC__)
    // development profile (MR_KWT_PROFILE=1): cycles of the phases of a block, summed over the blocks [16 + 8 * THIN + phase]
    long long pc = d.kwProf ? clock64() : 0;
#define KWS_PHASE(ph) do { if (d.kwProf && threadIdx.x == 0) { const long long c_ = clock64(); atomicAdd(&d.kwProf[16 + (THIN ? 8 : 0) + (ph)], (unsigned long long)(c_ - pc)); pc = c_; } } while (0)
#else
#define KWS_PHASE(ph) ((void)0)
#endif
    int status = KWS_DONE;
    auto quit = [&](int why) { if (live) { status = why; live = false; } };
    // ---- per task: everything static about the reach comes in one record (KwsRec, written once per network)
    KwsRec R;
    R.cls = 0; R.nGood = 0; R.nUps = 0; R.isr = 0; R.W = 1.0; R.scfB = 1.0; R.aK = 1.0; R.XMX = 1.0;
#pragma unroll
    for (int s = 0; s < KWS_BMAX; ++s) { R.U[s] = 0; R.scf[s] = 0.0; }
    if (live) R = d.kwRec[p];
    if (live && R.cls != 0) quit(KWS_TEAM);            // lake, lake outlet, ghost, more than KWS_BMAX upstream reaches
    if (live && d.expSlot) { if (d.expSlot[p] >= 0) quit(KWS_TEAM); }
    if (EXT) { if (live && d.wmFlux) { if (d.wmFlux[(size_t)t * N + p] != -9999.0) quit(KWS_TEAM); } }
    double *Qs = d.qSer[M_KWT] + (size_t)t * N;
    const double *qr0row = d.qrSer + (size_t)t * N, *qr1row = d.qrSer + (size_t)(t + 1) * N;
    const size_t row = (size_t)(live ? p : 0) * KWP;
    const double W = R.W;
    double qr1 = 0.0;
    int nPrev = 0, nrPrev = 0;
    if (live) {
        qr1 = qr1row[p];
        if (R.nGood == 0) {                            // no contributing area upstream, kwt_route.f90:181-205
            d.inflow[M_KWT][p] = 0.0;
            Qs[p] = qr1;
            d.kwN[b][p] = 1; d.kwNR[b][p] = 0;
            d.kwQF[b][row] = -9999.0; d.kwTI[b][row] = -9999.0; d.kwTR[b][row] = -9999.0;
            live = false;
        }
    }
    if (live) { nPrev = d.kwN[bp][p]; nrPrev = d.kwNR[bp][p]; }
    const int NUPB = live ? R.nUps : 0;
    const int first = nrPrev > 0 ? nrPrev - 1 : 0;
    int nOwn = nPrev > 0 ? nPrev - first : 1;

    // ---- the upstream reaches: basin series (T0, QR0), (T1, QR1) of each, wave series of those with contributing area:
    // points 0 .. sl-1 of the upstream wave, of which 1 .. nc are candidates (routed, not beyond the series end)
    int nc[KWS_BMAX], cmax[KWS_BMAX];
    double bq0[KWS_BMAX], bq1[KWS_BMAX];
    int sumNc = 0, nRead = 0, NUPR = 0;
    double qup = 0.0;                                  // REACH_INFLOW, kwt_route.f90:168-174
    bool noRouted = false;
#pragma unroll
    for (int i = 0; i < KWS_BMAX; ++i) {
        nc[i] = 0; cmax[i] = 0; bq0[i] = 0.0; bq1[i] = 0.0;
        if (i < NUPB) {
            const int u = R.U[i];
            bq0[i] = qr0row[u]; bq1[i] = qr1row[u];
            if (i < R.nGood) qup = qup + Qs[u];
            if (R.isr & (1 << i)) {
                const int NS = d.kwN[b][u], NR = d.kwNR[b][u];
                if (NS < 2 || NR < 1) noRouted = true; // "upstream wave has no routed element": the team code reports it
                else {
                    const int sl = NR + 1 < NS ? NR + 1 : NS;
                    int c = sl - 1 < NR - 1 ? sl - 1 : NR - 1;
                    if (c < 0) c = 0;
                    nc[i] = c; cmax[i] = NR < sl - 1 ? NR : sl - 1;
                    sumNc += c; nRead += sl; ++NUPR;
                }
            }
        }
    }
    if (noRouted) quit(KWS_TEAM);
    // at most one particle per wave candidate before T1 plus the one all series share at T1
    if (live && nOwn + sumNc + 1 > NL) quit(NL < KWS_NH ? KWS_HEAVY : KWS_TEAM);
    if (!THIN) { if (live && nOwn + sumNc + 1 > MR_MAXQPAR) quit(KWS_HEAVY); }
    if (!live) nOwn = 0;

    // ---- own wave KWAVE(NR-1:) of the previous step, kwt_route.f90:461-613
    double TX0 = 0.0;
    {
        const int nCopy = (live && nPrev > 0) ? nOwn : 0;
        const double *oq = d.kwQF[bp] + row + first, *oi = d.kwTI[bp] + row + first;
        const int m = MR_WARP_MAX(nCopy);
        bool neg = false;
#pragma unroll 4
        for (int i = 0; i < m; ++i) if (i < nCopy) { const double q = oq[i]; S.Q[i][lane] = q; S.T[i][lane] = oi[i]; if (q < 0.0) neg = true; }
        if (nCopy) TX0 = d.kwTR[bp][row + first];
        if (neg) quit(KWS_TEAM);                       // "negative flow extracted from upstream reach"
    }
    if (live) d.inflow[M_KWT][p] = qup;
    if (live && nPrev == 0) TX0 = T0;                  // cold start, kwt_route.f90:587-596 (particle 0 itself: below)
    if (live && TX0 > T0) quit(KWS_TEAM);              // interp_rch "bad bounds"

    // ---- qexmul_rch (kwt_route.f90:619-993).  The reference merges the candidates of all series in (time, series) order and
    // emits one particle per distinct time; here every candidate is a pool item that finds its own place in that order.
    // The basin series (lower series numbers) supply the one particle at T1; the wave series end with their end-of-step
    // point at T1, a duplicate of it; the candidates before T1 keep their order (ties between two waves: team code).
    int n = 0;
    bool merging = false;
    if (live) {
        if (NUPB == 1 && NUPR == 0) {                  // single headwater upstream, kwt_route.f90:743-759
            S.Q[nOwn][lane] = bq1[0] / W; S.T[nOwn][lane] = T1;
            n = nOwn + 1;
        } else merging = true;
    }
    if (w0) {                                          // what the pool phases need to know about this lane's task
        S.p[lane] = live ? p : 0; S.b[lane] = b; S.n[lane] = n; S.nOwn[lane] = nOwn; S.flag[lane] = 0; S.NR[lane] = 0;
        S.T0[lane] = T0; S.T1[lane] = T1; S.scfB[lane] = R.scfB; S.bq10[lane] = bq1[0];
        S.aK[lane] = R.aK; S.XMX[lane] = R.XMX;
        int meta = merging ? NUPB : 0, off = 0;
#pragma unroll
        for (int s = 0; s < KWS_BMAX; ++s) {
            if (merging && s < NUPB && (R.isr & (1 << s))) meta |= 4 << s;
            S.U[s][lane] = R.U[s]; S.bq0[s][lane] = bq0[s]; S.scf[s][lane] = R.scf[s];
            S.bsl[s][lane] = (merging && s < NUPB) ? (bq1[s] - bq0[s]) / (T1 - T0) : 0.0;  // SLOPE of the basin series: the same operands at every emission
            S.off[s][lane] = off; S.ncm[s][lane] = nc[s] | (cmax[s] << 8);
            off += merging ? nc[s] : 0;
        }
        S.off[KWS_BMAX][lane] = off;
        S.meta[lane] = meta;
    }
    KWS_PHASE(0);
    // pool of the candidates (+ the particle at T1) of every merging task, built by warp 0; every warp takes a contiguous share
    if (w0) kws_pool(S, merging ? sumNc + 1 : 0);
    KWS_BSYNC();
    KWS_PHASE(1);
    int total = S.pre[KWS_WNL];
    int share = ((total + KWS_WPB - 1) / KWS_WPB + KWS_WNL - 1) / KWS_WNL * KWS_WNL;
    int c0 = KWS_WARP * share, c1 = c0 + share < total ? c0 + share : total;
    // candidate c of task j: series s, point k of the upstream wave
    auto which = [&](int j, int c, int &s, int &k) {
        s = 0;
#pragma unroll
        for (int q = 1; q < KWS_BMAX; ++q) if (c >= S.off[q][j]) s = q;
        k = c - S.off[s][j] + 1;
    };
    // stage the candidate times in the exit-time column (free until kinwav)
    MR_NOUNROLL
    for (int k0 = c0; k0 < c1; k0 += KWS_WNL) {
        const int kk = k0 + lane;
        if (kk < c1) {
            const int j = S.map[kk], c = kk - S.pre[j];
            if (c < S.off[KWS_BMAX][j]) { int s, k; which(j, c, s, k); S.X[c][j] = d.kwTR[S.b[j]][(size_t)S.U[s][j] * KWP + k]; }
        }
    }
    KWS_BSYNC();
    // place and flow of every candidate
    MR_NOUNROLL
    for (int k0 = c0; k0 < c1; k0 += KWS_WNL) {
        const int kk = k0 + lane;
        if (kk < c1) {
            const int j = S.map[kk], c = kk - S.pre[j], nCand = S.off[KWS_BMAX][j];
            const bool atT1 = c == nCand;              // the particle the basins supply at T1
            const double t0 = S.T0[j], t1 = S.T1[j], sB = S.scfB[j];
            const int meta = S.meta[j], nb = meta & 3, bj = S.b[j];
            int sJ = -1, kJ = 1;
            double CT = t1;
            bool odd = false;
            if (!atT1) {
                which(j, c, sJ, kJ); CT = S.X[c][j];
                if (CT > t1) odd = true;
                if (kJ >= 2 && !(S.X[c - 1][j] < CT)) odd = true;                 // "expect process in order of time"
            }
            if (atT1 || CT < t1) {
                int rank = atT1 ? 0 : kJ - 1;
                double Q_AGG = 0.0;
#pragma unroll
                for (int s = 0; s < KWS_BMAX; ++s) {
                    if (s < nb) {
                        double SFLOW;
                        if (atT1 && s == 0) SFLOW = S.bq10[j] * sB;
                        else { const double PREDV = S.bq0[s][j] + S.bsl[s][j] * (CT - t0); SFLOW = PREDV * sB; }
                        Q_AGG = Q_AGG + SFLOW;
                    }
                }
#pragma unroll
                for (int s = 0; s < KWS_BMAX; ++s) {
                    if (meta & (4 << s)) {
                        const double *QF = d.kwQF[bj] + (size_t)S.U[s][j] * KWP, *TR = d.kwTR[bj] + (size_t)S.U[s][j] * KWP;
                        double SFLOW;
                        if (s == sJ) SFLOW = QF[kJ] * S.scf[s][j];
                        else {
                            // candidates of series s before this one in (time, series) order; its bracket follows
                            const int o = S.off[s][j], ncs = S.ncm[s][j] & 255, cm = S.ncm[s][j] >> 8;
                            int cnt = 0;               // (times ascend within a series: lower bound of CT)
                            {
                                int hi = ncs;
                                MR_NOUNROLL
                                while (cnt < hi) { const int mid = (cnt + hi) >> 1; if (S.X[o + mid][j] < CT) cnt = mid + 1; else hi = mid; }
                                if (!atT1 && cnt < ncs) { if (S.X[o + cnt][j] == CT) odd = true; }     // two waves tie: team code
                            }
                            rank += cnt;
                            const int cu = 1 + cnt > cm ? cm : 1 + cnt;
                            const double tb = TR[cu - 1], te = TR[cu], qb = QF[cu - 1], qe = QF[cu];
                            if (te < CT || tb > CT) odd = true;
                            const double SLOPE = (qe - qb) / (te - tb);
                            const double PREDV = qb + SLOPE * (CT - tb);
                            SFLOW = PREDV * S.scf[s][j];
                        }
                        Q_AGG = Q_AGG + SFLOW;
                    }
                }
                const int at = S.nOwn[j] + rank;
                S.Q[at][j] = Q_AGG; S.T[at][j] = CT;
                if (atT1) S.n[j] = at + 1;
                if (Q_AGG < 0.0) odd = true;           // "negative flow extracted from upstream reach"
            }
            if (odd) kws_flag(&S.flag[j], 3);
        }
    }
    KWS_BSYNC();
    KWS_PHASE(2);
    if (live) n = S.n[lane];
    if (live && S.flag[lane]) quit(KWS_TEAM);
    if (live && nPrev == 0) { S.Q[0][lane] = S.Q[nOwn][lane]; S.T[0][lane] = T0 - (T1 - T0); }     // cold start
    if (live) { if (S.Q[0][lane] < 0.0) quit(KWS_TEAM); }
    if (!live) n = 0;

    // ---- kinwav_rch (kwt_route.f90:1130-1439) for one particle of the pool: celerity, exit time, whether it would catch up
    // with the particle before it inside the reach (:1308-1319, a shock: wave breaking is the team code's), whether rUpdate
    // would have to move its exit time (:1431-1434), routed = exit before T_END.  Neighbouring particles of a task sit on
    // neighbouring lanes; the last item of the previous round (its lane 31) is carried over.
    double cWc = 0.0, cIwc = 0.0, cTe = 0.0, cX = 0.0;
    auto kin = [&](int k0) {
        const int k = k0 + lane;
        const bool has = k < c1;
        int j = 0, i = 1;
        double q = 1.0, e = 0.0, aK = 1.0, XMX = 1.0;
        if (has) { j = S.map[k]; i = 1 + (k - S.pre[j]); e = S.T[i][j]; aK = S.aK[j]; XMX = S.XMX[j]; q = S.Q[i][j]; }
        const double wc = aK * mr_pow04(q);
        const double iwc = 1.0 / wc;
        const double x = fmin(XMX / wc + e, DBL_MAX);
        double wcp = kws_up(wc), iwcp = kws_up(iwc), ep = kws_up(e), xp = kws_up(x);
        if (lane == 0) {
            if (k0 == c0 && has && i >= 2) {           // first item of this warp's share: its neighbour is another warp's
                const double qn = S.Q[i - 1][j];
                ep = S.T[i - 1][j]; wcp = aK * mr_pow04(qn); iwcp = 1.0 / wcp; xp = fmin(XMX / wcp + ep, DBL_MAX);
            } else { wcp = cWc; iwcp = cIwc; ep = cTe; xp = cX; }
        }
        cWc = kws_last(wc); cIwc = kws_last(iwc); cTe = kws_last(e); cX = kws_last(x);
        if (has) {
            const bool bad = q < 0.0 || wc < DBL_MIN;  // "negative flow", "zero flow": reported by the team code
            bool giveUp = false;
            if (i >= 2 && wc != 0.0 && wcp != 0.0) {
                const double WDIFF = iwcp - iwc;
                if (WDIFF != 0.0 && wc != wcp) {
                    const double XXB = (e - ep) / WDIFF;
                    if (!(XXB < 0.0 || XXB > XMX) && XXB != XMX) giveUp = true;
                }
            }
            const double t1 = S.T1[j];
            if (i == 1 ? x <= S.T0[j] : x <= xp) giveUp = true;
            const bool routed = x < t1;
            if (!routed && !(x >= t1)) giveUp = true;
            S.X[i][j] = x;
            if (bad) kws_flag(&S.flag[j], 3);
            else if (giveUp) kws_flag(&S.flag[j], 2);  // wave breaking or an exit time rUpdate has to move: the serial kinwav_rch below
            else if (routed) kws_count(&S.NR[j]);
        }
    };
    // pool of the particles 1 .. n-1 of every task
    if (w0) { S.n[lane] = n; kws_pool(S, n > 0 ? n - 1 : 0); }
    KWS_BSYNC();
    total = S.pre[KWS_WNL];
    share = ((total + KWS_WPB - 1) / KWS_WPB + KWS_WNL - 1) / KWS_WNL * KWS_WNL;
    c0 = KWS_WARP * share; c1 = c0 + share < total ? c0 + share : total;
    if (THIN) {
        // ---- remove_rch (kwt_route.f90:999-1123): greedy removal of the particle with the smallest interpolation error
        // until MR_MAXQPAR remain.  Errors of all interior particles from the pool; the removals per task: the errors stay in
        // their slots (a removed particle's is DBL_MAX like the two ends', so a strict < scan over the slots finds the
        // "first minimum" among the survivors), the survivors' neighbours are a doubly linked list.
        const bool thin = live && n > MR_MAXQPAR;
        const int last = n - 1;
        MR_NOUNROLL
        for (int k0 = c0; k0 < c1; k0 += KWS_WNL) {
            const int k = k0 + lane;
            if (k < c1) {
                const int j = S.map[k], i = 1 + (k - S.pre[j]);
                if (S.n[j] > MR_MAXQPAR && i < S.n[j] - 1)
                    S.X[i][j] = fabs((S.Q[i - 1][j] + ((S.Q[i + 1][j] - S.Q[i - 1][j]) / (S.T[i + 1][j] - S.T[i - 1][j])) * (S.T[i][j] - S.T[i - 1][j])) - S.Q[i][j]);
            }
        }
        KWS_BSYNC();
        if (thin) {
            S.X[0][lane] = DBL_MAX; S.X[last][lane] = DBL_MAX;
            MR_NOUNROLL
            for (int i = 0; i < n; ++i) { S.prv[i][lane] = (unsigned char)(i - 1); S.nxt[i][lane] = (unsigned char)(i + 1); }
        }
        auto terr = [&](int a, int m, int c) {         // |INTERP(T(m), Q(a), Q(c), T(a), T(c)) - Q(m)|, :1054,1062,1114-1121
            return fabs((S.Q[a][lane] + ((S.Q[c][lane] - S.Q[a][lane]) / (S.T[c][lane] - S.T[a][lane])) * (S.T[m][lane] - S.T[a][lane])) - S.Q[m][lane]);
        };
        // every warp of the block scans one segment of the slots of every task (lane = task), warp 0 puts the segments'
        // minima together in slot order and removes
        if (w0) S.rem[lane] = thin ? n - MR_MAXQPAR : 0;
        KWS_BSYNC();
        const int myRem = S.rem[lane], myLast = S.n[lane] - 1;
        const int nRemove = MR_WARP_MAX(myRem);
        constexpr int SEG = (NL - 2 + KWS_WPB - 1) / KWS_WPB;
        const int lo = 1 + KWS_WARP * SEG, hi = lo + SEG < myLast ? lo + SEG : myLast;
        bool stuck = false;
        MR_NOUNROLL
        for (int r = 0; r < nRemove; ++r) {
            double emin = DBL_MAX; int sel = -1;
            if (r < myRem) {
#pragma unroll 4
                for (int i = lo; i < hi; ++i) { const double e = S.X[i][lane]; if (e < emin) { emin = e; sel = i; } }
            }
            S.pe[KWS_WARP][lane] = emin; S.ps[KWS_WARP][lane] = sel;
            KWS_BSYNC();
            if (thin && !stuck && r < myRem) {
                emin = DBL_MAX; sel = -1;
#pragma unroll
                for (int q = 0; q < KWS_WPB; ++q) { const double e = S.pe[q][lane]; if (e < emin) { emin = e; sel = S.ps[q][lane]; } }
                if (sel < 0) stuck = true;             // "no interior particle to remove": the team code reports it
                else {
                    const int a = S.prv[sel][lane], c = S.nxt[sel][lane];
                    if (a > 0) S.X[a][lane] = terr(S.prv[a][lane], a, c);
                    if (c < last) S.X[c][lane] = terr(a, c, S.nxt[c][lane]);
                    S.X[sel][lane] = DBL_MAX;
                    S.nxt[a][lane] = (unsigned char)c; S.prv[c][lane] = (unsigned char)a;
                }
            }
            KWS_BSYNC();
        }
        if (thin && stuck) quit(KWS_TEAM);
        if (thin && live) {                            // compact the survivors (positions only move left)
            int pos = 0;
            MR_NOUNROLL
            for (int i = 0; i <= last; i = S.nxt[i][lane]) { S.Q[pos][lane] = S.Q[i][lane]; S.T[pos][lane] = S.T[i][lane]; ++pos; }
            n = pos;
        }
        if (!live) n = 0;
    }
    KWS_PHASE(3);
    if (THIN) {                                        // the pool again: thinned tasks are shorter
        if (w0) { S.n[lane] = n; kws_pool(S, n > 0 ? n - 1 : 0); }
        KWS_BSYNC();
        total = S.pre[KWS_WNL];
        share = ((total + KWS_WPB - 1) / KWS_WPB + KWS_WNL - 1) / KWS_WNL * KWS_WNL;
        c0 = KWS_WARP * share; c1 = c0 + share < total ? c0 + share : total;
    }
    KWS_PHASE(4);
    MR_NOUNROLL
    for (int k0 = c0; k0 < c1; k0 += KWS_WNL) { MR_WARP_SYNC(); kin(k0); }
    KWS_BSYNC();
    KWS_PHASE(5);
    if (live && S.flag[lane] == 3) quit(KWS_TEAM);
    // ---- kinwav_rch in full (kwt_route.f90:1130-1439) for the few tasks whose waves break or whose exit times rUpdate has
    // to move: the reference's serial algorithm, one lane per task.  Its work arrays live in the free upper part of the
    // heavy instantiation's columns, which limits it to KWS_NK particles; the light instantiation passes such tasks on.
    if (live && S.flag[lane] == 2) {
        if (!THIN) quit(KWS_HEAVY);
        else if (n - 1 > KWS_NK) quit(KWS_TEAM);
        else {
            const int NI = n - 1;
            int NN = NI;
            auto rT1 = [&](int i) -> double & { return S.Q[MR_MAXQPAR + i][lane]; };
            auto rQ1 = [&](int i) -> double & { return S.T[MR_MAXQPAR + i][lane]; };
            auto rQ2 = [&](int i) -> double & { return S.X[MR_MAXQPAR + i][lane]; };
            auto rWC = [&](int i) -> double & { const int c = i / KWS_WCS, o = MR_MAXQPAR + KWS_NK + 1 + i % KWS_WCS; return c == 0 ? S.Q[o][lane] : (c == 1 ? S.T[o][lane] : S.X[o][lane]); };
            unsigned char (*IX)[KWS_WNL] = S.prv, (*MF)[KWS_WNL] = S.nxt;
            const double K = d.kwK[p], XMX = R.XMX, p1 = 1.0 / (5.0 / 3.0);
            bool bad = false;
            MR_NOUNROLL
            for (int i = 1; i <= NI; ++i) {
                MF[i][lane] = (unsigned char)i; IX[i][lane] = (unsigned char)i;
                const double q = S.Q[i][lane];
                rQ1(i) = q; rQ2(i) = q; rT1(i) = S.T[i][lane];
                rWC(i) = R.aK * mr_pow04(q);
            }
            if (NN > 1) {                              // breaking waves, :1301-1349
                double X = 0.0;
                MR_NOUNROLL
                for (;;) {
                    double XB = XMX; int IXB = 0;
                    MR_NOUNROLL
                    for (int IW = 2; IW <= NN; ++IW) {
                        const int JW = IW - 1;
                        const double wi = rWC(IW), wj = rWC(JW);
                        if (wi == 0.0 || wj == 0.0) continue;
                        const double WDIFF = 1.0 / wj - 1.0 / wi;
                        if (WDIFF == 0.0) continue;
                        if (wi == wj) continue;
                        const double XXB = (rT1(IW) - rT1(JW)) / WDIFF;
                        if (XXB < X || XXB > XB) continue;
                        XB = XXB; IXB = IW;
                    }
                    if (XB == XMX) break;
                    NN = NN - 1;
                    const int JXB = IXB - 1;
                    const double q2n = fmax(rQ2(JXB), rQ2(IXB)), q1n = fmin(rQ1(JXB), rQ1(IXB));
                    const double A2 = mr_pow(q2n / K, p1), A1 = mr_pow(q1n / K, p1);
                    const double CM = (q2n - q1n) / (A2 - A1);
                    const double t1n = rT1(JXB) + XB / rWC(JXB) - XB / CM;
                    MR_NOUNROLL
                    for (int i = IX[IXB][lane]; i <= NI; ++i) MF[i][lane] = (unsigned char)(MF[i][lane] - 1);
                    MR_NOUNROLL
                    for (int i = IXB; i <= NN; ++i) { IX[i][lane] = IX[i + 1][lane]; rT1(i) = rT1(i + 1); rWC(i) = rWC(i + 1); rQ1(i) = rQ1(i + 1); rQ2(i) = rQ2(i + 1); }
                    rQ2(JXB) = q2n; rQ1(JXB) = q1n; rT1(JXB) = t1n; rWC(JXB) = CM;
                    X = XB;
                }
            }
            // exit times and rUpdate, :1363-1437: the emitted particles overwrite S.Q / S.T / S.X(1 : NQ2) in place (an entry
            // is never written before the originals it stands for have been read)
            int ICOUNT = 0, nRouted = 0;
            auto rupdate = [&](double QNEW, double TOLD, double TNEW) {
                ICOUNT = ICOUNT + 1;
                if (ICOUNT > NI) { bad = true; ICOUNT = NI; return; }
                S.Q[ICOUNT][lane] = QNEW; S.T[ICOUNT][lane] = TOLD;
                if (ICOUNT > 1) { if (TNEW <= S.X[ICOUNT - 1][lane]) TNEW = S.X[ICOUNT - 1][lane] + 1.0; }
                if (ICOUNT == 1 && TNEW <= T0) TNEW = T0 + 1.0;
                S.X[ICOUNT][lane] = TNEW;
                if (TNEW < T1) ++nRouted;
            };
            MR_NOUNROLL
            for (int IR = 1; IR <= NN && !bad; ++IR) {
                const double wc = rWC(IR);
                if (wc < DBL_MIN) { bad = true; break; }                           // "zero flow"
                const double TEXIT = fmin(XMX / wc + rT1(IR), DBL_MAX);
                const double TNEXT = IR < NN ? fmin(XMX / rWC(IR + 1) + rT1(IR + 1), DBL_MAX) : DBL_MAX;
                const double q1 = rQ1(IR), q2 = rQ2(IR), t1 = rT1(IR);
                if (q1 != q2) {
                    if (TEXIT < T1) {
                        const double TEXIT2 = fmin(TEXIT + 1.0, TEXIT + 0.5 * (fmin(TNEXT, T1) - TEXIT));
                        if (TEXIT2 == TEXIT) { bad = true; break; }                // "TEXIT equals TEXIT2 in kinwav"
                        rupdate(q1, t1, TEXIT);
                        rupdate(q2, t1, TEXIT2);
                    } else {
                        MR_NOUNROLL
                        for (int JR = 1; JR <= NI; ++JR) if (MF[JR][lane] == IR) rupdate(S.Q[JR][lane], S.T[JR][lane], TEXIT);
                    }
                } else {
                    rupdate(q1, t1, TEXIT);
                }
            }
            if (bad) quit(KWS_TEAM);
            else { n = ICOUNT + 1; S.n[lane] = n; S.NR[lane] = nRouted; }
        }
    }

    // ---- per task: interp_rch (:1444-1622) over the points (T_EXIT, Q)(0 : NR+1) -- the exit times increase strictly from
    // after T_START on, so its IBEG is 1 and its IEND the first non-routed particle NR+1 -- and the end-of-step point (:288-292)
    const int NR = live ? S.NR[lane] : 0;              // routed particles: 1 .. NR
    if (live && NR + 1 > n - 1) quit(KWS_TEAM);        // "no non-routed particle left"
    {
        double AREAM = 0.0;
        const int mMax = MR_WARP_MAX(live ? NR : 0);
        MR_NOUNROLL
        for (int i = 2; i <= mMax; ++i) if (live && i <= NR) AREAM = AREAM + (S.X[i][lane] - S.X[i - 1][lane]) * 0.5 * (S.Q[i - 1][lane] + S.Q[i][lane]);
        if (live) {
            const int L = NR + 1;
            const double q0 = S.Q[0][lane], q1 = S.Q[1][lane], x1 = S.X[1][lane];
            const double qp = S.Q[NR][lane], ep = S.T[NR][lane], xp = NR > 0 ? S.X[NR][lane] : TX0;
            const double q = S.Q[L][lane], e = S.T[L][lane], x = S.X[L][lane];
            double QNEW;
            const double SLOPE1 = (q1 - q0) / (x1 - TX0);
            if (NR == 0 && T1 < x1) {
                const double QEST0 = SLOPE1 * (T0 - TX0) + q0;
                const double QEST1 = SLOPE1 * (T1 - TX0) + q0;
                QNEW = 0.5 * (QEST0 + QEST1);
            } else {
                double AREAB = 0.0, AREAE = 0.0;
                if (T0 < x1) { const double QEST0 = SLOPE1 * (T0 - TX0) + q0; AREAB = (x1 - T0) * 0.5 * (QEST0 + q1); }
                if (T1 < x) {
                    const double SLOPE = (q - qp) / (x - xp);
                    const double QEST1 = SLOPE * (T1 - xp) + qp;
                    AREAE = (T1 - xp) * 0.5 * (qp + QEST1);
                }
                if (NR > 0) { if (T1 == x && T0 < xp) AREAM = AREAM + (x - xp) * 0.5 * (qp + q); }
                QNEW = (AREAB + AREAE + AREAM) / (T1 - T0);
            }
            const double Q_END = qp + ((q - qp) / (x - xp)) * (T1 - xp);
            const double TIMEI = ep + ((e - ep) / (x - xp)) * (T1 - xp);
            double *oQ = d.kwQF[b] + row, *oI = d.kwTI[b] + row, *oR = d.kwTR[b] + row;
            oQ[L] = Q_END; oI[L] = TIMEI; oR[L] = T1;
            oQ[0] = q0; oI[0] = S.T[0][lane]; oR[0] = TX0;
            Qs[p] = QNEW * W + qr1;                    // kwt_route.f90:273
            d.kwN[b][p] = n + 1;                       // NQ2 + 2
            d.kwNR[b][p] = NR + 2;
            if (d.kwCount) d.kwCount[p] += (unsigned)(nOwn + nRead + n + 1);
        }
    }

    KWS_PHASE(6);
    // ---- pool phase: the new wave KWAVE(0:NQ2+1) = routed(0:NR) | end-of-step point | non-routed (:299-311)
    if (w0 && !live) S.n[lane] = 0;
    KWS_BSYNC();
    MR_NOUNROLL
    for (int k0 = c0; k0 < c1; k0 += KWS_WNL) {
        const int k = k0 + lane;
        if (k < c1) {
            const int j = S.map[k], i = 1 + (k - S.pre[j]);
            if (i < S.n[j]) {
                const int jj = i <= S.NR[j] ? i : i + 1;
                const size_t o = (size_t)S.p[j] * KWP + jj;
                const int bj = S.b[j];
                d.kwQF[bj][o] = S.Q[i][j]; d.kwTI[bj][o] = S.T[i][j]; d.kwTR[bj][o] = S.X[i][j];
            }
        }
    }
    KWS_BSYNC();                                       // (the shared memory is reused by the block's next set of tasks)
    KWS_PHASE(7);
#if defined(__CUDACC__)
    if (d.kwProf && threadIdx.x == 0) atomicAdd(&d.kwProf[32 + (THIN ? 1 : 0)], 1ull);
#endif
    return status;
}

}  // namespace mr
