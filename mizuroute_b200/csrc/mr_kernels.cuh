// sm_100a kernels of the reach-routing path.  Compiled with --fmad=false: every a*b+c below is two
// IEEE roundings, as in the reference built without FMA contraction, so SUM / IRF / hillslope-UH
// results are bit-identical to a scalar CPU evaluation in the same operation order.
//
// Layout (all arrays in "stage order", see mr_topo.h; N = nRch):
//   per-reach scalars      x[p]
//   UH windows             x[k*N + p]            (slot-major: thread-per-reach accesses are coalesced)
//   per-step series        x[t*N + p]
//   KWT wave particles     x[p*KWP + k]          (reach-major rows of 192 B, read by the team that routes the reach)
//
// Kernels
//   k_basin       K1 basin2reach (process_remap.f90:372-420) fused with K2 hillslope UH
//                 (basinUH.f90:94-176) for all steps of a batch; the UH window is a ring (no shift copy)
//                 walked once per 64-step chunk
//   k_remap       remap_1D_runoff (process_remap.f90:164-262), optional
//   k_headwater<M>  reaches without upstream reaches, all steps of the batch in one launch
//   k_route<M>    one time-skewed wavefront of route_network (main_route.f90:356-403), thread per (reach, step):
//                 M=0 accum_inst_runoff (accum_runoff.f90:60-75), M=1 irf_rch+conv_upsbas_qr
//                 (irf_route.f90:82-150,235-262), M=3/4/5 the Euler schemes kw_rch / mc_rch / dfw_rch (mr_euler.cuh);
//                 lake reaches branch to lake_route (lake_route.f90:87-229)
//   k_route_kwt_light / _heavy / _team / _range   the same for kwt_rch and callees (kwt_route.f90:36-1622): lane per
//                 (reach, step) for the plain and the thinning tasks (mr_kwt_scalar.cuh), half-warp team per (reach, step) for
//                 the rest and for small wavefronts (mr_kwt.cuh)
//   k_export_pack / k_import_unpack   tributary -> mainstem hand-off records (mpi_process.f90:1238-1329)
#pragma once
#include "mr_dev.h"
#include "mr_kwt.cuh"
#include "mr_kwt_scalar.cuh"
#include "mr_irf.cuh"
#include "mr_euler.cuh"
#include "mr_lake.cuh"
#include "mr_ingest.h"

namespace mr {

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
    double *s_uh = s_dyn + BASIN_TC * BASIN_TPB;                                  // [2][nb] hillslope UH, lake UH
    const int N = d.nRch, nb = d.ntdhBas;
    const int p = blockIdx.x * BASIN_TPB + threadIdx.x;
    for (int k = threadIdx.x; k < nb; k += BASIN_TPB) { s_uh[k] = d.fracFuture[k]; s_uh[nb + k] = k == 0 ? 1.0 : 0.0; }
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
        // A group of BASIN_G consecutive slots is "clean" at a step when none of them is logical slot 0 and their
        // logical indices do not wrap inside the group: then the step is BASIN_G multiply-adds with consecutive UH
        // ordinates and nothing else, and clean steps come in runs (k0 counts down to 1).  Runs are evaluated 8 steps
        // at a time from a sliding window of 15 UH ordinates held in registers; the 8 steps around a wrap take the
        // general path.  Either way every slot sees  v = v + uh[k]*rr[t]  in step order (basinUH.f90:165-176).
        const int head0 = (int)((tau0 + c0) % nb);
        const double *uhp = lake ? s_uh + nb : s_uh;       // lakes: UH = [1, 0, 0, ...] (basinUH.f90:113-116)
        const bool fastOK = nb >= 2 * BASIN_G;
        for (int s0 = 0; s0 < nb; s0 += BASIN_G) {
            const bool full = s0 + BASIN_G <= nb;
            double v[BASIN_G];
#pragma unroll
            for (int g = 0; g < BASIN_G; ++g) v[g] = (s0 + g < nb) ? d.qfutBas[(size_t)(s0 + g) * N + p] : 0.0;
            int t = 0;
            while (t < nc) {
                int k0 = (s0 - head0 - t) % nb; if (k0 < 0) k0 += nb;       // logical index of slot s0 at step t
                if (fastOK && full && k0 >= 1 && k0 + BASIN_G - 1 <= nb - 1) {
                    int nrun = k0 < nc - t ? k0 : nc - t;                    // clean while k0 counts down to 1
                    int kb = k0;
                    const int tend = t + nrun;
                    for (; t + 8 <= tend; t += 8, kb -= 8) {
                        double w[BASIN_G + 7];
#pragma unroll
                        for (int i = 0; i < BASIN_G + 7; ++i) w[i] = uhp[kb - 7 + i];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const double x = s_rr[t + j][threadIdx.x];
#pragma unroll
                            for (int g = 0; g < BASIN_G; ++g) v[g] = v[g] + w[g - j + 7] * x;
                        }
                    }
                    for (; t < tend; ++t, --kb) {
                        const double x = s_rr[t][threadIdx.x];
#pragma unroll
                        for (int g = 0; g < BASIN_G; ++g) v[g] = v[g] + uhp[kb + g] * x;
                    }
                } else {
                    const double x = s_rr[t][threadIdx.x];
#pragma unroll
                    for (int g = 0; g < BASIN_G; ++g) {
                        if (s0 + g < nb) {
                            int kk = k0 + g; if (kk >= nb) kk -= nb;
                            v[g] = v[g] + uhp[kk] * x;
                            if (kk == 0) {            // this slot is BASIN_QR(1) of step t; it re-enters as slot nb-1, empty
                                d.qrSer[(size_t)(c0 + t + 1) * N + p] = v[g];
                                v[g] = 0.0;
                            }
                        }
                    }
                    ++t;
                }
            }
#pragma unroll
            for (int g = 0; g < BASIN_G; ++g) if (s0 + g < nb) d.qfutBas[(size_t)(s0 + g) * N + p] = v[g];
        }
    }
    if (live) d.basinQI[p] = rr;
}

// remap_1D_runoff (process_remap.f90:164-262): runoff on the forcing polygons -> river-network HRUs.  One thread per
// (mapping HRU, step): weighted sum over its overlapping polygons in file order, polygons without forcing skipped, values
// <= -1e-6 skipped, renormalised when the weights used do not sum to one.  Network HRUs absent from the mapping keep 0
// (get_basin_runoff.f90:73).
__global__ void k_remap(const double *forcing, double *out, const int *mapNet, const int *mapPtr, const int *ovIdx, const double *ovW,
                        int nMap, int nForcing, int nHRU, int K) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nMap) return;
    const int j = mapNet[i];
    if (j < 0) return;
    const int a = mapPtr[i], b = mapPtr[i + 1];
    const double xTol = 1.e-6;
    for (int t = blockIdx.y; t < K; t += gridDim.y) {
        const double *f = forcing + (size_t)t * nForcing;
        double sumW = 0.0, r = 0.0;
        for (int m = a; m < b; ++m) {
            const int q = ovIdx[m];
            if (q < 0) continue;
            const double v = f[q];
            if (v > -xTol) { sumW = sumW + ovW[m]; r = r + ovW[m] * v; }
        }
        if (sumW > xTol) { if (fabs(1.0 - sumW) > xTol) r = r / sumW; }
        out[(size_t)t * nHRU + j] = r;
    }
}

// forcing records -> runoff rows of the batch in river-network HRU order (ingest_value, mr_ingest.h); thread per HRU, steps in y
__global__ void k_ingest(const double *rec, double *out, const int *srcOfHru, const int *recPtr, const int *recIdx, const double *recFrac,
                         int nIn, int nHRU, int K, int rescale, double A, double B, double fill) {
    const int hru = blockIdx.x * blockDim.x + threadIdx.x;
    if (hru >= nHRU) return;
    const int src = srcOfHru[hru];
    for (int t = blockIdx.y; t < K; t += gridDim.y)
        out[(size_t)t * nHRU + hru] = ingest_value(rec, nIn, src, recIdx, recFrac, recPtr[t], recPtr[t + 1], rescale, A, B, fill);
}

// reach-level evaporation / precipitation of the lake reaches for the K steps of a batch (main_route.f90:174-199)
__global__ void k_lake_forcing(DevNet d, const int *lakePos, int K) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.nLake * K) return;
    const int slot = i % d.nLake, t = i / d.nLake, p = lakePos[slot];
    d.lakeEvap[(size_t)t * d.nLake + slot] = lake_basin2reach(d, p, d.evapo + (size_t)t * d.nHRU);
    d.lakePrecip[(size_t)t * d.nLake + slot] = lake_basin2reach(d, p, d.precip + (size_t)t * d.nHRU);
}

// gauge observations of the batch -> RCHFLX%Qobs / %Qelapsed rows (da_rows, mr_dev.h); coalesced over the reaches
__global__ void k_da_rows(double *obs, int *el, double *qobsState, int *elState, const unsigned char *hasRecord, int N, int K) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < N) da_rows(obs, el, qobsState, elState, hasRecord, N, p, K);
}

// carry BASIN_QR(1) of the previous batch into row 0 of the series
__global__ void k_carry_qr(double *qrSer, int N, int Kprev) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < N && Kprev > 0) qrSer[p] = qrSer[(size_t)Kprev * N + p];
}

// ------------------------------------------------------------------------------------------------
// per-reach bodies of the three methods (route_network loop body, main_route.f90:372-390)
// ------------------------------------------------------------------------------------------------
template <int M, bool HEAD, bool HY = false>
__device__ __forceinline__ void route_reach(const DevNet &d, int p, int t, long long tau) {
    const int N = d.nRch;
    const int flags = d.flags[p];
    if (flags & FLAG_GHOST) return;
    if (M != M_SUM && (flags & FLAG_LAKE)) { lake_reach<M, HY>(d, p, t, tau); return; }
    if (M == M_KWT && HEAD) {                          // no upstream reach => count(goodBas)=0, kwt_route.f90:181-205
        const int b = (int)(tau & 1);
        d.inflow[M_KWT][p] = 0.0;
        d.qSer[M_KWT][(size_t)t * N + p] = d.qrSer[(size_t)(t + 1) * N + p];
        d.kwN[b][p] = 1; d.kwNR[b][p] = 0;         // the sentinel particle (-9999) is written once by mr_set_network
        return;
    }
    if constexpr (M == M_SUM) {                        // accum_runoff.f90:60-75 (mr_irf.cuh)
        sum_reach(d, p, t);
    } else if constexpr (M == M_IRF) {                 // irf_route.f90:82-150,235-262 (mr_irf.cuh)
        irf_reach<HY>(d, p, t, tau);
    } else if constexpr (M == M_KW || M == M_DW) {     // kwe_route.f90 / dfw_route.f90 (mr_euler.cuh)
        kw_dw_reach<M, HY>(d, p, t);
    } else if constexpr (M == M_MC) {                  // mc_route.f90
        mc_reach<HY>(d, p, t);
    }
}

// headwater reaches (positions [0, nHead)): no upstream dependency, so one thread routes all K steps
template <int M, bool HY = false>
__global__ void __launch_bounds__(256) k_headwater(DevNet d, int K, long long tau0) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= d.nHead) return;
    for (int t = 0; t < K; ++t) route_reach<M, true, HY>(d, p, t, tau0 + t);
}

// one wavefront of interior reaches: positions [lo,hi) hold stages w-K+1..w; the reach at stage s does step t = w - s
template <int M, bool HY = false>
__global__ void __launch_bounds__(256) k_route(DevNet d, int lo, int hi, int w, long long tau0) {
    static_assert(M != M_KWT, "KWT wavefronts run in the k_route_kwt_* kernels");
    const int p = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= hi) return;
    const int t = w - d.stageOf[p];
    route_reach<M, false, HY>(d, p, t, tau0 + t);
}

// per-reach constants of kinwav_rch (kwt_route.f90:1283-1296), evaluated once with the device's sqrt/pow
__global__ void k_kwt_params(int N, const double *rslope, const double *rmann, double *K, double *aK) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const double ALFA = 5.0 / 3.0;
    const double k = sqrt(rslope[p]) / rmann[p];
    K[p] = k;
    aK[p] = ALFA * pow(k, 1.0 / ALFA);
}

// KWT wavefront: one team of MR_TEAM lanes per (reach, step); the wave particles sit in the team's shared-memory
// scratch.  The rare task that needs more room than the shared scratch offers borrows a full-capacity scratch
// from a global arena (64 slots per SM, claimed with an atomic bit mask).
#ifndef KWT_WARPS_N
#define KWT_WARPS_N 4
#endif
constexpr int KWT_WARPS = KWT_WARPS_N;                   // warps per block
constexpr int KWT_TEAMS = KWT_WARPS * (32 / MR_TEAM);    // teams (tasks in flight) per block
constexpr int KWT_ARENA_SMS = 256, KWT_ARENA_SLOTS = 64;
// One (reach, step) task of a KWT wavefront, by one team.  EVERY lane of the warp enters (a team without a task has
// active = false): the first attempt re-converges the teams of the warp at its phase boundaries (kwt_reach_team<.., true>).
template <bool HY>
__device__ __forceinline__ void kwt_task(const DevNet &d, KwtScratchSmall &S, int p, bool active, int w, long long tau0) {
    const int lane = MR_LANE;
    int t = 0;
    if (active) {
        t = w - d.stageOf[p];
        const int flags = d.flags[p];
        if (flags & FLAG_GHOST) {                      // this step's wave of a tributary outlet routed in another domain
            const int b = (int)((tau0 + t) & 1);
            const double *rec = d.impBuf + ((size_t)d.impSlot[p] * d.kmax + t) * d.recLen + d.nRoutes + 1;
            const size_t row = (size_t)p * KWP;
            for (int k = lane; k < KWP; k += MR_NL) { d.kwQF[b][row + k] = rec[2 + k]; d.kwTR[b][row + k] = rec[2 + KWP + k]; }
            if (lane == 0) { d.kwN[b][p] = (int)rec[0]; d.kwNR[b][p] = (int)rec[1]; }
            active = false;
        } else if (flags & FLAG_LAKE) {
            if (lane == 0) lake_reach<M_KWT, HY>(d, p, t, tau0 + t);
            active = false;
        }
    }
    int nPre = 0;
    const long long c0 = d.kwProf ? clock64() : 0;
    const double T0 = active ? d.T0s[t] : 0.0, T1 = active ? d.T1s[t] : 0.0;
    const int rc = kwt_reach_team<KwtScratchSmall, true, HY>(d, S, p, t, tau0 + t, T0, T1, &nPre, active);
    if (!active) return;
    if (d.kwProf && lane == 0) {                       // classes: particles before thinning 0 (no area), <=3, <=6, <=12, <=20, <=40, >40, retry
        const int cls = rc == KWT_RETRY ? 7 : (nPre == 0 ? 0 : nPre <= 3 ? 1 : nPre <= 6 ? 2 : nPre <= 12 ? 3 : nPre <= 20 ? 4 : nPre <= 40 ? 5 : 6);
        atomicAdd(&d.kwProf[2 * cls], (unsigned long long)(clock64() - c0)); atomicAdd(&d.kwProf[2 * cls + 1], 1ull);
    }
    if (rc != KWT_RETRY) return;
    // wide confluence: claim a full-capacity scratch of this SM (this team alone from here on)
    unsigned sm;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    sm &= KWT_ARENA_SMS - 1;
    int slot = 0;
    if (lane == 0) {
        for (;;) {
            const unsigned long long busy = atomicOr(&d.kwArenaMask[sm], 0ull);
            if (~busy == 0ull) continue;
            const int bit = __ffsll((long long)~busy) - 1;
            if (!((atomicOr(&d.kwArenaMask[sm], 1ull << bit) >> bit) & 1ull)) { slot = bit; break; }
        }
    }
    slot = team_bcast(slot, 0);
    KwtScratch &B = reinterpret_cast<KwtScratch *>(d.kwArena)[(size_t)sm * KWT_ARENA_SLOTS + slot];
    kwt_reach_team<KwtScratch, false, HY>(d, B, p, t, tau0 + t, T0, T1);
    MR_SYNC();
    if (lane == 0) { __threadfence(); atomicAnd(&d.kwArenaMask[sm], ~(1ull << slot)); }
}

#ifndef KWT_MIN_BLOCKS
#define KWT_MIN_BLOCKS 8
#endif
// One KWT wavefront = three launches (mr_kwt_scalar.cuh):
//   k_route_kwt_light   one LANE per (reach, step), 32 tasks per warp with their particles in shared-memory columns: routes the
//                       tasks of at most MR_MAXQPAR particles (four out of five); the others go to one of two lists in HBM
//                       (one atomic per warp and list);
//   k_route_kwt_heavy   the same code with room for 48 particles per task and remove_rch: the tasks that must thin;
//   k_route_kwt_team    what is left (wave breaking, lakes, ghosts, water management, exported outlets, wide confluences,
//                       errors), dealt over the whole GPU to teams of MR_TEAM lanes and routed by the cooperative code
//                       (kwt_task -> kwt_reach_team, mr_kwt.cuh).  The loop bounds are the same for all teams of a warp
//                       (required by the full-warp syncs inside kwt_task).
// static per-reach records of the lane-per-task code, once per network
__global__ void k_kws_records(DevNet d, KwsRec *out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < d.nRch) out[p] = kws_make_record(d, p);
}

// append p to a list in HBM if `yes` (all 32 lanes call)
__device__ __forceinline__ void kws_push(bool yes, int *cnt, int *list, int p) {
    const unsigned m = __ballot_sync(0xffffffffu, yes);
    if (m) {
        const int lane = threadIdx.x & 31;
        int base = 0;
        if (lane == __ffs(m) - 1) base = atomicAdd(cnt, __popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        if (yes) list[base + __popc(m & ((1u << lane) - 1u))] = p;
    }
}

#ifndef KWS_MIN_BLOCKS_L
#define KWS_MIN_BLOCKS_L 8
#endif
#ifndef KWS_MIN_BLOCKS_H
#define KWS_MIN_BLOCKS_H 4
#endif
// a block = KWS_WPB warps around ONE set of 32 tasks: warp 0 runs the per-task phases (lane = task), all warps the pool phases
template <bool HY = false>
__global__ void __launch_bounds__(32 * KWS_WPB, KWS_MIN_BLOCKS_L) k_route_kwt_light(DevNet d, int lo, int hi, int w, long long tau0, int *cntH, int *listH, int *cntT, int *listT) {
    __shared__ KwsWarp<KWS_NL, false> S;
    const int p = lo + blockIdx.x * 32 + (threadIdx.x & 31);
    const bool active = p < hi;                        // (every lane enters: the loops inside are warp-uniform)
    const int t = active ? w - d.stageOf[p] : 0;
    const int rc = kws_warp_route<HY, KWS_NL, false>(d, S, p, t, tau0 + t, d.T0s[t], d.T1s[t], active);
    kws_push(rc == KWS_HEAVY, cntH, listH, p);         // (only warp 0 holds tasks)
    kws_push(rc == KWS_TEAM, cntT, listT, p);
}

template <bool HY = false>
__global__ void __launch_bounds__(32 * KWS_WPB, KWS_MIN_BLOCKS_H) k_route_kwt_heavy(DevNet d, const int *cntH, const int *listH, int w, long long tau0, int *cntT, int *listT) {
    __shared__ KwsWarp<KWS_NH, true> S;
    const int cnt = *cntH;
    for (int base = blockIdx.x * 32; base < cnt; base += gridDim.x * 32) {
        const int i = base + (threadIdx.x & 31);
        const bool active = i < cnt;
        const int p = active ? listH[i] : 0;
        const int t = active ? w - d.stageOf[p] : 0;
        const int rc = kws_warp_route<HY, KWS_NH, true>(d, S, p, t, tau0 + t, d.T0s[t], d.T1s[t], active);
        kws_push(rc != KWS_DONE, cntT, listT, p);
    }
}

template <bool HY = false>
__global__ void __launch_bounds__(32 * KWT_WARPS, KWT_MIN_BLOCKS) k_route_kwt_team(DevNet d, const int *deferCnt, const int *deferList, int w, long long tau0) {
    __shared__ KwtScratchSmall S[KWT_TEAMS];
    const int team = threadIdx.x / MR_TEAM;
    const int cnt = *deferCnt;
    for (int base = blockIdx.x * KWT_TEAMS; base < cnt; base += gridDim.x * KWT_TEAMS) {
        const int i = base + team;
        const bool active = i < cnt;
        kwt_task<HY>(d, S[team], active ? deferList[i] : 0, active, w, tau0);
        MR_WSYNC();
    }
}

// a whole (small) wavefront on teams: positions [lo, hi).  Below a few ten thousand tasks a wavefront is bound by the latency
// of its slowest task, not by throughput, and one launch of the cooperative code is the shortest chain.
template <bool HY = false>
__global__ void __launch_bounds__(32 * KWT_WARPS, KWT_MIN_BLOCKS) k_route_kwt_range(DevNet d, int lo, int hi, int w, long long tau0) {
    __shared__ KwtScratchSmall S[KWT_TEAMS];
    const int team = threadIdx.x / MR_TEAM;
    for (int base = lo + blockIdx.x * KWT_TEAMS; base < hi; base += gridDim.x * KWT_TEAMS) {
        const int p = base + team;
        kwt_task<HY>(d, S[team], p, p < hi, w, tau0);
        MR_WSYNC();
    }
}

// ------------------------------------------------------------------------------------------------
// multi-domain hand-off (see include/mizuroute_b200.h)
// ------------------------------------------------------------------------------------------------
// export: REACH_Q of every route, BASIN_QR(1) and -- for outlets without a routed wave -- the sentinel counts
__global__ void k_export_pack(DevNet d, const int *expPos, int nExp, int K) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nExp * K) return;
    const int slot = i / K, t = i - slot * K, p = expPos[slot], N = d.nRch;
    double *rec = d.expBuf + ((size_t)slot * d.kmax + t) * d.recLen;
    for (int m = 0; m < N_METHODS; ++m) if (d.routeSlot[m] >= 0) rec[d.routeSlot[m]] = d.qSer[m][(size_t)t * N + p];
    rec[d.nRoutes] = d.qrSer[(size_t)(t + 1) * N + p];
    if (d.nGood[p] == 0 || d.routeSlot[M_KWT] < 0) { rec[d.nRoutes + 1] = 1.0; rec[d.nRoutes + 2] = 0.0; }
}
// import: the ghosts' REACH_Q and BASIN_QR(1) series for the whole batch (their waves are copied step by step
// inside the KWT wavefronts, kwt_task)
__global__ void k_import_unpack(DevNet d, const int *impPos, int nImp, int K) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nImp * K) return;
    const int slot = i / K, t = i - slot * K, p = impPos[slot], N = d.nRch;
    const double *rec = d.impBuf + ((size_t)slot * d.kmax + t) * d.recLen;
    for (int m = 0; m < N_METHODS; ++m) if (d.routeSlot[m] >= 0) d.qSer[m][(size_t)t * N + p] = rec[d.routeSlot[m]];
    d.qrSer[(size_t)(t + 1) * N + p] = rec[d.nRoutes];
}

// ------------------------------------------------------------------------------------------------
// history aggregation (histVars_data.f90:154-246): period means of one per-step series of a batch, thread per reach.  The sum
// of a period is formed in step order in double precision (the reference's `this%discharge = this%discharge + REACH_Q`),
// divided by the number of steps and rounded to float32; a period open at the end of the batch stays in acc / count.
// ------------------------------------------------------------------------------------------------
__global__ void k_history(const double *rows, double *acc, float *out, const int *pos2rch, int N, int K, int nAgg, int nAcc0, int series, int nSeries, int flush) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const int r = pos2rch[p];
    double a = acc[p];
    int c = nAcc0, per = 0;
    for (int t = 0; t < K; ++t) {
        if (c == 0) a = 0.0;
        a = a + rows[(size_t)t * N + p];
        if (++c == nAgg) { out[((size_t)per * nSeries + series) * N + r] = (float)(a / (double)c); ++per; c = 0; }
    }
    if (flush && c > 0) out[((size_t)per * nSeries + series) * N + r] = (float)(a / (double)c);
    acc[p] = a;
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
