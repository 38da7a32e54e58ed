// Lane abstraction for warp-cooperative device code.
//
// Device build (nvcc): a "team" is MR_TEAM consecutive lanes of a warp (a whole warp or half of one); collectives
// are shuffles / ballots / REDUX restricted to the team's lane mask, so the teams of a warp run independently.
// Host build (g++, tests/emul only): a team is ONE lane, every collective is the identity and every
// `for (i = MR_LANE; i < n; i += MR_NL)` loop visits all items in order.  The host build exists so the
// cooperative algorithms (which are written once, against these primitives) can be checked against the CPU
// oracle without a GPU; it is test infrastructure and is never linked into libmizuroute_b200.so.
#pragma once
#include <cfloat>
#include <cmath>

#if defined(__CUDACC__)
#ifndef MR_TEAM
#define MR_TEAM 16                       // lanes per team: 32 = a whole warp, 16 = two teams per warp (they re-converge at the
                                         // phase boundaries of a task, MR_WSYNC)
#endif
#define MR_DEV __device__ __forceinline__
#define MR_DEV_NOINLINE __device__ __noinline__
#define MR_NL MR_TEAM
#define MR_LANE ((int)(threadIdx.x & (MR_TEAM - 1)))
// lanes of this thread's team within its warp
#define MR_TMASK (MR_TEAM == 32 ? 0xffffffffu : (((1u << (MR_TEAM & 31)) - 1u) << ((threadIdx.x & 31u) & ~(unsigned)(MR_TEAM - 1))))
#define MR_SYNC() __syncwarp(MR_TMASK)
// all 32 lanes of the warp (every team of it): reconvergence point between the phases of a task
#if MR_TEAM < 32
#define MR_WSYNC() __syncwarp()
#else
#define MR_WSYNC() ((void)0)
#endif
#define MR_NOUNROLL _Pragma("unroll 1")
#else
#define MR_DEV inline
#define MR_DEV_NOINLINE inline
#define MR_LANE 0
#define MR_NL 1
#define MR_SYNC() ((void)0)
#define MR_WSYNC() ((void)0)
#define MR_NOUNROLL
#endif

// Whole-warp primitives of the thread-per-task code (mr_kwt_scalar.cuh): every lane routes its own task, and the loops whose
// trip count depends on the task run for the warp's maximum with the finished lanes idling, so that the 32 lanes stay
// converged (a data-dependent loop exit leaves them to run one after the other).  Identity in the host build.
#if defined(__CUDACC__)
#define MR_WARP_ANY(p) (__any_sync(0xffffffffu, (p)) != 0)
#define MR_WARP_MAX(v) ((int)__reduce_max_sync(0xffffffffu, (unsigned)(v)))
#define MR_WARP_SYNC() __syncwarp()
#else
#define MR_WARP_ANY(p) (p)
#define MR_WARP_MAX(v) (v)
#define MR_WARP_SYNC() ((void)0)
#endif

namespace mr {

#if defined(__CUDACC__)
MR_DEV bool team_any(bool pred) { return __any_sync(MR_TMASK, pred) != 0; }
MR_DEV double team_bcast(double v, int src) { return __shfl_sync(MR_TMASK, v, src, MR_TEAM); }
MR_DEV int team_bcast(int v, int src) { return __shfl_sync(MR_TMASK, v, src, MR_TEAM); }
MR_DEV unsigned team_bcast(unsigned v, int src) { return __shfl_sync(MR_TMASK, v, src, MR_TEAM); }
// number of lanes below this one whose pred is true; total = number of lanes with pred true
MR_DEV int team_rank(bool pred, int &total) {
    const unsigned m = __ballot_sync(MR_TMASK, pred) >> ((threadIdx.x & 31u) & ~(unsigned)(MR_TEAM - 1));
    total = __popc(m);
    return __popc(m & ((1u << MR_LANE) - 1u));
}
// exclusive prefix sum of v over lanes; total = sum over all lanes
MR_DEV int team_excl_scan(int v, int &total) {
    int x = v;
#pragma unroll
    for (int o = 1; o < MR_TEAM; o <<= 1) {
        const int y = __shfl_up_sync(MR_TMASK, x, o, MR_TEAM);
        if (MR_LANE >= o) x += y;
    }
    total = __shfl_sync(MR_TMASK, x, MR_TEAM - 1, MR_TEAM);
    return x - v;
}
// Minimum of NON-NEGATIVE doubles (or +inf / NaN, which order above every finite value) with the warp-reduce
// unit: for v >= 0 the IEEE bit pattern orders like the value, so two 32-bit REDUX.MIN give the 64-bit minimum.
MR_DEV double team_min_nonneg(double v, bool &mine) {
    const unsigned long long u = (unsigned long long)__double_as_longlong(v + 0.0);     // -0.0 -> +0.0
    const unsigned hi = (unsigned)(u >> 32), lo = (unsigned)u;
    const unsigned mhi = __reduce_min_sync(MR_TMASK, hi);
    const unsigned mlo = __reduce_min_sync(MR_TMASK, hi == mhi ? lo : 0xffffffffu);
    mine = hi == mhi && lo == mlo;
    return __longlong_as_double((long long)(((unsigned long long)mhi << 32) | mlo));
}
// lexicographic minimum of (v ascending, i ascending), v >= 0: Fortran MINLOC "first minimum"
MR_DEV void team_argmin_first(double &v, int &i) {
    bool mine;
    v = team_min_nonneg(v, mine);
    i = (int)__reduce_min_sync(MR_TMASK, mine ? (unsigned)i : 0xffffffffu);
}
// lexicographic minimum of (v ascending, i DESCENDING), v >= 0: a serial "accept if not greater" scan keeps the last
MR_DEV void team_argmin_last(double &v, int &i) {
    bool mine;
    v = team_min_nonneg(v, mine);
    i = (int)__reduce_max_sync(MR_TMASK, mine ? (unsigned)i : 0u);
}
MR_DEV unsigned team_or(unsigned x) { return __reduce_or_sync(MR_TMASK, x); }
MR_DEV int team_min(int x) { return __reduce_min_sync(MR_TMASK, x); }
MR_DEV int mr_popc(unsigned x) { return __popc(x); }
#else
inline bool team_any(bool pred) { return pred; }
inline double team_bcast(double v, int) { return v; }
inline int team_bcast(int v, int) { return v; }
inline unsigned team_bcast(unsigned v, int) { return v; }
inline int team_rank(bool pred, int &total) { total = pred ? 1 : 0; return 0; }
inline int team_excl_scan(int v, int &total) { total = v; return 0; }
inline void team_argmin_first(double &, int &) {}
inline void team_argmin_last(double &, int &) {}
inline unsigned team_or(unsigned x) { return x; }
inline int team_min(int x) { return x; }
inline int mr_popc(unsigned x) { return __builtin_popcount(x); }
#endif

}  // namespace mr
