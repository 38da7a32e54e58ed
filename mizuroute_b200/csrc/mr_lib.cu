// libmizuroute_b200.so -- handle, device memory, launch schedule and the C ABI of include/mizuroute_b200.h.
//
// One handle = one routing domain resident on one B200.  The per-step driver the reference runs as
//   basin2reach -> IRF_route_basin -> for each method: route_network   (main_route.f90:151-266)
// is executed for a whole batch of K time steps as
//   k_basin (all K steps)  ->  for each method, concurrently on its own stream: wavefronts w = 0 .. nStage+K-2 of
//                              k_route<M> / the KWT kernels (k_route_kwt_range, or _light + _heavy + _team)
// where wavefront w holds every (reach, step) pair with stage(reach) + step == w.  Legal because a reach at
// step t needs only its upstream reaches at step t (one stage behind => one wavefront earlier) and itself at
// t-1 (one wavefront earlier); K = 1 degenerates to the reference's upstream->downstream sweep.
#include <cuda_runtime.h>
#include <omp.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "mr_kernels.cuh"
#include "mr_topo.h"
#include "mr_uh.h"
#include "mr_calendar.h"
#include "mr_lakeparams.h"

using namespace mr;

struct mr_handle_s {
    mr_options opt{};
    bool on[N_METHODS] = {false, false, false, false, false, false};
    bool hasNet = false;
    Topology topo;
    std::vector<double> fracFuture, uhHost;      // uhHost slot-major [maxtdh][N], stage order
    std::vector<int> ntdh, flags;                // stage order
    std::vector<int> segIdCopy;                  // caller order
    int maxtdh = 1, ntdhBas = 1;
    DevNet d{};
    std::vector<void *> allocs;
    size_t devBytes = 0;
    long long stepsDone = 0;
    int lastK = 0;
    int launchesLast = 0;
    double timing[8] = {0};
    cudaStream_t stream = nullptr, ownStream = nullptr;
    cudaEvent_t ev[10] = {nullptr};
    cudaStream_t kwSide = nullptr; cudaEvent_t kwEv[2] = {nullptr, nullptr};   // KWT: team kernel of the special tasks beside the heavy pass
    cudaStream_t aux[N_METHODS - 1] = {nullptr, nullptr, nullptr, nullptr, nullptr};    // the routing methods of route_opt are independent: all but the last run on these
    cudaEvent_t mev[N_METHODS][2] = {};          // start / end of each method
    unsigned *dKwCount = nullptr;
    int *dKwDeferCnt = nullptr, *dKwDeferList = nullptr;   // KWT: per-wavefront count and list of the tasks the thread-per-task kernel deferred
    double *dRunoff = nullptr, *dT0s = nullptr, *dT1s = nullptr, *dOut = nullptr;
    int *dRch2pos = nullptr, *dPos2rch = nullptr;
    // history aggregation on the device (mr_history_means): sums of the open period per series, steps in them, staging of the means
    double *dHistAcc = nullptr; float *dHistOut = nullptr; size_t histOutCap = 0; int histCount = 0;
    size_t basinSmem = 0;
    int kwtGridMax = 148 * 8, kwsGridMax = 148 * 5;      // one resident wave of k_route_kwt_team / k_route_kwt_heavy blocks
    // runoff remapping (mr_set_remap): forcing arrives on nForcing polygons, k_remap fills dRunoffNet [max_batch][nHRU]
    int nForcing = 0, nMap = 0;
    double *dRunoffNet = nullptr, *dOvW = nullptr;
    int *dMapNet = nullptr, *dMapPtr = nullptr, *dOvIdx = nullptr;
    // software pipeline of mr_step_batch_async: copies on their own streams, forcing double-buffered
    cudaStream_t copyIn = nullptr, copyOut = nullptr;
    cudaEvent_t evIn[2] = {nullptr, nullptr}, evFree[2] = {nullptr, nullptr}, evOut = nullptr, evD2H = nullptr;
    double *dRunoffSlot[2] = {nullptr, nullptr};
    bool freeRec[2] = {false, false}, d2hRec = false;
    int asyncSlot = 0;
    // parametric lake models beyond Doll-2003: named per-reach parameters (caller order, consumed by mr_set_network) and the
    // simulation start datetime (mr_set_sim_start) from which the day of year of every step follows
    std::map<std::string, std::vector<double>> lakeParams;
    bool hasStart = false, hasHype = false, hasH06 = false; int startY = 0, startM = 1, startD = 1, noleap = 0; double startSec = 0.0;
    int *dStepDoy = nullptr; std::vector<int> stepDoyHost;      // [3][max_batch]: day of year, month, day of month
    // water management (mr_upload_wm): per-reach flux / target volume rows of the next batch, stage order; lakes that follow
    // the target volume (lake parameter LakeTargVol)
    int wmSteps = 0, wmJumpStart = 0; bool wmHasFlux = false, wmHasVol = false, wmActive = false, lakeForcingActive = false;
    double *dWmFlux = nullptr, *dWmVol = nullptr; unsigned char *dLakeTargVol = nullptr;
    std::vector<double> wmStage;
    // data assimilation by direct insertion (mr_set_da, mr_upload_obs): options, the rows of the next batch and the running
    // Qobs / Qelapsed of every reach (device, stage order)
    int qmodOption = 0, qBlendPeriod = 10, qErrTrend = 1, obsSteps = 0; bool daActive = false;
    double *dDaQobs = nullptr, *dQobsState = nullptr, *dQerr[N_METHODS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int *dDaEl = nullptr, *dElState = nullptr; unsigned char *dHasRecord = nullptr;
    std::vector<unsigned char> hasRecordHost;
    // lake forcing (mr_upload_lake_forcing): HRU-level rows of the next batch and their reach-level values at the lake reaches
    int nLake = 0, lakeForcingSteps = 0;
    int *dLakePos = nullptr;
    double *dEvapo = nullptr, *dPrecip = nullptr, *dLakeEvap = nullptr, *dLakePrecip = nullptr;
    // forcing ingest on the device (mr_set_ingest, mr_ingest_records)
    int ingestCols = 0, ingestRescale = 0; double ingestA = 1.0, ingestB = 0.0, ingestFill = -9999.0;
    int *dIngestSrc = nullptr, *dIngestPtr = nullptr, *dIngestIdx = nullptr; double *dIngestRec = nullptr, *dIngestFrac = nullptr;
    size_t ingestRecCap = 0, ingestIdxCap = 0;
    // multi-domain hand-off
    std::vector<int> ghostSegId, ghostKind; std::vector<double> ghostTotArea, ghostWidth;   // consumed by mr_set_network
    int nGhost = 0, nExport = 0;
    int *dExpPos = nullptr, *dImpPos = nullptr, *dExpSlot = nullptr, *dImpSlot = nullptr;
    double *xbuf[2] = {nullptr, nullptr};        // [0] export, [1] import
    bool xowned[2] = {false, false};
};

namespace {

void put_msg(char *message, const std::string &s) {
    if (!message) return;
    std::snprintf(message, MR_STRLEN, "%s", s.c_str());
}
int fail(char *message, int ierr, const std::string &s) { put_msg(message, s); return ierr; }

#define CU(call)                                                                                             \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess) return fail(message, 90, std::string(where) + "/CUDA: " + cudaGetErrorString(e_)); \
    } while (0)

template <typename T>
int dev_alloc(mr_handle h, T **ptr, size_t n, const char *where, char *message, bool zero = true) {
    const size_t bytes = sizeof(T) * (n ? n : 1);
    CU(cudaMalloc((void **)ptr, bytes));
    h->allocs.push_back(*ptr);
    h->devBytes += bytes;
    // The handle's stream is non-blocking: nothing orders it against the legacy stream the synchronous cudaMemcpy calls of
    // the set-up code run on.  The zeroing is therefore complete before dev_alloc returns, so a copy issued next (lake /
    // zero-area sentinels in mr_set_network, the rows of mr_upload_wm) can never be overwritten by a late memset.
    if (zero) { CU(cudaMemsetAsync(*ptr, 0, bytes, h->stream)); CU(cudaStreamSynchronize(h->stream)); }
    return 0;
}
// Every host <-> device copy of the set-up / state code goes through the handle's stream and is complete on return: a plain
// cudaMemcpy runs on the legacy stream, which is not ordered against the (non-blocking) stream the kernels run on, and for
// pageable sources returns once the data is staged, not once it has arrived.
cudaError_t copy_sync(mr_handle h, void *dst, const void *src, size_t bytes, cudaMemcpyKind kind) {
    if (!bytes) return cudaSuccess;
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, kind, h->stream);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(h->stream);
}

template <typename T>
int dev_upload(mr_handle h, T **ptr, const std::vector<T> &v, const char *where, char *message) {
    int e = dev_alloc(h, ptr, v.size(), where, message, false);
    if (e) return e;
    if (!v.empty()) CU(copy_sync(h, *ptr, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
    return 0;
}

template <typename T>
int dev_upload_const(mr_handle h, const T *&field, const std::vector<T> &v, const char *where, char *message) {
    T *tmp = nullptr;
    int e = dev_upload(h, &tmp, v, where, message);
    field = tmp;
    return e;
}

// columns of the caller's runoff array: forcing polygons when a remapping is set, network HRUs otherwise
size_t in_cols(mr_handle h) { return h->nForcing > 0 ? (size_t)h->nForcing : (size_t)h->d.nHRU; }

// after the upload of nSteps rows into `src`: remap to network HRUs if asked, and point the kernels at the result
void stage_runoff(mr_handle h, double *src, int nSteps, cudaStream_t st) {
    if (h->nForcing > 0) {
        dim3 grid((h->nMap + 127) / 128, nSteps < 64 ? nSteps : 64);
        k_remap<<<grid, 128, 0, st>>>(src, h->dRunoffNet, h->dMapNet, h->dMapPtr, h->dOvIdx, h->dOvW, h->nMap, h->nForcing, h->d.nHRU, nSteps);
        h->d.runoff = h->dRunoffNet;
    } else {
        h->d.runoff = src;
    }
}

const char *site_text(int site) {
    switch (site) {
        case E_NEG_RUNOFF: return "basin2reach/exceeded negative runoff tolerance";
        case E_LAKE_UPS: return "kwt_rch/getusq_rch/lake outlet reach should have one upstream lake";
        case E_NEG_FLOW: return "kwt_rch/negative flow extracted from upstream reach";
        case E_SCRATCH: return "kwt_rch/qexmul_rch/particle scratch capacity exceeded";
        case E_STUCK: return "kwt_rch/getusq_rch/qexmul_rch/stuck in the continuous do-loop";
        case E_TIME_ORDER: return "kwt_rch/getusq_rch/qexmul_rch/expect process in order of time";
        case E_BRACKET: return "kwt_rch/getusq_rch/qexmul_rch/the times are not ordered as we assume";
        case E_QD_BOUNDS: return "kwt_rch/getusq_rch/qexmul_rch/QD_TEMP bounds exceeded";
        case E_ZERO_FLOW: return "kwt_rch/kinwav_rch/zero flow";
        case E_TEXIT2: return "kwt_rch/kinwav_rch/TEXIT equals TEXIT2 in kinwav";
        case E_RUPDATE: return "kwt_rch/kinwav_rch/RUPDATE/array bounds exceeded";
        case E_NO_NONROUTED: return "kwt_rch/no non-routed particle left";
        case E_INTERP: return "kwt_rch/interp_rch/bad bounds";
        case E_LAKE_TYPE: return "lake_route/unable to identify the parametric lake model type";
        case E_LAKE_PARAM: return "lake_route/parameters of the lake model are not set (mr_set_lake_param)";
        case E_NO_CALENDAR: return "lake_route/the lake model needs the simulation start datetime (mr_set_sim_start)";
        case E_TOO_MANY_UPS: return "kwt_rch/qexmul_rch/more upstream series than the kernel merges";
        case E_THIN: return "kwt_rch/remove_rch/no interior particle to remove";
        case E_NO_ROUTED_UP: return "kwt_rch/qexmul_rch/upstream wave has no routed element";
        default: return "unknown";
    }
}

void free_device(mr_handle h) {
    for (void *p : h->allocs) cudaFree(p);
    h->allocs.clear();
    h->devBytes = 0;
    h->dKwCount = nullptr;
    h->dKwDeferCnt = h->dKwDeferList = nullptr;
}

// wavefront w of method M on stream st: every (reach, step) with stage + step == w (a contiguous position range)
template <int M>
void launch_wavefront(mr_handle h, cudaStream_t st, int w, int K, long long tau0) {
    const Topology &T = h->topo;
    const int slo = w - K + 1 > 0 ? w - K + 1 : 0;
    const int shi = w < T.nStage - 1 ? w : T.nStage - 1;
    const int lo = T.stagePtr[slo], hi = T.stagePtr[shi + 1];
    if (hi <= lo) return;                        // stages that hold only headwaters
    if constexpr (M == M_KWT) {
        // lane-per-task pass over the wavefront (k_route_kwt_light), then the tasks that must thin or whose waves break
        // (k_route_kwt_heavy) while, on a side stream, teams route what the light pass found to be special (lakes, ghosts,
        // water management, exported outlets, wide confluences: list T1); last, on teams again, the few the heavy pass gave up
        // on (list T2, mostly empty).  The sizes of the lists are known on the device only: one resident wave of blocks at
        // most, grid-stride.
        const bool ext = h->hasHype || h->hasH06 || h->wmActive || h->lakeForcingActive || h->daActive;
        int *cntH = h->dKwDeferCnt + 3 * w, *cntT1 = cntH + 1, *cntT2 = cntH + 2;
        int *listH = h->dKwDeferList, *listT1 = listH + h->d.nRch, *listT2 = listT1 + h->d.nRch;
        const int gridL = (hi - lo + 31) / 32;
        int gridH = gridL < h->kwsGridMax ? gridL : h->kwsGridMax;
        int gridT = (hi - lo + KWT_TEAMS - 1) / KWT_TEAMS;
        if (gridT > h->kwtGridMax) gridT = h->kwtGridMax;
        // By size (MR_KWT_SMALL / MR_KWT_LARGE: tuning knobs): a small wavefront is one launch of the team code over its
        // positions (latency-bound: the shortest chain wins); in a very large one the thinning tasks go to the teams as well
        // (throughput-bound: the team code routes a thinning task in fewer cycles than a lane does).
        static const int nSmall = std::getenv("MR_KWT_SMALL") ? std::atoi(std::getenv("MR_KWT_SMALL")) : 49152;
        static const int nLarge = std::getenv("MR_KWT_LARGE") ? std::atoi(std::getenv("MR_KWT_LARGE")) : 65536;
        if (hi - lo < nSmall) {
            const int grid = (hi - lo + KWT_TEAMS - 1) / KWT_TEAMS;
            if (ext) k_route_kwt_range<true><<<grid, 32 * KWT_WARPS, 0, st>>>(h->d, lo, hi, w, tau0);
            else k_route_kwt_range<false><<<grid, 32 * KWT_WARPS, 0, st>>>(h->d, lo, hi, w, tau0);
            h->launchesLast++;
            return;
        }
        const bool noHeavy = hi - lo >= nLarge;
        if (noHeavy) { cntH = cntT1; listH = listT1; }
        cudaStream_t side = h->kwSide;
        if (ext) k_route_kwt_light<true><<<gridL, 32 * KWS_WPB, 0, st>>>(h->d, lo, hi, w, tau0, cntH, listH, cntT1, listT1);
        else k_route_kwt_light<false><<<gridL, 32 * KWS_WPB, 0, st>>>(h->d, lo, hi, w, tau0, cntH, listH, cntT1, listT1);
        cudaEventRecord(h->kwEv[0], st);
        cudaStreamWaitEvent(side, h->kwEv[0], 0);
        if (ext) k_route_kwt_team<true><<<gridT, 32 * KWT_WARPS, 0, side>>>(h->d, cntT1, listT1, w, tau0);
        else k_route_kwt_team<false><<<gridT, 32 * KWT_WARPS, 0, side>>>(h->d, cntT1, listT1, w, tau0);
        cudaEventRecord(h->kwEv[1], side);
        if (!noHeavy) {
            if (ext) k_route_kwt_heavy<true><<<gridH, 32 * KWS_WPB, 0, st>>>(h->d, cntH, listH, w, tau0, cntT2, listT2);
            else k_route_kwt_heavy<false><<<gridH, 32 * KWS_WPB, 0, st>>>(h->d, cntH, listH, w, tau0, cntT2, listT2);
        }
        cudaStreamWaitEvent(st, h->kwEv[1], 0);
        if (!noHeavy) {
            const int gridT2 = gridT < 2 * h->kwtGridMax / 8 ? gridT : 2 * h->kwtGridMax / 8;
            if (ext) k_route_kwt_team<true><<<gridT2, 32 * KWT_WARPS, 0, st>>>(h->d, cntT2, listT2, w, tau0);
            else k_route_kwt_team<false><<<gridT2, 32 * KWT_WARPS, 0, st>>>(h->d, cntT2, listT2, w, tau0);
            h->launchesLast++;
        }
        h->launchesLast += 2;
        h->launchesLast++;
    } else {
        if (h->hasHype || h->hasH06 || h->wmActive || h->lakeForcingActive || h->daActive) k_route<M, true><<<(hi - lo + 255) / 256, 256, 0, st>>>(h->d, lo, hi, w, tau0);
        else k_route<M, false><<<(hi - lo + 255) / 256, 256, 0, st>>>(h->d, lo, hi, w, tau0);
    }
    h->launchesLast++;
}

__global__ void k_times(double T0, double dt, int K, double *T0s, double *T1s) {
    // TSEC(1)=TSEC(2); TSEC(2)=TSEC(1)+dt, init_model_data.f90:311-312
    double t0 = T0, t1 = T0 + dt;
    for (int t = 0; t < K; ++t) { T0s[t] = t0; T1s[t] = t1; t0 = t1; t1 = t0 + dt; }
}

int check_ready(mr_handle h, int nSteps, const char *where, char *message) {
    if (!h) return fail(message, 1, std::string(where) + "/null handle");
    if (!h->hasNet) return fail(message, 1, std::string(where) + "/mr_set_network has not been called");
    if (nSteps < 1 || nSteps > h->opt.max_batch)
        return fail(message, 1, std::string(where) + "/nSteps outside [1, max_batch]");
    CU(cudaSetDevice(h->opt.device));
    return 0;
}

// first use of data assimilation on this network: rows of one batch, the running Qobs / Qelapsed (0, init_model_data.f90:404-405)
// and Qerror of the methods that call direct_insertion (IRF, KW, MC, DW)
int ensure_da(mr_handle h, const char *where, char *message) {
    if (h->dDaQobs) return 0;
    const size_t N = (size_t)h->d.nRch, KB = (size_t)h->opt.max_batch;
    int e;
    if ((e = dev_alloc(h, &h->dDaQobs, KB * N, where, message))) return e;
    if ((e = dev_alloc(h, &h->dDaEl, KB * N, where, message))) return e;
    if ((e = dev_alloc(h, &h->dQobsState, N, where, message))) return e;
    if ((e = dev_alloc(h, &h->dElState, N, where, message))) return e;
    if ((e = dev_alloc(h, &h->dHasRecord, KB, where, message))) return e;
    for (int m : {M_IRF, M_KW, M_MC, M_DW})
        if (h->on[m] && (e = dev_alloc(h, &h->dQerr[m], N, where, message))) return e;
    return 0;
}

// device part of a batch: everything between "forcing is in HBM" and "REACH_Q series is in HBM"
int route_device(mr_handle h, int K, double T0, const char *where, char *message) {
    DevNet &d = h->d;
    const int N = d.nRch;
    h->launchesLast = 0;
    if (h->lakeForcingSteps && h->lakeForcingSteps != K) {      // refused before anything is launched
        h->lakeForcingSteps = 0;
        return fail(message, 1, std::string(where) + "/lake forcing was uploaded for a different number of steps");
    }
    if (h->wmSteps && h->wmSteps != K) {
        h->wmSteps = 0;
        return fail(message, 1, std::string(where) + "/water management was uploaded for a different number of steps");
    }
    if (h->obsSteps && h->obsSteps != K) {
        h->obsSteps = 0;
        return fail(message, 1, std::string(where) + "/gauge observations were uploaded for a different number of steps");
    }
    // exchange buffers are checked before anything is launched or consumed: a refusal leaves the state untouched
    if (h->nGhost && !d.impBuf) return fail(message, 1, std::string(where) + "/ghost reaches but no import buffer (mr_set_exchange_buffer)");
    if (h->nExport && !d.expBuf) return fail(message, 1, std::string(where) + "/export reaches but no export buffer (mr_set_exchange_buffer)");
    if (h->qmodOption == 1) { int e = ensure_da(h, where, message); if (e) return e; }
    d.wmFlux = d.wmVol = nullptr; d.volJumpStart = 0; d.lakeTargVol = h->dLakeTargVol;
    h->wmActive = false;
    if (h->wmSteps) {                                   // water management rows uploaded for this batch
        h->wmSteps = 0;
        d.wmFlux = h->wmHasFlux ? h->dWmFlux : nullptr; d.wmVol = h->wmHasVol ? h->dWmVol : nullptr; d.volJumpStart = h->wmJumpStart;
        h->wmActive = true;
    }
    d.stepDoy = d.stepMonth = d.stepDay = nullptr;
    d.lastK = h->lastK; d.noleap = h->noleap;
    if ((h->hasHype || h->hasH06) && h->hasStart) {     // calendar of simDatetime(1) of every step of the batch
        const int KB = h->opt.max_batch;
        h->stepDoyHost.assign((size_t)3 * KB, 0);
        for (int t = 0; t < K; ++t)
            step_calendar(h->startY, h->startM, h->startD, h->startSec, h->noleap != 0, h->opt.dt, h->stepsDone + t,
                          h->stepDoyHost[(size_t)KB + t], h->stepDoyHost[(size_t)2 * KB + t], h->stepDoyHost[t]);
        CU(cudaMemcpyAsync(h->dStepDoy, h->stepDoyHost.data(), sizeof(int) * 3 * KB, cudaMemcpyHostToDevice, h->stream));
        d.stepDoy = h->dStepDoy; d.stepMonth = h->dStepDoy + KB; d.stepDay = h->dStepDoy + 2 * KB;
    }
    CU(cudaEventRecord(h->ev[1], h->stream));
    k_times<<<1, 1, 0, h->stream>>>(T0, h->opt.dt, K, h->dT0s, h->dT1s);
    if (h->lastK > 0 && h->lastK != 0) k_carry_qr<<<(N + 255) / 256, 256, 0, h->stream>>>(d.qrSer, N, h->lastK);
    h->launchesLast += 2;
    k_basin<<<(N + BASIN_TPB - 1) / BASIN_TPB, BASIN_TPB, h->basinSmem, h->stream>>>(d, K, h->stepsDone);
    h->launchesLast++;
    // lake forcing uploaded for this batch: reach-level evaporation / precipitation of the lake reaches (main_route.f90:174-199)
    bool lakeForcing = false;
    d.lakeEvap = d.lakePrecip = nullptr; d.evapo = d.precip = nullptr;
    if (h->lakeForcingSteps) {
        h->lakeForcingSteps = 0;
        if (h->nLake) {
            d.evapo = h->dEvapo; d.precip = h->dPrecip; d.lakeEvap = h->dLakeEvap; d.lakePrecip = h->dLakePrecip;
            k_lake_forcing<<<(h->nLake * K + 255) / 256, 256, 0, h->stream>>>(d, h->dLakePos, K);
            h->launchesLast++;
            lakeForcing = true;
        }
    }
    h->lakeForcingActive = lakeForcing;
    // data assimilation (main_route.f90:125-148): Qobs / Qelapsed of every (reach, step) of the batch; a batch without an
    // upload is a stretch of the gauge file without records
    d.daQobs = nullptr; d.daElapsed = nullptr; h->daActive = false;
    if (h->qmodOption == 1) {
        if (!h->obsSteps) CU(cudaMemsetAsync(h->dHasRecord, 0, (size_t)K, h->stream));
        h->obsSteps = 0;
        k_da_rows<<<(N + 255) / 256, 256, 0, h->stream>>>(h->dDaQobs, h->dDaEl, h->dQobsState, h->dElState, h->dHasRecord, N, K);
        h->launchesLast++;
        d.daQobs = h->dDaQobs; d.daElapsed = h->dDaEl; d.qBlendPeriod = h->qBlendPeriod; d.qErrTrend = h->qErrTrend;
        for (int m = 0; m < N_METHODS; ++m) d.qerr[m] = h->dQerr[m];
        h->daActive = true;
    }
    if (h->nGhost) {
        k_import_unpack<<<(h->nGhost * K + 255) / 256, 256, 0, h->stream>>>(d, h->dImpPos, h->nGhost, K);
        h->launchesLast++;
    }
    CU(cudaEventRecord(h->ev[2], h->stream));
    // The methods of route_opt share only BASIN_QR (read-only here), so they run concurrently: the last one on the
    // handle's stream, the others on auxiliary streams forked after k_basin and joined before the export.
    // Exception: lake_route may cut the evaporation of a lake that runs dry, and the methods routed after it in the same
    // step see the cut value (RCHFLX%basinevapo is shared, lake_route.f90:169-172) -- with lake forcing and LakeInputOption
    // 0 / 2 the methods therefore share one stream, in route_opt order within every wavefront.
    const int hb = (d.nHead + 255) / 256, nr = h->opt.n_routes;
    // The same holds for Hanasaki reservoirs, whose inflow memory, monthly means and release coefficient are per reach.  On
    // one stream, wavefront by wavefront and method by method, every lake sees (step t, method 1), (step t, method 2),
    // (step t+1, method 1), ... -- the reference's order.
    const bool serialMethods = nr > 1 && (h->hasH06 || (lakeForcing && (h->opt.LakeInputOption == 0 || h->opt.LakeInputOption == 2)));
    cudaStream_t st[N_METHODS];
    for (int r = 0; r < nr; ++r) {
        st[r] = (serialMethods || r == nr - 1) ? h->stream : h->aux[r];
        if (st[r] != h->stream) CU(cudaStreamWaitEvent(st[r], h->ev[2], 0));
        CU(cudaEventRecord(h->mev[r][0], st[r]));
        if (hb) {
            const bool hy = h->hasHype || h->hasH06 || h->wmActive || h->lakeForcingActive || h->daActive;  // parametric reservoirs, lake forcing or water management: the instantiation that knows them
            switch (h->opt.route_methods[r]) {
                case M_SUM: if (hy) k_headwater<M_SUM, true><<<hb, 256, 0, st[r]>>>(d, K, h->stepsDone); else k_headwater<M_SUM, false><<<hb, 256, 0, st[r]>>>(d, K, h->stepsDone); break;
                case M_IRF: if (hy) k_headwater<M_IRF, true><<<hb, 256, 0, st[r]>>>(d, K, h->stepsDone); else k_headwater<M_IRF, false><<<hb, 256, 0, st[r]>>>(d, K, h->stepsDone); break;
                case M_KWT: if (hy) k_headwater<M_KWT, true><<<hb, 256, 0, st[r]>>>(d, K, h->stepsDone); else k_headwater<M_KWT, false><<<hb, 256, 0, st[r]>>>(d, K, h->stepsDone); break;
                case M_KW: if (hy) k_headwater<M_KW, true><<<hb, 256, 0, st[r]>>>(d, K, h->stepsDone); else k_headwater<M_KW, false><<<hb, 256, 0, st[r]>>>(d, K, h->stepsDone); break;
                case M_MC: if (hy) k_headwater<M_MC, true><<<hb, 256, 0, st[r]>>>(d, K, h->stepsDone); else k_headwater<M_MC, false><<<hb, 256, 0, st[r]>>>(d, K, h->stepsDone); break;
                case M_DW: if (hy) k_headwater<M_DW, true><<<hb, 256, 0, st[r]>>>(d, K, h->stepsDone); else k_headwater<M_DW, false><<<hb, 256, 0, st[r]>>>(d, K, h->stepsDone); break;
            }
            h->launchesLast++;
        }
    }
    for (int r = 0; r < nr; ++r)
        if (h->opt.route_methods[r] == M_KWT) CU(cudaMemsetAsync(h->dKwDeferCnt, 0, sizeof(int) * 3 * ((size_t)h->topo.nStage + K), st[r]));
    for (int w = 0; w < h->topo.nStage + K - 1; ++w)
        for (int r = 0; r < nr; ++r)
            switch (h->opt.route_methods[r]) {
                case M_SUM: launch_wavefront<M_SUM>(h, st[r], w, K, h->stepsDone); break;
                case M_IRF: launch_wavefront<M_IRF>(h, st[r], w, K, h->stepsDone); break;
                case M_KWT: launch_wavefront<M_KWT>(h, st[r], w, K, h->stepsDone); break;
                case M_KW: launch_wavefront<M_KW>(h, st[r], w, K, h->stepsDone); break;
                case M_MC: launch_wavefront<M_MC>(h, st[r], w, K, h->stepsDone); break;
                case M_DW: launch_wavefront<M_DW>(h, st[r], w, K, h->stepsDone); break;
            }
    for (int r = 0; r < nr; ++r) {
        CU(cudaEventRecord(h->mev[r][1], st[r]));
        if (st[r] != h->stream) CU(cudaStreamWaitEvent(h->stream, h->mev[r][1], 0));
    }
    if (h->nExport) {
        k_export_pack<<<(h->nExport * K + 255) / 256, 256, 0, h->stream>>>(d, h->dExpPos, h->nExport, K);
        h->launchesLast++;
    }
    CU(cudaEventRecord(h->ev[3], h->stream));
    CU(cudaGetLastError());
    h->stepsDone += K;
    h->lastK = K;
    return 0;
}

int check_device_error(mr_handle h, const char *where, char *message) {
    int e[4] = {0, 0, 0, 0};
    CU(cudaMemcpyAsync(e, h->d.err, sizeof(e), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (e[0] != 0) {
        const int rch = (e[1] >= 0 && e[1] < h->topo.nRch) ? h->topo.pos2rch[e[1]] : -1;
        CU(cudaMemsetAsync(h->d.err, 0, sizeof(e), h->stream));
        return fail(message, e[0], std::string(where) + "/main_route/" + site_text(e[2]) + " (reach index " + std::to_string(rch) + ")");
    }
    return 0;
}

void collect_timing(mr_handle h) {
    float ms = 0.f;
    auto span = [&](int a, int b) { ms = 0.f; cudaEventElapsedTime(&ms, h->ev[a], h->ev[b]); return (double)ms; };
    h->timing[1] = span(1, 2);
    h->timing[2] = span(2, 3);
    const int nr = h->opt.n_routes;
    for (int r = 0; r < 3; ++r) h->timing[5 + r] = 0.0;
    for (int r = 0; r < nr && r < 3; ++r) { ms = 0.f; cudaEventElapsedTime(&ms, h->mev[r][0], h->mev[r][1]); h->timing[5 + r] = ms; }   // the first three methods of route_opt
}

}  // namespace

__global__ void k_selftest_pow04(int n, const double *x, double *y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = mr::mr_pow04(x[i]);
}

// ================================================================================================
extern "C" {

int mr_create(const mr_options *opts, mr_handle *out, char *message) {
    const char *where = "mr_create";
    if (!opts || !out) return fail(message, 1, "mr_create/null argument");
    *out = nullptr;
    if (opts->n_routes < 1 || opts->n_routes > N_METHODS) return fail(message, 1, "mr_create/route_opt must name 1-6 methods");
    if (!(opts->dt > 0.0)) return fail(message, 1, "mr_create/dt_qsim must be positive");
    if (opts->max_batch < 1) return fail(message, 1, "mr_create/max_batch must be >= 1");
    mr_handle h = new mr_handle_s();
    h->opt = *opts;
    for (int r = 0; r < opts->n_routes; ++r) {
        const int m = opts->route_methods[r];
        if (m < 0 || m >= N_METHODS || h->on[m]) {   // read_control.f90:583-597
            delete h;
            return fail(message, 81, "mr_create/route_opt: routing method id expect digits 0-5, each at most once");
        }
        h->on[m] = true;
    }
    int nDev = 0;
    cudaError_t ce = cudaGetDeviceCount(&nDev);
    if (ce != cudaSuccess || nDev == 0 || opts->device < 0 || opts->device >= nDev) {
        delete h;
        return fail(message, 90, "mr_create/no usable CUDA device (this library has no CPU path)");
    }
    ce = cudaSetDevice(opts->device);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&h->ownStream, cudaStreamNonBlocking);
    h->stream = h->ownStream;
    for (int i = 0; i < 10 && ce == cudaSuccess; ++i) ce = cudaEventCreate(&h->ev[i]);
    for (int i = 0; i < N_METHODS - 1 && ce == cudaSuccess; ++i) ce = cudaStreamCreateWithFlags(&h->aux[i], cudaStreamNonBlocking);
    for (int i = 0; i < 2 * N_METHODS && ce == cudaSuccess; ++i) ce = cudaEventCreate(&h->mev[i / 2][i % 2]);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&h->kwSide, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && ce == cudaSuccess; ++i) ce = cudaEventCreateWithFlags(&h->kwEv[i], cudaEventDisableTiming);
    if (ce != cudaSuccess) { delete h; CU(ce); }
    put_msg(message, "");
    *out = h;
    return 0;
}

int mr_set_network(mr_handle h, int nRch, int nHRU, const int *segId, const int *downSegId, const int *hruSegId,
                   const double *hruArea, const double *length, const double *slope, const double *width,
                   const double *man_n, const int *islake, const int *lakeModelType, const double *D03_MaxStorage,
                   const double *D03_Coefficient, const double *D03_Power, const double *D03_S0, char *message) {
    const char *where = "mr_set_network";
    if (!h) return fail(message, 1, "mr_set_network/null handle");
    if (nRch < 1 || nHRU < 0 || !segId || !downSegId || !length || !slope || (nHRU > 0 && (!hruSegId || !hruArea)))
        return fail(message, 1, "mr_set_network/missing required network variable");
    CU(cudaSetDevice(h->opt.device));
    free_device(h);
    h->hasNet = false; h->stepsDone = 0; h->lastK = 0;
    const mr_options &o = h->opt;

    std::string terr;
    Topology &T = h->topo;
    T = Topology();
    std::vector<int> gKind; std::vector<double> gArea, gWidth;          // per reach, caller order
    h->nGhost = (int)h->ghostSegId.size();
    std::vector<int> ghostRch(h->nGhost, -1);
    if (h->nGhost) {
        gKind.assign(nRch, 0); gArea.assign(nRch, 0.0); gWidth.assign(nRch, 0.0);
        join_ids(h->ghostSegId.data(), h->nGhost, segId, nRch, ghostRch);
        for (int g = 0; g < h->nGhost; ++g) {
            if (ghostRch[g] < 0) return fail(message, 1, "mr_set_network/ghost reach id is not in the network");
            gKind[ghostRch[g]] = h->ghostKind[g]; gArea[ghostRch[g]] = h->ghostTotArea[g]; gWidth[ghostRch[g]] = h->ghostWidth[g];
        }
    }
    int ierr = build_topology(nRch, nHRU, segId, downSegId, hruSegId, hruArea, T, terr,
                              h->nGhost ? gKind.data() : nullptr, h->nGhost ? gArea.data() : nullptr);
    if (ierr) return fail(message, ierr, "mr_set_network/" + terr);
    const int N = nRch;
    h->segIdCopy.assign(segId, segId + nRch);
    h->nForcing = h->nMap = 0; h->dRunoffNet = nullptr; h->dOvW = nullptr; h->dMapNet = h->dMapPtr = h->dOvIdx = nullptr;
    h->dRunoffSlot[0] = h->dRunoffSlot[1] = nullptr; h->freeRec[0] = h->freeRec[1] = false; h->d2hRec = false; h->asyncSlot = 0;
    h->nExport = 0; h->dExpPos = h->dImpPos = h->dExpSlot = h->dImpSlot = nullptr;
    h->dStepDoy = nullptr; h->hasHype = false; h->hasH06 = false;
    h->wmSteps = 0; h->wmActive = false; h->dWmFlux = h->dWmVol = nullptr; h->dLakeTargVol = nullptr;
    h->obsSteps = 0; h->daActive = false; h->dDaQobs = h->dQobsState = nullptr; h->dDaEl = h->dElState = nullptr; h->dHasRecord = nullptr;
    for (int m = 0; m < N_METHODS; ++m) h->dQerr[m] = nullptr;
    h->ingestCols = 0; h->dIngestSrc = h->dIngestPtr = h->dIngestIdx = nullptr; h->dIngestRec = h->dIngestFrac = nullptr; h->ingestRecCap = h->ingestIdxCap = 0;
    h->nLake = 0; h->lakeForcingSteps = 0; h->dLakePos = nullptr; h->dEvapo = h->dPrecip = h->dLakeEvap = h->dLakePrecip = nullptr;
    for (int w = 0; w < 2; ++w) { h->xbuf[w] = nullptr; h->xowned[w] = false; }          // set again after mr_set_network

    // reach parameters in stage order (process_ntopo.f90:176-187,359-366)
    std::vector<double> rlen(N), rslp(N), rwid(N), rman(N), maxS(N, 0.0), coef(N, 0.0), pw(N, 0.0), s0(N, 0.0);
    std::vector<int> ltype(N, MR_LAKE_DOLL03);
    // channel geometry of the Euler schemes (process_ntopo.f90:174-203): bankfull depth dscale*sqrt(totalArea) with
    // <floodplain> T, else high_depth (globalData.f90:187-189); rectangular main channel; bankfull storage (hydraulic.f90:207-237)
    const bool euler = h->on[M_KW] || h->on[M_MC] || h->on[M_DW];
    std::vector<double> rdep, zside, zfld, rstor;
    if (euler) { rdep.assign(N, 100000.0); zside.assign(N, 0.0); zfld.assign(N, o.floodplainSlope > 0.0 ? o.floodplainSlope : 1000.0); rstor.assign(N, 0.0); }
    h->flags.assign(N, 0);
    for (int p = 0; p < N; ++p) {
        const int r = T.pos2rch[p];
        rlen[p] = length[r];
        rslp[p] = std::fmax(slope[r], 1.e-6);                        // min_slope, public_var.f90:30
        rwid[p] = width ? width[r] : o.wscale * std::sqrt(T.totArea[p]);
        rman[p] = man_n ? man_n[r] : o.mann_n;
        if (h->nGhost && gKind[r]) { h->flags[p] |= FLAG_GHOST; rwid[p] = gWidth[r]; }
        if (euler) {
            if (o.floodplain) rdep[p] = (o.dscale > 0.0 ? o.dscale : (double)0.000045f) * std::sqrt(T.totArea[p]);
            rstor[p] = rdep[p] * (rwid[p] + zside[p] * rdep[p]) * rlen[p];       // flow_area(y = bankDepth) * length
        }
        if (o.is_lake_sim) {
            if (islake && islake[r] == 1) h->flags[p] |= FLAG_LAKE;
            ltype[p] = (!o.lakeRegulate || !lakeModelType) ? MR_LAKE_DOLL03 : lakeModelType[r];
            if (D03_MaxStorage) maxS[p] = D03_MaxStorage[r];
            if (D03_Coefficient) coef[p] = D03_Coefficient[r];
            if (D03_Power) pw[p] = D03_Power[r];
            if (D03_S0) s0[p] = D03_S0[r];
        }
    }
    if (o.is_lake_sim)
        for (int p = 0; p < N; ++p)
            for (int m = T.upPtr[p]; m < T.upPtr[p + 1]; ++m)
                if (h->flags[T.upIdx[m]] & FLAG_LAKE) h->flags[p] |= FLAG_LAKE_UP;

    // unit hydrographs (process_param.f90)
    ierr = build_hillslope_uh(o.dt, o.fshape, o.tscale, h->fracFuture);
    if (ierr) return fail(message, ierr, "mr_set_network/basinUH/cannot identify the maximum number of bins for the tdh");
    h->ntdhBas = (int)h->fracFuture.size();
    h->ntdh.assign(N, 1);
    h->maxtdh = 1;
    h->uhHost.clear();
    if (h->on[M_IRF]) {
        std::vector<double> rowMajor((size_t)N * 240);
        int mx = 1;
#pragma omp parallel for schedule(dynamic, 1024) reduction(max : mx)
        for (int p = 0; p < N; ++p) {
            double *u = &rowMajor[(size_t)p * 240];
            const int n = build_reach_uh(rlen[p], o.dt, o.velo, o.diff, u);
            if (h->flags[p] & FLAG_LAKE) { for (int k = 0; k < n; ++k) u[k] = 0.0; u[0] = 1.0; }   // process_ntopo.f90:496-499
            h->ntdh[p] = n;
            if (n > mx) mx = n;
        }
        h->maxtdh = mx;
        h->uhHost.assign((size_t)mx * N, 0.0);
#pragma omp parallel for schedule(static)
        for (int p = 0; p < N; ++p)
            for (int k = 0; k < h->ntdh[p]; ++k) h->uhHost[(size_t)k * N + p] = rowMajor[(size_t)p * 240 + k];
    }

    // ---- device image
    DevNet &d = h->d;
    d = DevNet();
    d.nRch = N; d.nHRU = nHRU; d.nStage = T.nStage; d.nHead = T.nHead; d.ntdhBas = h->ntdhBas; d.maxtdh = h->maxtdh;
    d.dt = o.dt; d.runoffMin = o.runoffMin; d.tconv = o.time_conv; d.lconv = o.length_conv; d.minLengthRoute = o.min_length_route;
    d.doesBasinRoute = o.doesBasinRoute; d.hwDrain = o.hw_drain_point; d.isLakeSim = o.is_lake_sim; d.lakeInputOption = o.LakeInputOption;
    const int KB = o.max_batch;
    int e = 0;
#define UP(field, vec) do { e = dev_upload_const(h, d.field, vec, where, message); if (e) return e; } while (0)
#define AL(ptr, n) do { e = dev_alloc(h, &(ptr), (size_t)(n), where, message); if (e) return e; } while (0)
    UP(stageOf, T.stageOf); UP(upPtr, T.upPtr); UP(upIdx, T.upIdx); UP(nGood, T.nGood);
    UP(hruPtr, T.hruPtr); UP(hruIdx, T.hruIdx); UP(flags, h->flags); UP(ntdh, h->ntdh); UP(lakeType, ltype);
    UP(hruWgt, T.hruWgt); UP(basArea, T.basArea); UP(rlength, rlen); UP(rslope, rslp); UP(rwidth, rwid); UP(rmann, rman);
    UP(fracFuture, h->fracFuture);
    if (h->on[M_IRF]) UP(uh, h->uhHost);
    UP(d03MaxS, maxS); UP(d03Coef, coef); UP(d03Pow, pw); UP(d03S0, s0);
    if (euler) { UP(rdepth, rdep); UP(sideSlope, zside); UP(fldpSlope, zfld); UP(rstorage, rstor); }
    AL(d.qfutBas, (size_t)h->ntdhBas * N);
    AL(d.qrSer, (size_t)(KB + 1) * N);
    AL(d.basinQI, N);
    for (int m = 0; m < N_METHODS; ++m) {
        if (!h->on[m]) continue;
        AL(d.qSer[m], (size_t)KB * N);
        AL(d.vol0[m], N); AL(d.vol1[m], N); AL(d.inflow[m], N); AL(d.wb[m], N);
        if (n_molecule(m)) { AL(d.mol[m], (size_t)n_molecule(m) * N); AL(d.floodVol[m], N); AL(d.reachEle[m], N); }   // molecule%Q = 0, init_model_data.f90:463-497
    }
    if (h->on[M_IRF]) AL(d.qfutIrf, (size_t)h->maxtdh * N);
    if (h->on[M_KWT]) {
        double *kK = nullptr, *kAK = nullptr;
        AL(kK, N); AL(kAK, N);
        k_kwt_params<<<(N + 255) / 256, 256, 0, h->stream>>>(N, d.rslope, d.rmann, kK, kAK);
        d.kwK = kK; d.kwAK = kAK;
        {
            KwsRec *rec = nullptr;
            AL(rec, N);
            k_kws_records<<<(N + 255) / 256, 256, 0, h->stream>>>(d, rec);
            d.kwRec = rec;
        }
        KwtScratch *arena = nullptr; unsigned long long *amask = nullptr;
        AL(arena, (size_t)KWT_ARENA_SMS * KWT_ARENA_SLOTS); AL(amask, KWT_ARENA_SMS);
        d.kwArena = arena; d.kwArenaMask = amask;
        if (std::getenv("MR_KWT_PROFILE")) { unsigned long long *pr = nullptr; AL(pr, 40); d.kwProf = pr; }
        AL(h->dKwDeferCnt, 3 * ((size_t)T.nStage + KB + 1)); AL(h->dKwDeferList, 3 * (size_t)N);
        for (int b = 0; b < 2; ++b) {
            AL(d.kwN[b], N); AL(d.kwNR[b], N);
            AL(d.kwQF[b], (size_t)KWP * N); AL(d.kwTI[b], (size_t)KWP * N); AL(d.kwTR[b], (size_t)KWP * N);
        }
        // Reaches without contributing upstream area hold the sentinel particle (-9999, not routed; kwt_route.f90:181-205),
        // and so do lake reaches from the start (init_model_data.f90:440-456).  Their rows never change: written once here.
        {
            std::vector<int> n1(N, 0);
            std::vector<double> s9((size_t)KWP * N, 0.0);
            for (int p = 0; p < N; ++p) {
                const bool lake = (h->flags[p] & FLAG_LAKE) != 0;
                if (lake) n1[p] = 1;
                if (lake || T.nGood[p] == 0) s9[(size_t)p * KWP] = -9999.0;
            }
            for (int b = 0; b < 2; ++b) {
                CU(copy_sync(h, d.kwN[b], n1.data(), sizeof(int) * N, cudaMemcpyHostToDevice));
                CU(copy_sync(h, d.kwQF[b], s9.data(), sizeof(double) * s9.size(), cudaMemcpyHostToDevice));
                CU(copy_sync(h, d.kwTI[b], s9.data(), sizeof(double) * s9.size(), cudaMemcpyHostToDevice));
                CU(copy_sync(h, d.kwTR[b], s9.data(), sizeof(double) * s9.size(), cudaMemcpyHostToDevice));
            }
        }
    }
    if (o.is_lake_sim) {                              // lake reaches by position, for the optional lake forcing
        std::vector<int> slot(N, -1), pos;
        for (int p = 0; p < N; ++p) if (h->flags[p] & FLAG_LAKE) { slot[p] = (int)pos.size(); pos.push_back(p); }
        h->nLake = (int)pos.size();
        if (h->nLake) { UP(lakeSlot, slot); e = dev_upload(h, &h->dLakePos, pos, where, message); if (e) return e; }
        d.nLake = h->nLake;
        h->hasHype = false;
        for (int p : pos) if (ltype[p] == MR_LAKE_HYPE) h->hasHype = true;
        const LakeParamLookup par = [&](const std::string &nm) -> const std::vector<double> * {
            auto it = h->lakeParams.find(nm);
            return (it == h->lakeParams.end() || (int)it->second.size() != nRch) ? nullptr : &it->second; };
        if (h->hasHype) {                             // HYP_* by lake slot (dataTypes.f90:202-213)
            std::vector<HypeParams> bySlot;
            const std::string missing = build_hype_params(h->nLake, pos.data(), T.pos2rch.data(), par, bySlot);
            if (!missing.empty()) return fail(message, 20, "mr_set_network/HYPE lakes need the parameter " + missing + " for every reach (mr_set_lake_param)");
            UP(hyp, bySlot);
        }
        h->hasH06 = false;
        for (int p : pos) if (ltype[p] == MR_LAKE_HANASAKI06) h->hasH06 = true;
        if (h->hasH06) {                              // H06_* by lake slot (dataTypes.f90:215-254) + the inflow memory
            std::vector<H06Lake> lk;
            long long memDoubles = 0;
            const std::string missing = build_h06_lakes(h->nLake, pos.data(), T.pos2rch.data(), ltype.data(), o.dt, par, lk, memDoubles);
            if (!missing.empty()) return fail(message, 20, "mr_set_network/Hanasaki lakes need the parameter " + missing + " for every reach (mr_set_lake_param)");
            for (int sIdx = 0; sIdx < h->nLake; ++sIdx) {
                if (ltype[pos[sIdx]] != MR_LAKE_HANASAKI06) continue;
                if (pos[sIdx] < T.nHead && o.n_routes > 1)
                    return fail(message, 20, "mr_set_network/a Hanasaki reservoir without upstream reaches cannot be routed with several methods (its state is shared by them)");
                if (lk[sIdx].memF && lk[sIdx].LFnoleap < 1) return fail(message, 20, "mr_set_network/H06_I_mem_L must cover at least one step");
            }
            H06Lake *dl = nullptr;
            e = dev_upload(h, &dl, lk, where, message); if (e) return e;
            d.h06 = dl;
            AL(d.h06Mem, (size_t)(memDoubles > 0 ? memDoubles : 1));
        }
        if (h->hasHype || h->hasH06) AL(h->dStepDoy, (size_t)3 * KB);
        if (const std::vector<double> *tv = par("LakeTargVol")) {          // NETOPO%LakeTargVol: the lake follows REACH_WM_VOL
            std::vector<unsigned char> f(N, 0);
            for (int p = 0; p < N; ++p) f[p] = (*tv)[T.pos2rch[p]] != 0.0 ? 1 : 0;
            e = dev_upload(h, &h->dLakeTargVol, f, where, message); if (e) return e;
        }
    }
    AL(d.err, 4);
    d.expSlot = d.impSlot = nullptr; d.expBuf = nullptr; d.impBuf = nullptr;
    d.nRoutes = o.n_routes; d.kmax = KB; d.recLen = o.n_routes + 3 + 2 * KWP;
    for (int m = 0; m < N_METHODS; ++m) d.routeSlot[m] = -1;
    for (int r = 0; r < o.n_routes; ++r) d.routeSlot[o.route_methods[r]] = r;
    if (h->nGhost) {
        std::vector<int> slot(N, -1), pos(h->nGhost);
        for (int g = 0; g < h->nGhost; ++g) { pos[g] = T.rch2pos[ghostRch[g]]; slot[pos[g]] = g; }
        e = dev_upload(h, &h->dImpSlot, slot, where, message); if (e) return e;
        e = dev_upload(h, &h->dImpPos, pos, where, message); if (e) return e;
        d.impSlot = h->dImpSlot;
    }
    AL(h->dRunoff, (size_t)KB * (nHRU > 0 ? nHRU : 1));
    AL(h->dT0s, KB); AL(h->dT1s, KB);
    AL(h->dOut, (size_t)o.n_routes * KB * N);
    e = dev_upload(h, &h->dRch2pos, T.rch2pos, where, message); if (e) return e;
    e = dev_upload(h, &h->dPos2rch, T.pos2rch, where, message); if (e) return e;
    h->dHistAcc = nullptr; h->dHistOut = nullptr; h->histOutCap = 0; h->histCount = 0;
    d.runoff = h->dRunoff; d.T0s = h->dT0s; d.T1s = h->dT1s;
#undef UP
#undef AL

    {
        int nSM = 148, perSM = 8;
        cudaDeviceGetAttribute(&nSM, cudaDevAttrMultiProcessorCount, o.device);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_route_kwt_team<false>, 32 * KWT_WARPS, 0) != cudaSuccess || perSM < 1) perSM = 8;
        h->kwtGridMax = nSM * perSM;               // one resident wave of team blocks, grid-stride over the deferred list
        int perSMh = 5;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSMh, k_route_kwt_heavy<false>, 32 * KWS_WPB, 0) != cudaSuccess || perSMh < 1) perSMh = 5;
        h->kwsGridMax = nSM * perSMh;
        // tuning knob (development): MR_KWT_WAVES = resident waves the KWT grid may span; 0 = one block per KWT_TEAMS tasks
        if (const char *ev = std::getenv("MR_KWT_WAVES")) {
            const double wv = std::atof(ev);
            h->kwtGridMax = wv <= 0.0 ? nSM * perSM : (int)(wv * nSM * perSM);
        }
    }
    h->basinSmem = sizeof(double) * ((size_t)BASIN_TC * BASIN_TPB + 2 * (size_t)h->ntdhBas);
    CU(cudaFuncSetAttribute(k_basin, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->basinSmem));
    CU(cudaStreamSynchronize(h->stream));
    h->hasNet = true;
    put_msg(message, "");
    return 0;
}

int mr_upload_runoff(mr_handle h, int nSteps, const double *runoff, char *message) {
    const char *where = "mr_upload_runoff";
    int e = check_ready(h, nSteps, where, message); if (e) return e;
    if (!runoff) return fail(message, 1, "mr_upload_runoff/null runoff");
    CU(cudaMemcpyAsync(h->dRunoff, runoff, sizeof(double) * (size_t)nSteps * in_cols(h), cudaMemcpyHostToDevice, h->stream));
    stage_runoff(h, h->dRunoff, nSteps, h->stream);
    CU(cudaStreamSynchronize(h->stream));
    put_msg(message, "");
    return 0;
}

int mr_set_ingest(mr_handle h, int nForcing, const int *forcingOfHru, double scale, double offset, double fill, char *message) {
    const char *where = "mr_set_ingest";
    int e = check_ready(h, 1, where, message); if (e) return e;
    if (h->nForcing > 0) return fail(message, 1, "mr_set_ingest/the handle remaps its forcing (mr_set_remap); ingest maps forcing HRUs one to one");
    if (nForcing < 1 || !forcingOfHru) return fail(message, 1, "mr_set_ingest/invalid arguments");
    const int nH = h->d.nHRU;
    std::vector<int> src(forcingOfHru, forcingOfHru + nH);
    for (int i = 0; i < nH; ++i) if (src[i] >= nForcing) return fail(message, 1, "mr_set_ingest/forcing column outside the records");
    if (!h->dIngestSrc) { e = dev_alloc(h, &h->dIngestSrc, (size_t)nH, where, message, false); if (e) return e;
                          e = dev_alloc(h, &h->dIngestPtr, (size_t)h->opt.max_batch + 1, where, message, false); if (e) return e; }
    CU(copy_sync(h, h->dIngestSrc, src.data(), sizeof(int) * (size_t)nH, cudaMemcpyHostToDevice));
    // scale_forcing (get_basin_runoff.f90:375-423): -9999 = not given
    h->ingestRescale = (scale != -9999.0 || offset != -9999.0) ? 1 : 0;
    h->ingestA = scale == -9999.0 ? 1.0 : scale; h->ingestB = offset == -9999.0 ? 0.0 : offset; h->ingestFill = fill;
    h->ingestCols = nForcing;
    put_msg(message, "");
    return 0;
}

int mr_ingest_records(mr_handle h, int nSteps, int nRec, const double *records, const int *recPtr, const int *recIdx, const double *recFrac, char *message) {
    const char *where = "mr_ingest_records";
    int e = check_ready(h, nSteps, where, message); if (e) return e;
    if (!h->ingestCols) return fail(message, 1, "mr_ingest_records/mr_set_ingest has not been called");
    if (nRec < 1 || !records || !recPtr || !recIdx) return fail(message, 1, "mr_ingest_records/invalid arguments");
    const size_t nIdx = (size_t)recPtr[nSteps];
    for (int t = 0; t < nSteps; ++t) if (recPtr[t + 1] <= recPtr[t]) return fail(message, 30, "timeMap_sim_forc/a simulation step without forcing record");
    for (size_t j = 0; j < nIdx; ++j) if (recIdx[j] < 0 || recIdx[j] >= nRec) return fail(message, 30, "timeMap_sim_forc/record index outside the uploaded records");
    CU(cudaStreamSynchronize(h->stream));
    const size_t need = (size_t)nRec * h->ingestCols;
    if (need > h->ingestRecCap) { e = dev_alloc(h, &h->dIngestRec, need, where, message, false); if (e) return e; h->ingestRecCap = need; }   // (the smaller one stays until mr_destroy)
    if (nIdx > h->ingestIdxCap) { e = dev_alloc(h, &h->dIngestIdx, nIdx, where, message, false); if (e) return e;
                                  e = dev_alloc(h, &h->dIngestFrac, nIdx, where, message, false); if (e) return e; h->ingestIdxCap = nIdx; }
    CU(cudaMemcpyAsync(h->dIngestRec, records, sizeof(double) * need, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->dIngestPtr, recPtr, sizeof(int) * ((size_t)nSteps + 1), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->dIngestIdx, recIdx, sizeof(int) * nIdx, cudaMemcpyHostToDevice, h->stream));
    if (recFrac) CU(cudaMemcpyAsync(h->dIngestFrac, recFrac, sizeof(double) * nIdx, cudaMemcpyHostToDevice, h->stream));
    dim3 grid((h->d.nHRU + 255) / 256, nSteps < 64 ? nSteps : 64);
    k_ingest<<<grid, 256, 0, h->stream>>>(h->dIngestRec, h->dRunoff, h->dIngestSrc, h->dIngestPtr, h->dIngestIdx, recFrac ? h->dIngestFrac : nullptr,
                                          h->ingestCols, h->d.nHRU, nSteps, h->ingestRescale, h->ingestA, h->ingestB, h->ingestFill);
    h->d.runoff = h->dRunoff;
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    put_msg(message, "");
    return 0;
}

int mr_set_lake_param(mr_handle h, const char *name, int n, const double *values, char *message) {
    if (!h || !name || !values || n < 1) return fail(message, 1, "mr_set_lake_param/null argument");
    static const char *known[] = {"HYP_E_emr", "HYP_E_lim", "HYP_E_min", "HYP_E_zero", "HYP_Qrate_emr", "HYP_Erate_emr", "HYP_Qrate_prim",
                                  "HYP_Qrate_amp", "HYP_Qrate_phs", "HYP_prim_F", "HYP_A_avg", "HYP_Qsim_mode",
                                  "H06_Smax", "H06_alpha", "H06_envfact", "H06_S_ini", "H06_c1", "H06_c2", "H06_exponent", "H06_denominator",
                                  "H06_c_compare", "H06_frac_Sdead", "H06_E_rel_ini", "H06_purpose", "H06_I_mem_F", "H06_D_mem_F", "H06_I_mem_L", "H06_D_mem_L",
                                  "LakeTargVol"};
    bool ok = !std::strncmp(name, "H06_I_", 6) || !std::strncmp(name, "H06_D_", 6);      // H06_I_Jan .. H06_D_Dec (and the mem_* above)
    for (const char *k : known) ok = ok || !std::strcmp(k, name);
    if (!ok) return fail(message, 20, std::string("mr_set_lake_param/unknown or unsupported lake parameter ") + name);
    h->lakeParams[name].assign(values, values + n);
    put_msg(message, "");
    return 0;
}

int mr_set_sim_start(mr_handle h, int year, int month, int day, double secOfDay, int noleap, char *message) {
    if (!h) return fail(message, 1, "mr_set_sim_start/null handle");
    if (month < 1 || month > 12 || day < 1 || day > 31 || secOfDay < 0.0 || secOfDay >= 86400.0) return fail(message, 1, "mr_set_sim_start/invalid date");
    h->hasStart = true; h->startY = year; h->startM = month; h->startD = day; h->startSec = secOfDay; h->noleap = noleap ? 1 : 0;
    put_msg(message, "");
    return 0;
}

int mr_upload_wm(mr_handle h, int nSteps, const double *flux_wm, const double *vol_wm, int volJumpStart, char *message) {
    const char *where = "mr_upload_wm";
    int e = check_ready(h, nSteps, where, message); if (e) return e;
    if (!flux_wm && !vol_wm) return fail(message, 1, "mr_upload_wm/neither fluxes nor target volumes given");
    const size_t N = (size_t)h->d.nRch, KB = (size_t)h->opt.max_batch;
    const Topology &T = h->topo;
    h->wmStage.resize((size_t)nSteps * N);
    CU(cudaStreamSynchronize(h->stream));                 // the rows of the previous batch may still be read
    for (int w = 0; w < 2; ++w) {
        const double *src = w == 0 ? flux_wm : vol_wm;
        double **dst = w == 0 ? &h->dWmFlux : &h->dWmVol;
        if (!src) continue;
        if (!*dst) { e = dev_alloc(h, dst, KB * N, where, message); if (e) return e; }
        for (int t = 0; t < nSteps; ++t)                  // caller's reach order -> stage order
            for (size_t p = 0; p < N; ++p) h->wmStage[(size_t)t * N + p] = src[(size_t)t * N + T.pos2rch[p]];
        CU(copy_sync(h, *dst, h->wmStage.data(), sizeof(double) * (size_t)nSteps * N, cudaMemcpyHostToDevice));
    }
    h->wmHasFlux = flux_wm != nullptr; h->wmHasVol = vol_wm != nullptr && h->opt.is_lake_sim;      // REACH_WM_VOL is 0 without is_lake_sim
    h->wmJumpStart = volJumpStart ? 1 : 0;
    h->wmSteps = nSteps;
    put_msg(message, "");
    return 0;
}

int mr_set_da(mr_handle h, int qmodOption, int qBlendPeriod, int QerrTrend, char *message) {
    if (!h) return fail(message, 1, "mr_set_da/null handle");
    if (qmodOption != 0 && qmodOption != 1) return fail(message, 1, "main_route/Error: qmodOption invalid");                  // main_route.f90:146-147
    if (qmodOption == 1 && (QerrTrend < 1 || QerrTrend > 4))                                                              // data_assimilation.f90:87
        return fail(message, 81, "direct_insertion/discharge error trend model must be 1(const),2(liear), or 3(logistic)");
    h->qmodOption = qmodOption; h->qBlendPeriod = qBlendPeriod; h->qErrTrend = QerrTrend;
    put_msg(message, "");
    return 0;
}

int mr_upload_obs(mr_handle h, int nSteps, const int *hasRecord, const double *obs, char *message) {
    const char *where = "mr_upload_obs";
    int e = check_ready(h, nSteps, where, message); if (e) return e;
    if (h->qmodOption != 1) return fail(message, 1, "mr_upload_obs/qmodOption is not 1 (mr_set_da)");
    if (!obs) return fail(message, 1, "mr_upload_obs/null observations");
    if ((e = ensure_da(h, where, message))) return e;
    const size_t N = (size_t)h->d.nRch;
    const Topology &T = h->topo;
    h->wmStage.resize((size_t)nSteps * N);
    for (int t = 0; t < nSteps; ++t)                      // caller's reach order -> stage order
        for (size_t p = 0; p < N; ++p) h->wmStage[(size_t)t * N + p] = obs[(size_t)t * N + T.pos2rch[p]];
    h->hasRecordHost.assign((size_t)nSteps, 1);
    if (hasRecord) for (int t = 0; t < nSteps; ++t) h->hasRecordHost[t] = hasRecord[t] ? 1 : 0;
    CU(cudaStreamSynchronize(h->stream));                 // the rows of the previous batch may still be read
    CU(copy_sync(h, h->dDaQobs, h->wmStage.data(), sizeof(double) * (size_t)nSteps * N, cudaMemcpyHostToDevice));
    CU(copy_sync(h, h->dHasRecord, h->hasRecordHost.data(), (size_t)nSteps, cudaMemcpyHostToDevice));
    h->obsSteps = nSteps;
    put_msg(message, "");
    return 0;
}

int mr_upload_lake_forcing(mr_handle h, int nSteps, const double *basinEvapo, const double *basinPrecip, char *message) {
    const char *where = "mr_upload_lake_forcing";
    int e = check_ready(h, nSteps, where, message); if (e) return e;
    if (!basinEvapo || !basinPrecip) return fail(message, 1, "mr_upload_lake_forcing/null evaporation or precipitation");
    if (!h->opt.is_lake_sim) return fail(message, 1, "mr_upload_lake_forcing/is_lake_sim is off");
    const size_t nH = (size_t)(h->d.nHRU > 0 ? h->d.nHRU : 1), KB = (size_t)h->opt.max_batch, nL = (size_t)(h->nLake > 0 ? h->nLake : 1);
    if (!h->dEvapo) {                                  // first use: rows of one batch, HRU level and lake-reach level
        e = dev_alloc(h, &h->dEvapo, KB * nH, where, message); if (e) return e;
        e = dev_alloc(h, &h->dPrecip, KB * nH, where, message); if (e) return e;
        e = dev_alloc(h, &h->dLakeEvap, KB * nL, where, message); if (e) return e;
        e = dev_alloc(h, &h->dLakePrecip, KB * nL, where, message); if (e) return e;
    }
    CU(cudaMemcpyAsync(h->dEvapo, basinEvapo, sizeof(double) * (size_t)nSteps * h->d.nHRU, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->dPrecip, basinPrecip, sizeof(double) * (size_t)nSteps * h->d.nHRU, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->lakeForcingSteps = nSteps;
    put_msg(message, "");
    return 0;
}

int mr_route_resident(mr_handle h, int nSteps, double T0, char *message) {
    const char *where = "mr_route_resident";
    int e = check_ready(h, nSteps, where, message); if (e) return e;
    CU(cudaEventRecord(h->ev[0], h->stream));
    e = route_device(h, nSteps, T0, where, message); if (e) return e;
    CU(cudaEventRecord(h->ev[5], h->stream));
    e = check_device_error(h, where, message); if (e) return e;
    collect_timing(h);
    float ms = 0.f; cudaEventElapsedTime(&ms, h->ev[0], h->ev[5]); h->timing[0] = ms;
    h->timing[3] = h->timing[4] = 0.0;
    put_msg(message, "");
    return 0;
}

// the same without waiting: kernels are enqueued on the handle's stream and the call returns; mr_wait collects errors
int mr_route_resident_async(mr_handle h, int nSteps, double T0, char *message) {
    const char *where = "mr_route_resident_async";
    int e = check_ready(h, nSteps, where, message); if (e) return e;
    CU(cudaEventRecord(h->ev[0], h->stream));
    e = route_device(h, nSteps, T0, where, message); if (e) return e;
    CU(cudaEventRecord(h->ev[5], h->stream));
    put_msg(message, "");
    return 0;
}

int mr_wait(mr_handle h, char *message) {
    const char *where = "mr_wait";
    if (!h || !h->hasNet) return fail(message, 1, "mr_wait/handle has no network");
    CU(cudaSetDevice(h->opt.device));
    int e = check_device_error(h, where, message); if (e) return e;
    if (h->copyOut) CU(cudaStreamSynchronize(h->copyOut));
    if (h->lastK > 0) {
        collect_timing(h);
        float ms = 0.f; cudaEventElapsedTime(&ms, h->ev[0], h->ev[5]); h->timing[0] = ms;
        h->timing[3] = h->timing[4] = 0.0;
    }
    put_msg(message, "");
    return 0;
}

static int download_q(mr_handle h, int nSteps, double *q_out, const char *where, char *message) {
    const int N = h->d.nRch;
    for (int r = 0; r < h->opt.n_routes; ++r) {
        const int m = h->opt.route_methods[r];
        dim3 grid((N + 255) / 256, nSteps < 64 ? nSteps : 64);
        k_unpermute_rows<<<grid, 256, 0, h->stream>>>(h->d.qSer[m], h->dOut + (size_t)r * nSteps * N, h->dRch2pos, N, nSteps);
        h->launchesLast++;
    }
    CU(cudaMemcpyAsync(q_out, h->dOut, sizeof(double) * (size_t)h->opt.n_routes * nSteps * N, cudaMemcpyDeviceToHost, h->stream));
    return 0;
}

int mr_download_q(mr_handle h, int nSteps, double *q_out, char *message) {
    const char *where = "mr_download_q";
    int e = check_ready(h, nSteps, where, message); if (e) return e;
    if (!q_out) return fail(message, 1, "mr_download_q/null output");
    if (nSteps > h->lastK) return fail(message, 1, "mr_download_q/more steps requested than the last batch routed");
    e = download_q(h, nSteps, q_out, where, message); if (e) return e;
    CU(cudaStreamSynchronize(h->stream));
    put_msg(message, "");
    return 0;
}

int mr_download_basin_q(mr_handle h, int nSteps, double *qr_out, char *message) {
    const char *where = "mr_download_basin_q";
    int e = check_ready(h, nSteps, where, message); if (e) return e;
    if (!qr_out) return fail(message, 1, "mr_download_basin_q/null output");
    if (nSteps > h->lastK) return fail(message, 1, "mr_download_basin_q/more steps requested than the last batch routed");
    const int N = h->d.nRch;
    dim3 grid((N + 255) / 256, nSteps < 64 ? nSteps : 64);
    k_unpermute_rows<<<grid, 256, 0, h->stream>>>(h->d.qrSer + N, h->dOut, h->dRch2pos, N, nSteps);      // rows 1..nSteps = BASIN_QR(1) after each step
    CU(cudaMemcpyAsync(qr_out, h->dOut, sizeof(double) * (size_t)nSteps * N, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    put_msg(message, "");
    return 0;
}

int mr_history_means(mr_handle h, int nSteps, int nAgg, int wantDlay, int flush, int maxPeriods, float *out, int *nPeriods, char *message) {
    const char *where = "mr_history_means";
    int e = check_ready(h, nSteps, where, message); if (e) return e;
    if (!out || !nPeriods || nAgg < 1 || maxPeriods < 0) return fail(message, 1, "mr_history_means/invalid arguments");
    if (nSteps > h->lastK) return fail(message, 1, "mr_history_means/more steps requested than the last batch routed");
    const int N = h->d.nRch, nSeries = h->opt.n_routes + (wantDlay ? 1 : 0);
    const int total = h->histCount + nSteps;
    const int nPer = total / nAgg + ((flush && total % nAgg) ? 1 : 0);
    if (nPer > maxPeriods) return fail(message, 1, "mr_history_means/more periods complete than the output buffer holds");
    if (!h->dHistAcc) { e = dev_alloc(h, &h->dHistAcc, (size_t)(h->opt.n_routes + 1) * N, where, message); if (e) return e; }
    const size_t need = (size_t)(nPer > 0 ? nPer : 1) * nSeries * N;
    if (need > h->histOutCap) { e = dev_alloc(h, &h->dHistOut, need, where, message, false); if (e) return e; h->histOutCap = need; }   // (the smaller one stays until mr_destroy)
    for (int s = 0; s < nSeries; ++s) {
        const double *rows = s < h->opt.n_routes ? h->d.qSer[h->opt.route_methods[s]] : h->d.qrSer + N;       // rows 1.. = BASIN_QR(1) after each step
        k_history<<<(N + 255) / 256, 256, 0, h->stream>>>(rows, h->dHistAcc + (size_t)s * N, h->dHistOut, h->dPos2rch, N, nSteps, nAgg, h->histCount, s, nSeries, flush ? 1 : 0);
        h->launchesLast++;
    }
    CU(cudaGetLastError());
    CU(copy_sync(h, out, h->dHistOut, sizeof(float) * (size_t)nPer * nSeries * N, cudaMemcpyDeviceToHost));
    h->histCount = flush ? 0 : total % nAgg;
    *nPeriods = nPer;
    put_msg(message, "");
    return 0;
}

int mr_step_batch(mr_handle h, int nSteps, double T0, const double *runoff, double *q_out, char *message) {
    const char *where = "mr_step_batch";
    int e = check_ready(h, nSteps, where, message); if (e) return e;
    if (!runoff) return fail(message, 1, "mr_step_batch/null runoff");
    CU(cudaEventRecord(h->ev[0], h->stream));
    CU(cudaMemcpyAsync(h->dRunoff, runoff, sizeof(double) * (size_t)nSteps * in_cols(h), cudaMemcpyHostToDevice, h->stream));
    stage_runoff(h, h->dRunoff, nSteps, h->stream);
    e = route_device(h, nSteps, T0, where, message); if (e) return e;
    CU(cudaEventRecord(h->ev[4], h->stream));
    if (q_out) { e = download_q(h, nSteps, q_out, where, message); if (e) return e; }
    CU(cudaEventRecord(h->ev[5], h->stream));
    e = check_device_error(h, where, message); if (e) return e;
    collect_timing(h);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev[0], h->ev[5]); h->timing[0] = ms;
    cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]); h->timing[3] = ms;
    cudaEventElapsedTime(&ms, h->ev[4], h->ev[5]); h->timing[4] = ms;
    put_msg(message, "");
    return 0;
}

// mr_step_batch without host waits, as a three-stage software pipeline over consecutive calls:
//   copy-in stream   H2D of this batch's forcing (double-buffered on the device)
//   handle's stream  routing of this batch, then the re-ordering of REACH_Q into the caller's reach order
//   copy-out stream  D2H of this batch's REACH_Q while the next batch is being routed
// runoff / q_out must stay valid (and should be pinned) until mr_wait returns.  Results are those of mr_step_batch.
int mr_step_batch_async(mr_handle h, int nSteps, double T0, const double *runoff, double *q_out, char *message) {
    const char *where = "mr_step_batch_async";
    int e = check_ready(h, nSteps, where, message); if (e) return e;
    if (!runoff) return fail(message, 1, "mr_step_batch_async/null runoff");
    if (!h->copyIn) {
        CU(cudaStreamCreateWithFlags(&h->copyIn, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&h->copyOut, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) { CU(cudaEventCreateWithFlags(&h->evIn[i], cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&h->evFree[i], cudaEventDisableTiming)); }
        CU(cudaEventCreateWithFlags(&h->evOut, cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&h->evD2H, cudaEventDisableTiming));
    }
    if (!h->dRunoffSlot[1]) {
        h->dRunoffSlot[0] = h->dRunoff;
        e = dev_alloc(h, &h->dRunoffSlot[1], (size_t)h->opt.max_batch * (in_cols(h) > 0 ? in_cols(h) : 1), where, message, false); if (e) return e;
    }
    const int slot = h->asyncSlot; h->asyncSlot ^= 1;
    if (h->freeRec[slot]) CU(cudaStreamWaitEvent(h->copyIn, h->evFree[slot], 0));          // the batch that used this slot has been routed
    CU(cudaMemcpyAsync(h->dRunoffSlot[slot], runoff, sizeof(double) * (size_t)nSteps * in_cols(h), cudaMemcpyHostToDevice, h->copyIn));
    CU(cudaEventRecord(h->evIn[slot], h->copyIn));
    CU(cudaStreamWaitEvent(h->stream, h->evIn[slot], 0));
    CU(cudaEventRecord(h->ev[0], h->stream));
    stage_runoff(h, h->dRunoffSlot[slot], nSteps, h->stream);
    e = route_device(h, nSteps, T0, where, message);
    if (e) return e;
    CU(cudaEventRecord(h->evFree[slot], h->stream)); h->freeRec[slot] = true;
    if (q_out) {
        const int N = h->d.nRch;
        if (h->d2hRec) CU(cudaStreamWaitEvent(h->stream, h->evD2H, 0));                    // the staging buffer has been copied out
        for (int r = 0; r < h->opt.n_routes; ++r) {
            const int m = h->opt.route_methods[r];
            dim3 grid((N + 255) / 256, nSteps < 64 ? nSteps : 64);
            k_unpermute_rows<<<grid, 256, 0, h->stream>>>(h->d.qSer[m], h->dOut + (size_t)r * nSteps * N, h->dRch2pos, N, nSteps);
            h->launchesLast++;
        }
        CU(cudaEventRecord(h->evOut, h->stream));
        CU(cudaStreamWaitEvent(h->copyOut, h->evOut, 0));
        CU(cudaMemcpyAsync(q_out, h->dOut, sizeof(double) * (size_t)h->opt.n_routes * nSteps * N, cudaMemcpyDeviceToHost, h->copyOut));
        CU(cudaEventRecord(h->evD2H, h->copyOut)); h->d2hRec = true;
    }
    CU(cudaEventRecord(h->ev[5], h->stream));
    put_msg(message, "");
    return 0;
}

int mr_step(mr_handle h, double T0, double T1, const double *basinRunoff, char *message) {
    if (h && h->hasNet && T1 - T0 != h->opt.dt) return fail(message, 1, "mr_step/T1-T0 differs from dt_qsim");
    int e = mr_step_batch(h, 1, T0, basinRunoff, nullptr, message);
    if (e && message) { std::string s(message); if (s.rfind("mr_step_batch", 0) == 0) put_msg(message, "mr_step" + s.substr(13)); }
    return e;
}

int mr_get_flux(mr_handle h, int method, int field, double *out, char *message) {
    const char *where = "mr_get_flux";
    if (!h || !h->hasNet || !out) return fail(message, 1, "mr_get_flux/handle has no network or null output");
    CU(cudaSetDevice(h->opt.device));
    const int N = h->d.nRch;
    const Topology &T = h->topo;
    const double *src = nullptr;
    std::vector<double> tmp(N);
    const bool perMethod = (field == MR_REACH_Q || field == MR_REACH_VOL1 || field == MR_REACH_VOL0 || field == MR_REACH_INFLOW || field == MR_WB || field == MR_QERROR);
    if (perMethod && (method < 0 || method >= N_METHODS || !h->on[method])) return fail(message, 1, "mr_get_flux/routing method is not active");
    switch (field) {
        case MR_REACH_Q: src = h->lastK > 0 ? h->d.qSer[method] + (size_t)(h->lastK - 1) * N : nullptr; break;
        case MR_REACH_VOL1: src = h->d.vol1[method]; break;
        case MR_REACH_VOL0: src = h->d.vol0[method]; break;
        case MR_REACH_INFLOW: src = h->d.inflow[method]; break;
        case MR_WB: src = h->d.wb[method]; break;
        case MR_QERROR: src = h->dQerr[method]; break;            // nullptr (zeros) before the first step with qmodOption 1
        case MR_BASIN_QI: src = h->d.basinQI; break;
        case MR_BASIN_QR1: src = h->d.qrSer + (size_t)h->lastK * N; break;
        case MR_BASIN_QR0: src = h->lastK > 0 ? h->d.qrSer + (size_t)(h->lastK - 1) * N : nullptr; break;
        case MR_R_WIDTH: src = h->d.rwidth; break;
        case MR_R_SLOPE: src = h->d.rslope; break;
        case MR_BASAREA: for (int r = 0; r < N; ++r) out[r] = T.basArea[T.rch2pos[r]]; put_msg(message, ""); return 0;
        case MR_TOTAREA: for (int r = 0; r < N; ++r) out[r] = T.totArea[T.rch2pos[r]]; put_msg(message, ""); return 0;
        case MR_NGOOD: for (int r = 0; r < N; ++r) out[r] = (double)T.nGood[T.rch2pos[r]]; put_msg(message, ""); return 0;
        default: return fail(message, 1, "mr_get_flux/unknown field");
    }
    if (!src) { for (int r = 0; r < N; ++r) out[r] = 0.0; put_msg(message, ""); return 0; }
    CU(cudaStreamSynchronize(h->stream));
    CU(copy_sync(h, tmp.data(), src, sizeof(double) * N, cudaMemcpyDeviceToHost));
    for (int r = 0; r < N; ++r) out[r] = tmp[T.rch2pos[r]];
    put_msg(message, "");
    return 0;
}

// ---- state in the restart schema -----------------------------------------------------------------
static long state_bytes(mr_handle h, int var) {
    const long N = h->d.nRch;
    switch (var) {
        case MR_ST_BASIN_QFUTURE: return 8L * N * h->ntdhBas;
        case MR_ST_BASIN_QR: return 8L * N * 2;
        case MR_ST_IRF_QFUTURE: return 8L * N * h->maxtdh;
        case MR_ST_IRF_VOL: return 8L * N;
        case MR_ST_KWT_NWAVE: return 4L * N;
        case MR_ST_KWT_QWAVE: case MR_ST_KWT_TENTRY: case MR_ST_KWT_TEXIT: return 8L * N * KWS;
        case MR_ST_KWT_ROUTED: return 4L * N * KWS;
        case MR_ST_LAKE_VOL: return 8L * N * h->opt.n_routes;
        case MR_ST_MOLECULE_KW: return 8L * N * n_molecule(M_KW);
        case MR_ST_MOLECULE_MC: return 8L * N * n_molecule(M_MC);
        case MR_ST_MOLECULE_DW: return 8L * N * n_molecule(M_DW);
        case MR_ST_QERROR: return 8L * N * h->opt.n_routes;
        case MR_ST_DA_QOBS: return 8L * N;
        case MR_ST_DA_QELAPSED: return 4L * N;
        default: return -1;
    }
}

int mr_get_state(mr_handle h, int var, void *buf, long nbytes, char *message) {
    const char *where = "mr_get_state";
    if (!h || !h->hasNet || !buf) return fail(message, 1, "mr_get_state/handle has no network or null buffer");
    if (state_bytes(h, var) != nbytes) return fail(message, 1, "mr_get_state/buffer size does not match the variable");
    CU(cudaSetDevice(h->opt.device));
    CU(cudaStreamSynchronize(h->stream));
    const int N = h->d.nRch;
    const Topology &T = h->topo;
    const long long tau = h->stepsDone;
    double *out = (double *)buf; int *iout = (int *)buf;
    auto pull = [&](const void *src, void *dst, size_t bytes) { return copy_sync(h, dst, src, bytes, cudaMemcpyDeviceToHost); };
    switch (var) {
        case MR_ST_BASIN_QFUTURE: {
            const int nb = h->ntdhBas;
            std::vector<double> tmp((size_t)nb * N);
            CU(pull(h->d.qfutBas, tmp.data(), tmp.size() * 8));
            for (int r = 0; r < N; ++r) { const int p = T.rch2pos[r];
                for (int k = 0; k < nb; ++k) out[(size_t)r * nb + k] = tmp[(size_t)((tau + k) % nb) * N + p]; }
            break; }
        case MR_ST_BASIN_QR: {
            std::vector<double> a(N, 0.0), b(N);
            if (h->lastK > 0) CU(pull(h->d.qrSer + (size_t)(h->lastK - 1) * N, a.data(), 8L * N));
            CU(pull(h->d.qrSer + (size_t)h->lastK * N, b.data(), 8L * N));
            for (int r = 0; r < N; ++r) { out[2 * r] = a[T.rch2pos[r]]; out[2 * r + 1] = b[T.rch2pos[r]]; }
            break; }
        case MR_ST_IRF_QFUTURE: {
            if (!h->on[M_IRF]) return fail(message, 1, "mr_get_state/IRF is not active");
            const int mx = h->maxtdh;
            std::vector<double> tmp((size_t)mx * N);
            CU(pull(h->d.qfutIrf, tmp.data(), tmp.size() * 8));
            for (int r = 0; r < N; ++r) { const int p = T.rch2pos[r], nt = h->ntdh[p];
                for (int k = 0; k < mx; ++k) out[(size_t)r * mx + k] = k < nt ? tmp[(size_t)((tau + k) % nt) * N + p] : 0.0; }
            break; }
        case MR_ST_IRF_VOL: {
            if (!h->on[M_IRF]) return fail(message, 1, "mr_get_state/IRF is not active");
            std::vector<double> tmp(N);
            CU(pull(h->d.vol1[M_IRF], tmp.data(), 8L * N));
            for (int r = 0; r < N; ++r) out[r] = tmp[T.rch2pos[r]];
            break; }
        case MR_ST_MOLECULE_KW: case MR_ST_MOLECULE_MC: case MR_ST_MOLECULE_DW: {
            const int m = var == MR_ST_MOLECULE_KW ? M_KW : var == MR_ST_MOLECULE_MC ? M_MC : M_DW, nm = n_molecule(m);
            if (!h->on[m]) return fail(message, 1, "mr_get_state/routing method is not active");
            std::vector<double> tmp((size_t)nm * N);
            CU(pull(h->d.mol[m], tmp.data(), tmp.size() * 8));
            for (int r = 0; r < N; ++r) { const int p = T.rch2pos[r]; for (int k = 0; k < nm; ++k) out[(size_t)r * nm + k] = tmp[(size_t)k * N + p]; }
            break; }
        case MR_ST_LAKE_VOL: {
            std::vector<double> tmp(N);
            for (int q = 0; q < h->opt.n_routes; ++q) {
                CU(pull(h->d.vol1[h->opt.route_methods[q]], tmp.data(), 8L * N));
                for (int r = 0; r < N; ++r) out[(size_t)q * N + r] = tmp[T.rch2pos[r]];
            }
            break; }
        case MR_ST_KWT_NWAVE: case MR_ST_KWT_QWAVE: case MR_ST_KWT_TENTRY: case MR_ST_KWT_TEXIT: case MR_ST_KWT_ROUTED: {
            if (!h->on[M_KWT]) return fail(message, 1, "mr_get_state/KWT is not active");
            const int b = (int)((tau + 1) & 1);          // buffer written by the last completed step
            std::vector<int> n(N), nr(N);
            CU(pull(h->d.kwN[b], n.data(), 4L * N)); CU(pull(h->d.kwNR[b], nr.data(), 4L * N));
            std::vector<double> tmp;
            if (var == MR_ST_KWT_QWAVE || var == MR_ST_KWT_TENTRY || var == MR_ST_KWT_TEXIT) {
                tmp.resize((size_t)KWP * N);
                const double *src = var == MR_ST_KWT_QWAVE ? h->d.kwQF[b] : (var == MR_ST_KWT_TENTRY ? h->d.kwTI[b] : h->d.kwTR[b]);
                CU(pull(src, tmp.data(), tmp.size() * 8));
            }
            // what the reference holds between steps is KWAVE(NR-1:) (kwt_route.f90:840-844,325-344)
            for (int r = 0; r < N; ++r) {
                const int p = T.rch2pos[r];
                const int first = nr[p] > 0 ? nr[p] - 1 : 0, cnt = n[p] - first;
                if (var == MR_ST_KWT_NWAVE) { iout[r] = cnt; continue; }
                for (int k = 0; k < KWS; ++k) {
                    if (var == MR_ST_KWT_ROUTED) iout[(size_t)r * KWS + k] = (k < cnt && first + k < nr[p]) ? 1 : 0;
                    else out[(size_t)r * KWS + k] = k < cnt ? tmp[(size_t)p * KWP + first + k] : -9999.0;
                }
            }
            break; }
        case MR_ST_QERROR: case MR_ST_DA_QOBS: case MR_ST_DA_QELAPSED: {
            if (!h->dDaQobs) { std::memset(buf, 0, (size_t)nbytes); break; }          // as initialised, init_model_data.f90:404-405
            if (var == MR_ST_DA_QELAPSED) {
                std::vector<int> tmp(N);
                CU(pull(h->dElState, tmp.data(), 4L * N));
                for (int r = 0; r < N; ++r) iout[r] = tmp[T.rch2pos[r]];
                break;
            }
            std::vector<double> tmp(N);
            const int nq = var == MR_ST_QERROR ? h->opt.n_routes : 1;
            for (int q = 0; q < nq; ++q) {
                const double *src = var == MR_ST_QERROR ? h->dQerr[h->opt.route_methods[q]] : h->dQobsState;
                if (!src) { for (int r = 0; r < N; ++r) out[(size_t)q * N + r] = 0.0; continue; }   // SUM, KWT: no direct_insertion
                CU(pull(src, tmp.data(), 8L * N));
                for (int r = 0; r < N; ++r) out[(size_t)q * N + r] = tmp[T.rch2pos[r]];
            }
            break; }
        default: return fail(message, 1, "mr_get_state/unknown state variable");
    }
    put_msg(message, "");
    return 0;
}

int mr_set_state(mr_handle h, int var, const void *buf, long nbytes, char *message) {
    const char *where = "mr_set_state";
    if (!h || !h->hasNet || !buf) return fail(message, 1, "mr_set_state/handle has no network or null buffer");
    if (state_bytes(h, var) != nbytes) return fail(message, 1, "mr_set_state/buffer size does not match the variable");
    CU(cudaSetDevice(h->opt.device));
    CU(cudaStreamSynchronize(h->stream));
    const int N = h->d.nRch;
    const Topology &T = h->topo;
    const long long tau = h->stepsDone;
    const double *in = (const double *)buf; const int *iin = (const int *)buf;
    auto push = [&](void *dst, const void *src, size_t bytes) { return copy_sync(h, dst, src, bytes, cudaMemcpyHostToDevice); };
    switch (var) {
        case MR_ST_BASIN_QFUTURE: {
            const int nb = h->ntdhBas;
            std::vector<double> tmp((size_t)nb * N);
            for (int r = 0; r < N; ++r) { const int p = T.rch2pos[r];
                for (int k = 0; k < nb; ++k) tmp[(size_t)((tau + k) % nb) * N + p] = in[(size_t)r * nb + k]; }
            CU(push(h->d.qfutBas, tmp.data(), tmp.size() * 8));
            break; }
        case MR_ST_BASIN_QR: {
            // BASIN_QR(1) becomes the carry-in row of the next batch; BASIN_QR(0) is overwritten by it at the next step
            std::vector<double> b(N);
            for (int r = 0; r < N; ++r) b[T.rch2pos[r]] = in[2 * r + 1];
            CU(push(h->d.qrSer, b.data(), 8L * N));
            h->lastK = 0;
            break; }
        case MR_ST_IRF_QFUTURE: {
            if (!h->on[M_IRF]) return fail(message, 1, "mr_set_state/IRF is not active");
            const int mx = h->maxtdh;
            std::vector<double> tmp((size_t)mx * N, 0.0);
            for (int r = 0; r < N; ++r) { const int p = T.rch2pos[r], nt = h->ntdh[p];
                for (int k = 0; k < nt; ++k) tmp[(size_t)((tau + k) % nt) * N + p] = in[(size_t)r * mx + k]; }
            CU(push(h->d.qfutIrf, tmp.data(), tmp.size() * 8));
            break; }
        case MR_ST_IRF_VOL: {
            if (!h->on[M_IRF]) return fail(message, 1, "mr_set_state/IRF is not active");
            std::vector<double> tmp(N);
            for (int r = 0; r < N; ++r) tmp[T.rch2pos[r]] = in[r];
            CU(push(h->d.vol1[M_IRF], tmp.data(), 8L * N));
            break; }
        case MR_ST_MOLECULE_KW: case MR_ST_MOLECULE_MC: case MR_ST_MOLECULE_DW: {
            const int m = var == MR_ST_MOLECULE_KW ? M_KW : var == MR_ST_MOLECULE_MC ? M_MC : M_DW, nm = n_molecule(m);
            if (!h->on[m]) return fail(message, 1, "mr_set_state/routing method is not active");
            std::vector<double> tmp((size_t)nm * N);
            for (int r = 0; r < N; ++r) { const int p = T.rch2pos[r]; for (int k = 0; k < nm; ++k) tmp[(size_t)k * N + p] = in[(size_t)r * nm + k]; }
            CU(push(h->d.mol[m], tmp.data(), tmp.size() * 8));
            break; }
        case MR_ST_LAKE_VOL: {
            std::vector<double> tmp(N);
            for (int q = 0; q < h->opt.n_routes; ++q) {
                for (int r = 0; r < N; ++r) tmp[T.rch2pos[r]] = in[(size_t)q * N + r];
                CU(push(h->d.vol1[h->opt.route_methods[q]], tmp.data(), 8L * N));
            }
            break; }
        case MR_ST_KWT_NWAVE: case MR_ST_KWT_ROUTED: {
            if (!h->on[M_KWT]) return fail(message, 1, "mr_set_state/KWT is not active");
            const int b = (int)((tau + 1) & 1);
            std::vector<int> tmp(N);
            if (var == MR_ST_KWT_NWAVE) {
                for (int r = 0; r < N; ++r) { if (iin[r] < 0 || iin[r] > KWS) return fail(message, 1, "mr_set_state/numWaves outside [0, MR_KW_SLOTS]"); tmp[T.rch2pos[r]] = iin[r]; }
                CU(push(h->d.kwN[b], tmp.data(), 4L * N));
            } else {                                      // routed flags must be a prefix (they always are, kwt_route.f90:303-311,1436)
                for (int r = 0; r < N; ++r) {
                    int nr = 0; bool gap = false;
                    for (int k = 0; k < KWS; ++k) { const int f = iin[(size_t)r * KWS + k] != 0; if (f && gap) return fail(message, 1, "mr_set_state/routed flags are not a prefix of the wave array"); if (f) ++nr; else gap = true; }
                    tmp[T.rch2pos[r]] = nr;
                }
                CU(push(h->d.kwNR[b], tmp.data(), 4L * N));
            }
            break; }
        case MR_ST_KWT_QWAVE: case MR_ST_KWT_TENTRY: case MR_ST_KWT_TEXIT: {
            if (!h->on[M_KWT]) return fail(message, 1, "mr_set_state/KWT is not active");
            const int b = (int)((tau + 1) & 1);
            std::vector<double> tmp((size_t)KWP * N, -9999.0);
            for (int r = 0; r < N; ++r) { const int p = T.rch2pos[r];
                for (int k = 0; k < KWS; ++k) tmp[(size_t)p * KWP + k] = in[(size_t)r * KWS + k]; }
            double *dst = var == MR_ST_KWT_QWAVE ? h->d.kwQF[b] : (var == MR_ST_KWT_TENTRY ? h->d.kwTI[b] : h->d.kwTR[b]);
            CU(push(dst, tmp.data(), tmp.size() * 8));
            break; }
        case MR_ST_QERROR: case MR_ST_DA_QOBS: case MR_ST_DA_QELAPSED: {
            int e = ensure_da(h, where, message); if (e) return e;
            CU(cudaStreamSynchronize(h->stream));
            if (var == MR_ST_DA_QELAPSED) {
                std::vector<int> tmp(N);
                for (int r = 0; r < N; ++r) tmp[T.rch2pos[r]] = iin[r];
                CU(push(h->dElState, tmp.data(), 4L * N));
                break;
            }
            std::vector<double> tmp(N);
            const int nq = var == MR_ST_QERROR ? h->opt.n_routes : 1;
            for (int q = 0; q < nq; ++q) {
                double *dst = var == MR_ST_QERROR ? h->dQerr[h->opt.route_methods[q]] : h->dQobsState;
                if (!dst) continue;
                for (int r = 0; r < N; ++r) tmp[T.rch2pos[r]] = in[(size_t)q * N + r];
                CU(push(dst, tmp.data(), 8L * N));
            }
            break; }
        default: return fail(message, 1, "mr_set_state/unknown state variable");
    }
    put_msg(message, "");
    return 0;
}

int mr_set_steps_done(mr_handle h, long steps, char *message) {
    if (!h || !h->hasNet) return fail(message, 1, "mr_set_steps_done/handle has no network");
    if (steps < 0) return fail(message, 1, "mr_set_steps_done/negative step count");
    // ring phases and the particle buffer parity are functions of the step count: re-seat the state
    const char *where = "mr_set_steps_done";
    if (steps == h->stepsDone) { put_msg(message, ""); return 0; }
    std::vector<std::vector<char>> keep;
    std::vector<int> vars = {MR_ST_BASIN_QFUTURE};
    if (h->on[M_IRF]) vars.push_back(MR_ST_IRF_QFUTURE);
    if (h->on[M_KWT]) { vars.push_back(MR_ST_KWT_NWAVE); vars.push_back(MR_ST_KWT_ROUTED); vars.push_back(MR_ST_KWT_QWAVE); vars.push_back(MR_ST_KWT_TENTRY); vars.push_back(MR_ST_KWT_TEXIT); }
    for (int v : vars) { keep.emplace_back(state_bytes(h, v)); int e = mr_get_state(h, v, keep.back().data(), (long)keep.back().size(), message); if (e) return e; }
    // carry row of BASIN_QR(1)
    if (h->lastK > 0) { CU(copy_sync(h, h->d.qrSer, h->d.qrSer + (size_t)h->lastK * h->d.nRch, 8L * h->d.nRch, cudaMemcpyDeviceToDevice)); h->lastK = 0; }
    h->stepsDone = steps;
    for (size_t i = 0; i < vars.size(); ++i) { int e = mr_set_state(h, vars[i], keep[i].data(), (long)keep[i].size(), message); if (e) return e; }
    put_msg(message, "");
    return 0;
}

int mr_get_basin_uh(mr_handle h, double *frac_future, char *message) {
    if (!h || !h->hasNet || !frac_future) return fail(message, 1, "mr_get_basin_uh/handle has no network or null output");
    std::memcpy(frac_future, h->fracFuture.data(), sizeof(double) * h->fracFuture.size());
    put_msg(message, "");
    return 0;
}

int mr_get_reach_uh(mr_handle h, int *ntdh, double *uh, char *message) {
    if (!h || !h->hasNet || !ntdh || !uh) return fail(message, 1, "mr_get_reach_uh/handle has no network or null output");
    if (!h->on[M_IRF]) return fail(message, 1, "mr_get_reach_uh/IRF is not active");
    const int N = h->d.nRch, mx = h->maxtdh;
    for (int r = 0; r < N; ++r) {
        const int p = h->topo.rch2pos[r];
        ntdh[r] = h->ntdh[p];
        for (int k = 0; k < mx; ++k) uh[(size_t)r * mx + k] = h->uhHost[(size_t)k * N + p];
    }
    put_msg(message, "");
    return 0;
}

long mr_get_info(mr_handle h, int key) {
    if (!h) return -1;
    switch (key) {
        case MR_INFO_NRCH: return h->d.nRch;
        case MR_INFO_NHRU: return h->d.nHRU;
        case MR_INFO_NFORCING: return (long)in_cols(h);
        case MR_INFO_NSTAGE: return h->topo.nStage;
        case MR_INFO_NTDH_BAS: return h->ntdhBas;
        case MR_INFO_MAXTDH: return h->maxtdh;
        case MR_INFO_LAUNCHES_LAST: return h->launchesLast;
        case MR_INFO_STEPS_DONE: return (long)h->stepsDone;
        case MR_INFO_MAX_BATCH: return h->opt.max_batch;
        case MR_INFO_MAX_NUPS: return h->topo.maxUps;
        case MR_INFO_DEVICE_BYTES: return (long)(h->devBytes >> 10);
        case MR_INFO_NHEAD: return h->topo.nHead;
        case MR_INFO_SUM_NTDH: { long s = 0; for (int v : h->ntdh) s += v; return s; }
        case MR_INFO_SUM_NUPS: return (long)h->topo.upIdx.size();
        case MR_INFO_KWT_TOUCHED: {
            if (!h->hasNet || !h->dKwCount) return 0;
            cudaSetDevice(h->opt.device);
            cudaStreamSynchronize(h->stream);
            std::vector<unsigned> c(h->d.nRch);
            copy_sync(h, c.data(), h->dKwCount, 4L * h->d.nRch, cudaMemcpyDeviceToHost);
            cudaMemsetAsync(h->dKwCount, 0, 4L * h->d.nRch, h->stream); cudaStreamSynchronize(h->stream);
            long tot = 0; for (unsigned v : c) tot += v;
            return tot;
        }
        case MR_INFO_KWT_PARTICLES: {
            if (!h->hasNet || !h->on[M_KWT]) return 0;
            cudaSetDevice(h->opt.device);
            cudaStreamSynchronize(h->stream);
            const int N = h->d.nRch, b = (int)((h->stepsDone + 1) & 1);
            std::vector<int> n(N), nr(N);
            copy_sync(h, n.data(), h->d.kwN[b], 4L * N, cudaMemcpyDeviceToHost);
            copy_sync(h, nr.data(), h->d.kwNR[b], 4L * N, cudaMemcpyDeviceToHost);
            long tot = 0;
            for (int p = 0; p < N; ++p) tot += n[p] - (nr[p] > 0 ? nr[p] - 1 : 0);
            return tot;
        }
        default: return -1;
    }
}

int mr_set_remap(mr_handle h, int nForcing, int nMap, const int *mapHruIndex, const int *numQhru, const int *qhruIndex, const double *weight, char *message) {
    const char *where = "mr_set_remap";
    if (!h || !h->hasNet) return fail(message, 1, "mr_set_remap/handle has no network");
    if (h->nForcing) return fail(message, 1, "mr_set_remap/already set for this network");
    if (nForcing < 1 || nMap < 1 || !mapHruIndex || !numQhru || !qhruIndex || !weight) return fail(message, 1, "mr_set_remap/missing argument");
    CU(cudaSetDevice(h->opt.device));
    CU(cudaStreamSynchronize(h->stream));
    std::vector<int> ptr(nMap + 1, 0), net(nMap);
    for (int i = 0; i < nMap; ++i) {
        if (numQhru[i] < 0) return fail(message, 1, "mr_set_remap/negative number of overlapping polygons");
        ptr[i + 1] = ptr[i] + numQhru[i];
        net[i] = (mapHruIndex[i] >= 0 && mapHruIndex[i] < h->d.nHRU) ? mapHruIndex[i] : -1;
    }
    const int nOv = ptr[nMap];
    std::vector<int> ov(qhruIndex, qhruIndex + nOv);
    for (int &q : ov) if (q < 0 || q >= nForcing) q = -1;
    std::vector<double> w(weight, weight + nOv);
    int e;
    e = dev_upload(h, &h->dMapNet, net, where, message); if (e) return e;
    e = dev_upload(h, &h->dMapPtr, ptr, where, message); if (e) return e;
    e = dev_upload(h, &h->dOvIdx, ov, where, message); if (e) return e;
    e = dev_upload(h, &h->dOvW, w, where, message); if (e) return e;
    const size_t KB = (size_t)h->opt.max_batch;
    e = dev_alloc(h, &h->dRunoffNet, KB * (h->d.nHRU > 0 ? h->d.nHRU : 1), where, message); if (e) return e;    // zero: HRUs outside the mapping
    // the input staging buffers now hold forcing polygons
    e = dev_alloc(h, &h->dRunoff, KB * nForcing, where, message, false); if (e) return e;
    h->dRunoffSlot[0] = h->dRunoffSlot[1] = nullptr;
    h->nForcing = nForcing; h->nMap = nMap;
    h->d.runoff = h->dRunoffNet;
    CU(cudaStreamSynchronize(h->stream));
    put_msg(message, "");
    return 0;
}

int mr_set_ghosts(mr_handle h, int nGhost, const int *ghostSegId, const int *kind, const double *totArea, const double *width, char *message) {
    if (!h) return fail(message, 1, "mr_set_ghosts/null handle");
    if (nGhost < 0 || (nGhost > 0 && (!ghostSegId || !kind || !totArea || !width))) return fail(message, 1, "mr_set_ghosts/missing argument");
    for (int g = 0; g < nGhost; ++g) if (kind[g] != 1 && kind[g] != 2) return fail(message, 1, "mr_set_ghosts/kind must be 1 or 2");
    h->ghostSegId.assign(ghostSegId, ghostSegId + nGhost); h->ghostKind.assign(kind, kind + nGhost);
    h->ghostTotArea.assign(totArea, totArea + nGhost); h->ghostWidth.assign(width, width + nGhost);
    put_msg(message, "");
    return 0;
}

int mr_set_export(mr_handle h, int nExport, const int *exportSegId, char *message) {
    const char *where = "mr_set_export";
    if (!h || !h->hasNet) return fail(message, 1, "mr_set_export/handle has no network");
    if (nExport < 0 || (nExport > 0 && !exportSegId)) return fail(message, 1, "mr_set_export/missing argument");
    if (h->nExport) return fail(message, 1, "mr_set_export/already set for this network");
    CU(cudaSetDevice(h->opt.device));
    const Topology &T = h->topo;
    const int N = h->d.nRch;
    // id -> caller index through the handle's own reach ids is not kept; resolve against positions via downIndex order
    std::vector<int> slot(N, -1), pos(nExport);
    std::vector<int> idx;
    join_ids(exportSegId, nExport, h->segIdCopy.data(), N, idx);
    for (int k = 0; k < nExport; ++k) {
        if (idx[k] < 0) return fail(message, 1, "mr_set_export/reach id is not in the network");
        pos[k] = T.rch2pos[idx[k]]; slot[pos[k]] = k;
    }
    int e = dev_upload(h, &h->dExpSlot, slot, where, message); if (e) return e;
    e = dev_upload(h, &h->dExpPos, pos, where, message); if (e) return e;
    h->d.expSlot = h->dExpSlot;
    h->nExport = nExport;
    put_msg(message, "");
    return 0;
}

long mr_exchange_bytes(mr_handle h, int which) {
    if (!h || !h->hasNet || which < 0 || which > 1) return -1;
    const long n = which == 0 ? h->nExport : h->nGhost;
    return 8L * n * h->d.kmax * h->d.recLen;
}

int mr_set_exchange_buffer(mr_handle h, int which, void *dev, long nbytes, char *message) {
    const char *where = "mr_set_exchange_buffer";
    if (!h || !h->hasNet || which < 0 || which > 1) return fail(message, 1, "mr_set_exchange_buffer/bad handle or selector");
    CU(cudaSetDevice(h->opt.device));
    CU(cudaStreamSynchronize(h->stream));
    const long need = mr_exchange_bytes(h, which);
    double *ptr = (double *)dev;
    if (!ptr) {
        int e = dev_alloc(h, &ptr, (size_t)(need / 8), where, message); if (e) return e;
        CU(cudaStreamSynchronize(h->stream));
        h->xowned[which] = true;
    } else {
        if (nbytes < need) return fail(message, 1, "mr_set_exchange_buffer/buffer smaller than mr_exchange_bytes");
        h->xowned[which] = false;
    }
    h->xbuf[which] = ptr;
    if (which == 0) h->d.expBuf = ptr; else h->d.impBuf = ptr;
    put_msg(message, "");
    return 0;
}

int mr_get_exchange_buffer(mr_handle h, int which, void **dev, long *nbytes, char *message) {
    if (!h || !h->hasNet || which < 0 || which > 1 || !dev || !nbytes) return fail(message, 1, "mr_get_exchange_buffer/bad argument");
    *dev = h->xbuf[which]; *nbytes = mr_exchange_bytes(h, which);
    put_msg(message, "");
    return 0;
}

int mr_copy_exchange(mr_handle src, mr_handle dst, int srcSlot0, int dstSlot0, int nSlots, char *message) {
    const char *where = "mr_copy_exchange";
    if (!src || !dst || !src->hasNet || !dst->hasNet) return fail(message, 1, "mr_copy_exchange/handle has no network");
    if (src->d.recLen != dst->d.recLen || src->d.kmax != dst->d.kmax) return fail(message, 1, "mr_copy_exchange/route_opt or max_batch differ between the domains");
    if (nSlots < 0 || srcSlot0 < 0 || dstSlot0 < 0 || srcSlot0 + nSlots > src->nExport || dstSlot0 + nSlots > dst->nGhost)
        return fail(message, 1, "mr_copy_exchange/slot range outside the buffers");
    if (!src->xbuf[0] || !dst->xbuf[1]) return fail(message, 1, "mr_copy_exchange/exchange buffers are not set");
    const size_t rec = (size_t)src->d.kmax * src->d.recLen;
    CU(cudaStreamSynchronize(src->stream));
    CU(cudaMemcpyAsync(dst->xbuf[1] + (size_t)dstSlot0 * rec, src->xbuf[0] + (size_t)srcSlot0 * rec, 8 * rec * nSlots, cudaMemcpyDeviceToDevice, dst->stream));
    CU(cudaStreamSynchronize(dst->stream));
    put_msg(message, "");
    return 0;
}

int mr_set_stream(mr_handle h, void *cuda_stream, char *message) {
    const char *where = "mr_set_stream";
    if (!h) return fail(message, 1, "mr_set_stream/null handle");
    CU(cudaSetDevice(h->opt.device));
    CU(cudaStreamSynchronize(h->stream));
    h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->ownStream;
    put_msg(message, "");
    return 0;
}

int mr_set_counting(mr_handle h, int enabled, char *message) {
    const char *where = "mr_set_counting";
    if (!h || !h->hasNet) return fail(message, 1, "mr_set_counting/handle has no network");
    CU(cudaSetDevice(h->opt.device));
    CU(cudaStreamSynchronize(h->stream));
    if (enabled && !h->dKwCount) { int e = dev_alloc(h, &h->dKwCount, (size_t)h->d.nRch, where, message); if (e) return e; }
    h->d.kwCount = enabled ? h->dKwCount : nullptr;
    CU(cudaStreamSynchronize(h->stream));
    put_msg(message, "");
    return 0;
}

int mr_selftest_pow04(mr_handle h, int n, const double *x, double *y, char *message) {
    const char *where = "mr_selftest_pow04";
    if (!h || n < 1 || !x || !y) return fail(message, 1, "mr_selftest_pow04/invalid arguments");
    CU(cudaSetDevice(h->opt.device));
    double *dx = nullptr, *dy = nullptr;
    CU(cudaMalloc((void **)&dx, sizeof(double) * n)); CU(cudaMalloc((void **)&dy, sizeof(double) * n));
    CU(copy_sync(h, dx, x, sizeof(double) * n, cudaMemcpyHostToDevice));
    k_selftest_pow04<<<(n + 255) / 256, 256, 0, h->stream>>>(n, dx, dy);
    CU(copy_sync(h, y, dy, sizeof(double) * n, cudaMemcpyDeviceToHost));
    cudaFree(dx); cudaFree(dy);
    put_msg(message, "");
    return 0;
}

int mr_get_timing(mr_handle h, double *ms) {
    if (!h || !ms) return 1;
    for (int i = 0; i < 8; ++i) ms[i] = h->timing[i];
    return 0;
}

void mr_destroy(mr_handle h) {
    if (!h) return;
    cudaSetDevice(h->opt.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->hasNet && h->d.kwProf) {                  // MR_KWT_PROFILE=1: cycles per task class (development)
        unsigned long long pr[40];
        if (copy_sync(h, pr, h->d.kwProf, sizeof(pr), cudaMemcpyDeviceToHost) == cudaSuccess) {
            static const char *ph[8] = {"per-task: gather + merge decisions", "pool build", "pool: merged flows", "per-task: thinning", "pool build", "pool: kinwav", "per-task: time average", "pool: stores"};
            for (int k = 0; k < 2; ++k) {
                double tp = 0; for (int c = 0; c < 8; ++c) tp += (double)pr[16 + 8 * k + c];
                for (int c = 0; c < 8; ++c) if (tp > 0) std::fprintf(stderr, "kwt %s phase %-36s %5.1f%% of the block cycles\n", k ? "heavy" : "light", ph[c], 100.0 * pr[16 + 8 * k + c] / tp);
                if (pr[32 + k]) std::fprintf(stderr, "kwt %s blocks %llu, cycles per block %.0f\n", k ? "heavy" : "light", pr[32 + k], tp / pr[32 + k]);
            }
            static const char *nm[8] = {"n=0", "n<=3", "n<=6", "n<=12", "n<=20", "n<=40", "n>40", "retry"};
            double tot = 0; for (int c = 0; c < 8; ++c) tot += (double)pr[2 * c];
            for (int c = 0; c < 8; ++c) if (pr[2 * c + 1])
                std::fprintf(stderr, "kwt tasks %-6s count %12llu  share of cycles %5.1f%%  cycles/task %9.0f\n", nm[c], pr[2 * c + 1], 100.0 * pr[2 * c] / tot, (double)pr[2 * c] / pr[2 * c + 1]);
        }
    }
    free_device(h);
    for (auto &e : h->ev) if (e) cudaEventDestroy(e);
    for (auto &m : h->mev) for (auto &e : m) if (e) cudaEventDestroy(e);
    for (auto &a : h->aux) if (a) { cudaStreamSynchronize(a); cudaStreamDestroy(a); }
    if (h->kwSide) { cudaStreamSynchronize(h->kwSide); cudaStreamDestroy(h->kwSide); }
    for (auto &e : h->kwEv) if (e) cudaEventDestroy(e);
    for (cudaStream_t c : {h->copyIn, h->copyOut}) if (c) { cudaStreamSynchronize(c); cudaStreamDestroy(c); }
    for (cudaEvent_t ev : {h->evIn[0], h->evIn[1], h->evFree[0], h->evFree[1], h->evOut, h->evD2H}) if (ev) cudaEventDestroy(ev);
    if (h->ownStream) cudaStreamDestroy(h->ownStream);
    delete h;
}

}  // extern "C"
