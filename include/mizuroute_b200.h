/*
 * mizuroute_b200.h -- C ABI of the B200-native reach-routing solver.
 *
 * The reference (ESCOMP/mizuRoute, Fortran) has no FFI/plug-in layer.  The seam this library sits
 * behind is the Fortran module interface of the per-step routing driver
 *
 *     main_route(basinRunoff_in, ..., NETOPO_in, RPARAM_in, RCHFLX_out, RCHSTA_out, ..., ierr, message)
 *         route/build/src/main_route.f90:29-43   (called by mpi_route, mpi_process.f90:1217,1294)
 *
 * and, below it, the abstract per-reach operator  base_route_rch%route  (base_route.f90:27-63)
 * with its IRF / KWT / SUM implementations (irf_route.f90:40, kwt_route.f90:36, accum_runoff.f90:32).
 * Each entry point below names the reference interface it replaces.  INTEGRATION.md shows the
 * bind(C) shim a mizuRoute maintainer would add to call it from mpi_route.
 *
 * Conventions
 *   - plain C: pointers and sizes only, all arrays caller-owned HOST memory unless named *_dev*;
 *   - every function returns ierr (0 = ok; reference codes 10/20/30/40/60 are kept where the
 *     reference raises them) and writes a NUL-terminated, path-like message ("mr_step/kwt_rch/...")
 *     into the caller's 256-byte buffer, mirroring `ierr, message` (main_route.f90:42);
 *   - reach-dimensioned arrays are in the caller's reach order (the order of segId passed to
 *     mr_set_network); HRU-dimensioned arrays in the caller's HRU order (NETOPO%HRUIX indexing,
 *     main_route.f90:50);
 *   - all reals are double (real(dp)), all integers 32-bit (integer(i4b));
 *   - a handle is bound to one CUDA device and is not re-entrant (like the module globals it replaces).
 */
#ifndef MIZUROUTE_B200_H
#define MIZUROUTE_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define MR_STRLEN 256            /* strLen, public_var.f90:25 */
#define MR_MAXQPAR 20            /* MAXQPAR, public_var.f90:37 */
#define MR_KW_SLOTS 22           /* KWAVE(0:NQ2+1) can momentarily hold 21 entries (kwt_route.f90:299) */

/* routing-method ids, public_var.f90:74-80 */
enum { MR_ACCUM_RUNOFF = 0, MR_IMPULSE_RESPONSE_FUNC = 1, MR_KINEMATIC_WAVE_TRACKING = 2,
       MR_KINEMATIC_WAVE = 3, MR_MUSKINGUM_CUNGE = 4, MR_DIFFUSIVE_WAVE = 5 };

/* flux fields of STRFLX (dataTypes.f90:346-377) readable with mr_get_flux */
enum {
    MR_REACH_Q = 0, MR_REACH_VOL1 = 1, MR_REACH_INFLOW = 2, MR_WB = 3,
    MR_BASIN_QI = 4, MR_BASIN_QR1 = 5, MR_BASIN_QR0 = 6, MR_REACH_VOL0 = 7,
    MR_QERROR = 8,               /* ROUTE(:)%Qerror of IRF / KW / MC / DW under data assimilation (dataTypes.f90:354) */
    /* derived reach parameters (RCHPRP, dataTypes.f90:183-255) */
    MR_R_WIDTH = 10, MR_TOTAREA = 11, MR_BASAREA = 12, MR_R_SLOPE = 13,
    MR_NGOOD = 14                /* count(goodBas), network_topo.f90:769-775 (as a double) */
};

/* lake model types, public_var.f90 / lake_route.f90:196-438 */
enum { MR_LAKE_ENDORHEIC = 0, MR_LAKE_DOLL03 = 1, MR_LAKE_HANASAKI06 = 2, MR_LAKE_HYPE = 3 };

/* state variables in the reference's restart schema (write_restart_pio.f90:544,1039-1134; read_restart.f90:402-470) */
enum {
    MR_ST_BASIN_QFUTURE = 0,   /* double [nRch][ntdh_bas]      "qfuture"      */
    MR_ST_BASIN_QR      = 1,   /* double [nRch][2]             BASIN_QR(0:1)  */
    MR_ST_IRF_QFUTURE   = 2,   /* double [nRch][maxtdh]        "irf_qfuture" (rows padded with 0 beyond ntdh(reach)) */
    MR_ST_IRF_VOL       = 3,   /* double [nRch]                "volume_irf"   */
    MR_ST_KWT_NWAVE     = 4,   /* int    [nRch]                "numWaves"     */
    MR_ST_KWT_QWAVE     = 5,   /* double [nRch][MR_KW_SLOTS]   "qwave"        */
    MR_ST_KWT_TENTRY    = 6,   /* double [nRch][MR_KW_SLOTS]   "tentry"       */
    MR_ST_KWT_TEXIT     = 7,   /* double [nRch][MR_KW_SLOTS]   "texit"        */
    MR_ST_KWT_ROUTED    = 8,   /* int    [nRch][MR_KW_SLOTS]   "routed"       */
    MR_ST_LAKE_VOL      = 9,   /* double [nRoutes][nRch]       REACH_VOL(1) of every active method ("volume_irf|kwt|kw|mc|dw") */
    MR_ST_MOLECULE_KW   = 10,  /* double [nRch][20]            "q_sub_kw": molecule%Q of the kinematic wave (init_model_data.f90:388) */
    MR_ST_MOLECULE_MC   = 11,  /* double [nRch][2]             "q_sub_mc": inflow / outflow of the previous step */
    MR_ST_MOLECULE_DW   = 12,  /* double [nRch][20]            "q_sub_dw" */
    MR_ST_QERROR        = 13,  /* double [nRoutes][nRch]       Qerror of every active method (restart variables of ixIRF%qerror .. ixDW%qerror,
                                  read_restart.f90:381,538,611,684; zero for SUM and KWT, which are not corrected) */
    MR_ST_DA_QOBS       = 14,  /* double [nRch]                RCHFLX%Qobs, the last gauge value seen (not in the reference's restart file, which
                                  starts again from 0: init_model_data.f90:511-512) */
    MR_ST_DA_QELAPSED   = 15   /* int    [nRch]                RCHFLX%Qelapsed, steps since then (likewise) */
};

/* integer facts about a handle, mr_get_info */
enum {
    MR_INFO_NRCH = 0, MR_INFO_NHRU = 1, MR_INFO_NSTAGE = 2, MR_INFO_NTDH_BAS = 3, MR_INFO_MAXTDH = 4,
    MR_INFO_LAUNCHES_LAST = 5,   /* kernels launched by the last mr_step / mr_step_batch / mr_route_resident */
    MR_INFO_STEPS_DONE = 6,      /* iTime-1 (globalData iTime) */
    MR_INFO_MAX_BATCH = 7, MR_INFO_MAX_NUPS = 8, MR_INFO_KWT_PARTICLES = 9, /* live particles in the KWT state */
    MR_INFO_DEVICE_BYTES = 10,   /* bytes of HBM held by the handle (KiB) */
    MR_INFO_KWT_TOUCHED = 11, MR_INFO_NHEAD = 12, MR_INFO_SUM_NTDH = 13, MR_INFO_SUM_NUPS = 14,
    MR_INFO_NFORCING = 15        /* columns of the runoff arrays the step calls expect (nHRU, or the forcing polygons of mr_set_remap) */
};

typedef struct mr_handle_s *mr_handle;

/* Control-file keys the routing path reads (read_control.f90; defaults public_var.f90:100-145) and the
 * spatially-constant parameters of namelist param_nml (read_param.f90:26-38). */
typedef struct {
    double dt;                   /* <dt_qsim> [s] */
    int    n_routes;             /* number of digits in <route_opt> */
    int    route_methods[8];     /* the digits, in order (read_control.f90:583-597) */
    int    doesBasinRoute;       /* 1: hillslope UH (default) */
    int    hw_drain_point;       /* 1 top / 2 bottom (default) of headwater reach, irf_route.f90:91-111 */
    double min_length_route;     /* irf_route.f90:237 */
    int    is_lake_sim, lakeRegulate, LakeInputOption;
    double runoffMin;            /* public_var.f90:145 */
    double time_conv, length_conv;   /* from <units_qsim>, read_control.f90:443-474 */
    double fshape, tscale;       /* &HSLOPE */
    double velo, diff;           /* &IRF_UH */
    double mann_n, wscale;       /* &KWT    */
    int    device;               /* CUDA device ordinal */
    int    max_batch;            /* largest nSteps a batch call may pass (sizes the resident series) */
    /* Euler schemes (route methods 3/4/5) only; zero = the reference's defaults */
    int    floodplain;           /* <floodplain>: 1 = finite bankfull depth dscale*sqrt(totalArea), 0 = high_depth (process_ntopo.f90:174-196) */
    double dscale;               /* bankfull depth scaling, 0 -> real(0.000045) = 4.5000000682193786e-05, a default-real literal (globalData.f90:187) */
    double floodplainSlope;      /* floodplain slope h:v, 0 -> 1000 (globalData.f90:188) */
} mr_options;

/* Replaces init_route_method + the option globals (init_model_data.f90:753-805, public_var.f90). */
int mr_create(const mr_options *opts, mr_handle *out, char *message);

/* Replaces init_ntopo/augment_ntopo/put_data_struct (process_ntopo.f90:39-513): takes the variables
 * read_streamSeg.f90:getData (:44) reads from the river-network file and derives topology, areas,
 * widths, unit hydrographs, and the device layout.  Optional arrays may be NULL.  Cold-start state
 * (init_model_data.f90:399-463). */
int mr_set_network(mr_handle h, int nRch, int nHRU,
                   const int *segId, const int *downSegId,
                   const int *hruSegId, const double *hruArea,
                   const double *length, const double *slope,
                   const double *width, const double *man_n,
                   const int *islake, const int *lakeModelType,
                   const double *D03_MaxStorage, const double *D03_Coefficient,
                   const double *D03_Power, const double *D03_S0,
                   char *message);

/* Optional: runoff arrives on other polygons than the river-network HRUs (<is_remap> T).  Replaces remap_1D_runoff
 * (process_remap.f90:164-262, called from get_hru_runoff, get_basin_runoff.f90:91-93) ON THE DEVICE: after this call
 * every runoff array passed to mr_step / mr_step_batch* / mr_upload_runoff is [nSteps][nForcing].
 *   mapHruIndex[nMap]  network-HRU index (0-based, caller's HRU order) of each mapping-layer HRU, -1 = not in the network
 *   numQhru[nMap]      overlapping forcing polygons per mapping HRU (ragged rows of the next two arrays, file order)
 *   qhruIndex[sum]     index of the polygon in the forcing vector, -1 = polygon without forcing
 *   weight[sum]        areal weights */
int mr_set_remap(mr_handle h, int nForcing, int nMap, const int *mapHruIndex, const int *numQhru, const int *qhruIndex, const double *weight, char *message);

/* Replaces one main_route call (main_route.f90:29-268) for the time step [T0,T1] = TSEC(1:2). */
int mr_step(mr_handle h, double T0, double T1, const double *basinRunoff /* [nHRU] */, char *message);

/* nSteps consecutive main_route calls starting at T0 (T advances as update_time does,
 * init_model_data.f90:311-312), executed as one time-skewed wavefront.  runoff is [nSteps][nHRU];
 * q_out, if not NULL, receives REACH_Q as [n_routes][nSteps][nRch]. */
int mr_step_batch(mr_handle h, int nSteps, double T0, const double *runoff, double *q_out, char *message);

/* The same in three stages, for callers that keep forcing resident in HBM:
 * upload -> route (device only, no host<->device traffic) -> download. */
int mr_upload_runoff(mr_handle h, int nSteps, const double *runoff, char *message);
int mr_route_resident(mr_handle h, int nSteps, double T0, char *message);
/* Forcing ingest on the device, for callers that hold raw forcing records instead of per-step rows (replaces what
 * get_basin_runoff does to the runoff, get_basin_runoff.f90:19-251, when no areal remapping is needed):
 *   mr_set_ingest      forcingOfHru [nHRU]: column of every river-network HRU in the forcing records, -1 = none (the IX_in of
 *                      sort_flux inverted, process_remap.f90:271-314); scale / offset = <scale_factor_runoff> / <offset_value_runoff>
 *                      (-9999 = not given, scale_forcing :375-423); fill = the records' fill value.  After mr_set_network.
 *   mr_ingest_records  records [nRec][nForcing]; step t takes records recIdx[recPtr[t] .. recPtr[t+1]) with the shares recFrac of
 *                      the step (timeMap_sim_forc, :256-369; recFrac NULL = exactly one record per step, taken as it is); the
 *                      time-weighted mean skips fill values (read_1D_forcing, read_runoff.f90:298-325); HRUs without forcing and
 *                      negative values become 0.  Leaves nSteps runoff rows resident, as mr_upload_runoff does: follow with
 *                      mr_route_resident(h, nSteps, T0). */
int mr_set_ingest(mr_handle h, int nForcing, const int *forcingOfHru, double scale, double offset, double fill, char *message);
int mr_ingest_records(mr_handle h, int nSteps, int nRec, const double *records, const int *recPtr, const int *recIdx, const double *recFrac, char *message);
/* Lake evaporation / precipitation of the NEXT routing call (mr_step, mr_step_batch*, mr_route_resident*), which must
 * route exactly nSteps steps: basinEvapo / basinPrecip [nSteps][nHRU] in the units of the runoff, river-network HRU
 * order (the optional arguments basinEvapo_in / basinPrecip_in of main_route, main_route.f90:33-34,174-199).  They pass
 * through the same basin2reach as the runoff and enter lake_route when LakeInputOption is 0 or 2 (lake_route.f90:166-174)
 * and the lake water balance.  A routing call without a preceding upload uses exactly zero for both. */
int mr_upload_lake_forcing(mr_handle h, int nSteps, const double *basinEvapo, const double *basinPrecip, char *message);
/* Water management of the NEXT routing call (which must route exactly nSteps steps): flux_wm [nSteps][nRch] = REACH_WM_FLUX,
 * abstraction (+) / injection (-) in m3/s, -9999 = none at that reach (<is_flux_wm>, the reachflux_in argument of main_route,
 * main_route.f90:110-116); vol_wm [nSteps][nRch] = REACH_WM_VOL, the target volume of the lakes flagged with the lake parameter
 * "LakeTargVol" (<is_vol_wm>, :117-123); volJumpStart = <is_vol_wm_jumpstart>.  Either array may be NULL (= that option off).
 * Caller's reach order.  On the device: the abstraction cascade of irf_rch and of the Euler schemes (irf_route.f90:114-142),
 * extract_from_rch of KWT (kwt_route.f90:351-455; note that it reads the sign the other way round) and the lake fluxes /
 * target volumes (lake_route.f90:137-139,176-203). */
int mr_upload_wm(mr_handle h, int nSteps, const double *flux_wm, const double *vol_wm, int volJumpStart, char *message);
/* Data assimilation by direct insertion (<qmodOption> 1, <qBlendPeriod>, <QerrTrend> 1 constant / 2 linear / 3 logistic /
 * 4 exponential; public_var.f90:189-191): where a gauge value is known, REACH_Q of IRF, KW, MC and DW is pulled to it, the
 * correction fading over qBlendPeriod steps (data_assimilation.f90:23-97; irf_route.f90:188-199 and the Euler schemes); the
 * water balance of those methods is then not evaluated (:200-202).  SUM and KWT are not corrected.  qmodOption 0 = off.
 * Errors as the reference: ierr 1 "qmodOption invalid" (main_route.f90:146-147), ierr 81 for an unknown trend model. */
int mr_set_da(mr_handle h, int qmodOption, int qBlendPeriod, int QerrTrend, char *message);
/* Gauge observations of the NEXT routing call (which must route exactly nSteps steps): obs [nSteps][nRch] in m3/s, caller's
 * reach order, NaN or negative = no value at that reach (main_route.f90:136-142: gage_obs_data%read_obs + link_ix);
 * hasRecord [nSteps], 0 = the gauge file has no record at that step (time_ix = integerMissing: every reach's Qelapsed goes up
 * by one, :143-145), NULL = every step has one.  A routing call without a preceding upload is a stretch without records. */
int mr_upload_obs(mr_handle h, int nSteps, const int *hasRecord, const double *obs, char *message);
/* Per-reach parameters of the parametric lake models beyond Doll-2003, by their name in RCHPRP (dataTypes.f90:202-213):
 * HYP_E_emr, HYP_E_lim, HYP_E_min, HYP_E_zero, HYP_Qrate_emr, HYP_Erate_emr, HYP_Qrate_prim, HYP_Qrate_amp, HYP_Qrate_phs,
 * HYP_prim_F, HYP_A_avg, HYP_Qsim_mode, and (dataTypes.f90:215-254) H06_Smax, H06_alpha, H06_envfact, H06_S_ini, H06_c1, H06_c2,
 * H06_exponent, H06_denominator, H06_c_compare, H06_frac_Sdead, H06_E_rel_ini, H06_I_Jan..H06_I_Dec, H06_D_Jan..H06_D_Dec,
 * H06_purpose, H06_I_mem_F, H06_D_mem_F, H06_I_mem_L, H06_D_mem_L (integers and logicals as doubles); values[n = nRch] in the
 * caller's reach order; "LakeTargVol" (NETOPO%LakeTargVol, 0/1) flags the lakes that follow the target volume of mr_upload_wm.
 * Call BEFORE mr_set_network (like mr_set_ghosts).  lakeModelType 3 (HYPE) needs all HYP_*, lakeModelType
 * 2 (Hanasaki 2006) all H06_* (the demand memory H06_D_mem_* belongs to water management and is not used). */
int mr_set_lake_param(mr_handle h, const char *name, int n, const double *values, char *message);
/* Datetime of the first simulation step (simDatetime(1) at iTime = 1, init_model_data.f90) and the calendar (0 standard /
 * gregorian / proleptic_gregorian, 1 noleap): HYPE's seasonal spillway reads the day of year (lake_route.f90:404). */
int mr_set_sim_start(mr_handle h, int year, int month, int day, double secOfDay, int noleap, char *message);
int mr_download_q(mr_handle h, int nSteps, double *q_out, char *message);
/* BASIN_QR(1) ("dlayRunoff", the hillslope-routed lateral inflow) of the last batch as [nSteps][nRch] */
int mr_download_basin_q(mr_handle h, int nSteps, double *qr_out, char *message);
/* History aggregation on the device (histVars_data.f90:154-246, aggregate / finalize): the period means of REACH_Q of every
 * routing method (route_opt order) and, with wantDlay, of BASIN_QR(1) over consecutive groups of nAgg steps, formed from the
 * first nSteps steps of the LAST batch in step order and rounded to float32, the history file's type.  A period may span
 * calls: the sums and the number of steps in them stay on the device.  flush != 0 also closes the period still open after
 * the last step (the end of the run).  out receives [nPeriods][n_routes (+1)][nRch] in the caller's reach order, at most
 * maxPeriods periods (ierr 1 if more complete); *nPeriods = periods written.  The device -> host copy is nPeriods records
 * instead of nSteps: with daily means of hourly steps 1/48 of the bytes of mr_download_q. */
int mr_history_means(mr_handle h, int nSteps, int nAgg, int wantDlay, int flush, int maxPeriods, float *out, int *nPeriods, char *message);
/* mr_route_resident without the final wait: the kernels are enqueued on the handle's stream and the call returns,
 * so a caller can keep several domains (tributaries of the next batch, mainstem of this one) in flight on
 * different streams.  mr_wait blocks until the stream is idle and reports a device-side error (ierr, message). */
int mr_route_resident_async(mr_handle h, int nSteps, double T0, char *message);
/* mr_step_batch as a software pipeline over consecutive calls: the forcing of this batch is uploaded on a copy stream
 * (double-buffered), routed on the handle's stream, and its REACH_Q is downloaded on a second copy stream while the
 * next batch is routed.  Host buffers must stay valid (pinned for real overlap) until mr_wait. */
int mr_step_batch_async(mr_handle h, int nSteps, double T0, const double *runoff, double *q_out, char *message);
int mr_wait(mr_handle h, char *message);

/* RCHFLX_out(:)%ROUTE(method)%<field> / %BASIN_* after the last step, caller's reach order. */
int mr_get_flux(mr_handle h, int method, int field, double *out /* [nRch] */, char *message);

/* Restart-schema state access (read_restart.f90 / write_restart_pio.f90). buf sized per the enum above. */
int mr_get_state(mr_handle h, int var, void *buf, long nbytes, char *message);
int mr_set_state(mr_handle h, int var, const void *buf, long nbytes, char *message);
/* set iTime (number of completed steps); a restart sets it > 0 so lakes skip their cold start. */
int mr_set_steps_done(mr_handle h, long steps, char *message);

/* Unit hydrographs produced by process_param.f90 (basinUH :13-92, make_uh :99-262). */
int mr_get_basin_uh(mr_handle h, double *frac_future /* [ntdh_bas] */, char *message);
int mr_get_reach_uh(mr_handle h, int *ntdh /* [nRch] */, double *uh /* [nRch][maxtdh] */, char *message);

/* ---- multi-domain hand-off (mpi_process.f90:1238-1329) -----------------------------------------------------
 * The reference routes tributary domains on every rank, gathers the tributary-outlet fluxes (mpi_comm_river_flux
 * :1747-1967), BASIN_QR (mpi_comm_flux :1632-1742) and KWT wave state (mpi_comm_kwt_state :2501-2720) to rank 0,
 * and routes the mainstem there with the outlets appended as extra reaches (:593-607).  Here:
 *   tributary handle : mr_set_export names the outlet reaches; every routed step leaves one RECORD per outlet in
 *                      the export buffer;
 *   mainstem handle  : mr_set_ghosts (before mr_set_network) names the reaches that are ghosts of outlets routed
 *                      elsewhere; their per-step values are read from the import buffer;
 *   the caller moves export -> import (NCCL send/recv between processes, mr_copy_exchange inside one process).
 * Buffers are device memory: [slot][max_batch][recLen] doubles, recLen = n_routes + 3 + 2*24:
 *   REACH_Q of each route (route_opt order) | BASIN_QR(1) | numWaves | numRouted | QF[24] | TR[24].
 * No message travels back: owners strip their own routed particles (kwt_route.f90:840-844 is a pure function of
 * the outlet's own state). */
int mr_set_ghosts(mr_handle h, int nGhost, const int *ghostSegId, const int *kind /* 1 headwater-like, 2 interior */,
                  const double *totArea, const double *width, char *message);
int mr_set_export(mr_handle h, int nExport, const int *exportSegId, char *message);
/* which: 0 export, 1 import.  dev == NULL: the library allocates.  Otherwise the caller's device buffer is used
 * (e.g. a torch tensor that NCCL sends/receives in place); nbytes must be >= mr_exchange_bytes. */
long mr_exchange_bytes(mr_handle h, int which);
int mr_set_exchange_buffer(mr_handle h, int which, void *dev, long nbytes, char *message);
int mr_get_exchange_buffer(mr_handle h, int which, void **dev, long *nbytes, char *message);
/* device-to-device: nSlots export records of src (from srcSlot0) -> import records of dst (from dstSlot0) */
int mr_copy_exchange(mr_handle src, mr_handle dst, int srcSlot0, int dstSlot0, int nSlots, char *message);

/* Launch on the caller's CUDA stream (a cudaStream_t passed as void*; NULL restores the handle's own stream),
 * so callers can bracket the work with their own events and overlap it with other streams. */
int mr_set_stream(mr_handle h, void *cuda_stream, char *message);

/* Algorithmic-traffic accounting (bench only): when enabled the KWT kernel also accumulates, per reach, the
 * number of wave particles it reads (own + upstream) and writes.  mr_get_info(MR_INFO_KWT_TOUCHED) returns and
 * clears the total. */
int mr_set_counting(mr_handle h, int enabled, char *message);

long mr_get_info(mr_handle h, int key);
/* device milliseconds of the last batch call, GPTL-region style (mpi_process.f90:1184-1339):
 * [0] whole call, [1] basin2reach+hillslope UH, [2] route_network (all methods), [3] H2D, [4] D2H,
 * [5..7] route_network of the 1st..3rd method of route_opt */
int mr_get_timing(mr_handle h, double *ms /* [8] */);
/* Self-test: y[i] = x[i]**((ALFA-1)/ALFA) as the KWT kernels evaluate it on the device (mr_pow04, kwt_route.f90:1296), for the
 * accuracy test of that routine (tests/test_fastpow.py).  Uses the handle's device and stream only. */
int mr_selftest_pow04(mr_handle h, int n, const double *x, double *y, char *message);

void mr_destroy(mr_handle h);

#ifdef __cplusplus
}
#endif
#endif
