/*
 * mr_oracle.c -- CPU restatement of mizuRoute's per-timestep reach-routing path.
 *
 * THIS FILE IS TEST INFRASTRUCTURE.  It is the checker the CUDA path is compared
 * against (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference
 * legs).  Nothing under mizuroute_b200/ may include, link, import or call it.
 *
 * PARITY UNPINNED: the reference (ESCOMP/mizuRoute, Fortran) ships no golden vectors,
 * unit tests or fixtures for this path and cannot be compiled in this environment
 * (no Fortran compiler / MPI / netCDF).  This restatement follows the Fortran line by
 * line (citations below, paths relative to /root/reference/route/build/src) and is
 * cross-checked against an independently written Python twin (oracle/twin.py) and
 * the invariants in tests/test_oracle_invariants.py.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (see oracle/Makefile).
 * -ffp-contract=off keeps a*b+c as two roundings, as gfortran does on baseline x86-64.
 *
 * Reference map
 *   gamma functions            gamma_func.f90:16-121
 *   hillslope UH (FRAC_FUTURE) process_param.f90:13-92
 *   reach UH (make_uh)         process_param.f90:99-262
 *   topology / areas / goodBas network_topo.f90:46-196,202-311,637-779,958-985
 *   width, slope floor         process_ntopo.f90:176-187,359-366
 *   basin2reach                process_remap.f90:319-422
 *   hillslope convolution      basinUH.f90:70-178
 *   main_route step order      main_route.f90:104-266
 *   accum_inst_runoff          accum_runoff.f90:32-93
 *   irf_rch / conv_upsbas_qr   irf_route.f90:40-264
 *   kwt_rch and callees        kwt_route.f90:36-1622
 *   lake_route (endorheic,Doll)lake_route.f90:28-229,466-470
 *   comp_reach_wb              water_balance.f90:22-112
 */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXQPAR 20                 /* public_var.f90:37 */
#define KW_CAP  24                 /* per-reach particle capacity (reference can hold 0:NQ2+1 <= 21) */
#define VERYSMALL DBL_MIN          /* tiny(1.0_dp), public_var.f90:29 */
#define HUGE_DP   DBL_MAX          /* huge(1.0_dp) */
#define MIN_SLOPE 1.e-6            /* public_var.f90:30 */
#define NEG_RUNOFF_TOL (-1.e-3)    /* public_var.f90:31 */
#define LAKE_WB_TOL 2.e-2
#define PI_MR 3.14159265359        /* public_var.f90:15 */
#define SECPRDAY 86400.0

/* routing method ids = the digits of <route_opt> (public_var.f90:74-80) */
enum { M_SUM = 0, M_IRF = 1, M_KWT = 2, M_KW = 3, M_MC = 4, M_DW = 5, N_METHOD = 6 };
/* computational molecules of the Euler schemes (init_model_data.f90:386-393) */
static const int N_MOLECULE[N_METHOD] = {0, 0, 0, 20, 2, 20};
#define MAX_MOLECULE 20
#define HIGH_DEPTH 100000.0        /* globalData.f90:189 */

/* path-coverage counters (tests assert that thinning / shock merging / disaggregation were exercised);
   0 remove_rch calls, 1 shock merges, 2 merged-and-exited, 3 disaggregated groups, 4 max merged series length,
   5 duplicate times skipped in the merge */
static long g_cnt[8];
long mro_counter(int k) { return g_cnt[k]; }
void mro_reset_counters(void) { int k; for (k = 0; k < 8; k++) g_cnt[k] = 0; }
enum { LAKE_ENDORHEIC = 0, LAKE_DOLL03 = 1, LAKE_H06 = 2, LAKE_HYPE = 3 };

typedef struct {
    int n;                         /* number of entries KWAVE(0:n-1); 0 == not allocated */
    double QF[KW_CAP], TI[KW_CAP], TR[KW_CAP];
    unsigned char RF[KW_CAP];
} kwave_t;

typedef struct {
    /* sizes */
    int nRch, nHRU;
    /* options (control-file keys, read_control.f90; public_var.f90:100-145) */
    double dt;
    int doesBasinRoute, hw_drain_point, is_lake_sim, lakeRegulate, LakeInputOption;
    double min_length_route, runoffMin, time_conv, length_conv;
    int onRoute[N_METHOD];
    int nRoutes, routeOrder[N_METHOD];
    /* spatially constant parameters (param.nml) */
    double fshape, tscale, velo, diff, mann_n, wscale;
    /* topology */
    int *segId, *downSegId, *downIndex;
    int *up_ptr, *up_idx; unsigned char *goodBas; int *nGood;
    int *hru_ptr, *hru_idx; double *hru_wgt;
    int *order;                    /* a valid upstream->downstream processing order */
    int nLevel, *lev_ptr, *lev_idx; /* level sets for OpenMP sweep */
    /* reach parameters */
    double *RLENGTH, *R_SLOPE, *R_WIDTH, *R_MAN_N, *BASAREA, *UPSAREA, *TOTAREA;
    /* channel geometry of the Euler schemes (process_ntopo.f90:174-203); <floodplain>, dscale, floodplainSlope (globalData.f90:187-188) */
    double *R_DEPTH, *SIDE_SLOPE, *FLDP_SLOPE, *R_STORAGE;
    int floodplain; double dscale, floodplainSlope;
    unsigned char *isLake, *lakeInlet; int *lakeModelType;
    double *D03_MaxStorage, *D03_Coefficient, *D03_Power, *D03_S0;
    /* unit hydrographs */
    int ntdh_bas; double *FRAC_FUTURE;
    int *uh_ptr; double *uh_val; int maxtdh;
    /* fluxes / state */
    double *BASIN_QI, *BASIN_QR0, *BASIN_QR1, *QFUTURE; unsigned char *qfuture_alloc;
    double *REACH_Q[N_METHOD], *REACH_VOL0[N_METHOD], *REACH_VOL1[N_METHOD];
    double *REACH_INFLOW[N_METHOD], *WB[N_METHOD];
    double *QFUTURE_IRF;
    kwave_t *KW;
    double *MOL[N_METHOD];         /* molecule%Q of KW / MC / DW, [nRch][N_MOLECULE] */
    double *FLOOD_VOL1[N_METHOD], *REACH_ELE[N_METHOD];
    double *reachRunoff;
    /* lake forcing: reach-level evaporation / precipitation [m3/s] of the current step (RCHFLX%basinevapo / basinprecip,
       main_route.f90:174-199,243-249); hasEP = 0: no such forcing given, both are exactly zero */
    double *reachEvapo, *reachPrecip; int hasEP;
    /* water management (is_flux_wm / is_vol_wm, main_route.f90:110-123): abstraction (+) / injection (-) per reach, target
       lake volumes, the volume jump start, and REACH_WM_FLUX_actual per method; wmFlux = NULL: is_flux_wm off (flux 0) */
    const double *wmFlux, *wmVol; int volJumpStart; double *WM_ACTUAL[N_METHOD];
    /* data assimilation by direct insertion (qmodOption = 1; main_route.f90:124-148, data_assimilation.f90): last observed
       discharge and steps since then per reach, discharge error per reach and method */
    int qmodOption, qBlendPeriod, QerrTrend; double *Qobs, *Qerror[N_METHOD]; int *Qelapsed;
    int obsNow; const double *obsRow;
    /* parametric lake models beyond Doll-2003: per-reach parameters by name (dataTypes.f90:202-254) and the simulation
       start datetime (simDatetime(1) of step 1) the HYPE / Hanasaki formulations read the calendar from */
    double *LP[64];
    /* Hanasaki-2006 inflow memory QPASTUP_IRF(12, past_length) of every lake reach (shared by the routing methods, as the
       monthly means H06_I_* and H06_E_rel_ini it feeds are: they live in RPARAM / RCHFLX, not in ROUTE(:)) */
    double **h06Mem; int *h06Len;
    int hasStart, startYear, startMonth, startDay, noleap; double startSec;
    long iTime;                    /* globalData iTime, 1 on the first step */
    int nThreads;
    char message[256];
} mro_t;

/* ------------------------------------------------------------------------------------------ */
/* gamma_func.f90                                                                              */
/* ------------------------------------------------------------------------------------------ */
static double gammln(double xx)    /* gamma_func.f90:104-121 */
{
    static const double coef[6] = {76.18009172947146, -86.50532032941677, 24.01409824083091,
                                   -1.231739572450155, 0.1208650973866179e-2, -0.5395239384953e-5};
    const double stp = 2.5066282746310005;
    double x = xx, tmp, ser = 0.0, y;
    int j;
    tmp = x + 5.5;
    tmp = (x + 0.5) * log(tmp) - tmp;
    /* sum(coef(:)/arth(x+1,1,6)): arth accumulates by repeated addition (nr_utils.f90:70-82) */
    y = x + 1.0;
    for (j = 0; j < 6; j++) { ser += coef[j] / y; y = y + 1.0; }
    return tmp + log(stp * (1.000000000190015 + ser) / x);
}

static double gser(double a, double x)   /* gamma_func.f90:30-60 */
{
    const int ITMAX = 100; const double EPS = DBL_EPSILON;
    double ap, del, summ; int n;
    if (x == 0.0) return 0.0;
    ap = a; summ = 1.0 / a; del = summ;
    for (n = 1; n <= ITMAX; n++) {
        ap = ap + 1.0;
        del = del * x / ap;
        summ = summ + del;
        if (fabs(del) < fabs(summ) * EPS) break;
    }
    return summ * exp(-x + a * log(x) - gammln(a));
}

static double gcf(double a, double x)    /* gamma_func.f90:65-99 */
{
    const int ITMAX = 100; const double EPS = DBL_EPSILON, FPMIN = DBL_MIN / DBL_EPSILON;
    double an, b, c, d, del, h; int i;
    if (x == 0.0) return 1.0;
    b = x + 1.0 - a; c = 1.0 / FPMIN; d = 1.0 / b; h = d;
    for (i = 1; i <= ITMAX; i++) {
        an = -i * (i - a);
        b = b + 2.0;
        d = an * d + b;
        if (fabs(d) < FPMIN) d = FPMIN;
        c = b + an / c;
        if (fabs(c) < FPMIN) c = FPMIN;
        d = 1.0 / d;
        del = d * c;
        h = h * del;
        if (fabs(del - 1.0) <= EPS) break;
    }
    return exp(-x + a * log(x) - gammln(a)) * h;
}

static double gammp(double a, double x)  /* gamma_func.f90:16-25 */
{
    if (x < a + 1.0) return gser(a, x);
    return 1.0 - gcf(a, x);
}

/* ------------------------------------------------------------------------------------------ */
/* process_param.f90:13-92  basinUH -> FRAC_FUTURE                                             */
/* ------------------------------------------------------------------------------------------ */
static int make_basin_uh(double dt, double fshape, double tscale, int *ntdh_out, double **frac_out)
{
    const int MAXTRY = 100;
    double ntdh_min, ntdh_max, ntdh_try, x_value, cumprob, psave, tfuture, s;
    int itry, ntdh, jtim; double *ff;
    x_value = dt / tscale;
    cumprob = gammp(fshape, x_value);
    if (cumprob > 0.999) {
        ntdh_try = 1.999;
    } else {
        ntdh_min = 1.0; ntdh_max = 1000.0;
        ntdh_try = 0.5 * (ntdh_min + ntdh_max);
        for (itry = 1; itry <= MAXTRY; itry++) {
            x_value = dt * ntdh_try / tscale;
            cumprob = gammp(fshape, x_value);
            if (cumprob < 0.99) ntdh_min = ntdh_try;
            if (cumprob > 0.999) ntdh_max = ntdh_try;
            if (cumprob > 0.99 && cumprob < 0.999) break;
            ntdh_try = 0.5 * (ntdh_min + ntdh_max);
            if (itry == MAXTRY) return 20;
        }
    }
    ntdh = (int)ceil(ntdh_try);
    ff = (double *)malloc(sizeof(double) * ntdh);
    psave = 0.0;
    for (jtim = 1; jtim <= ntdh; jtim++) {
        tfuture = (double)jtim * dt;
        cumprob = gammp(fshape, tfuture / tscale);
        ff[jtim - 1] = fmax(0.0, cumprob - psave);
        psave = cumprob;
    }
    s = 0.0; for (jtim = 0; jtim < ntdh; jtim++) s += ff[jtim];
    for (jtim = 0; jtim < ntdh; jtim++) ff[jtim] = ff[jtim] / s;
    *ntdh_out = ntdh; *frac_out = ff;
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* process_param.f90:99-262  make_uh for one segment; returns ntdh, writes into out[<=240]      */
/* ------------------------------------------------------------------------------------------ */
static int make_uh_one(double seg_length, double dt, double velo, double diff, double *out)
{
    enum { nTMAX = 240, nHr = 240 };
    const double dTUH = 3600.0;
    const double thr1 = (double)0.99999f;   /* single-precision literals, process_param.f90:205,211 */
    const double thr2 = (double)0.9999f;    /* process_param.f90:242 */
    double UHM[nHr + 1], UHQ[nTMAX + 1], fr[nTMAX + 1];
    double INTE, sec, POT, H, UHQ0, d;
    int nTsub, iHr, jHr, iHrStrt = 1, iHrLast = 1, ntdh, iTagg, k;

    nTsub = (int)ceil(dt / dTUH);
    for (k = 1; k <= nTMAX; k++) fr[k] = 0.0;
    for (k = 1; k <= nTsub && k <= nTMAX; k++) fr[k] = 1.0 / nTsub;

    INTE = 0.0; sec = 0.0;
    for (iHr = 1; iHr <= nHr; iHr++) UHM[iHr] = 0.0;
    for (iHr = 1; iHr <= nHr; iHr++) {
        sec = sec + dTUH;
        if (velo > 0.0) {
            d = velo * sec - seg_length;
            POT = (d * d) / (4.0 * diff * sec);
            if (POT > 69.0) H = 0.0;
            else H = 1.0 / (2.0 * sqrt(PI_MR * diff * sec)) * seg_length * exp(-POT);
        } else H = 0.0;
        UHM[iHr] = H;
        INTE = INTE + H;
    }
    if (INTE > 0.0) for (iHr = 1; iHr <= nHr; iHr++) UHM[iHr] = UHM[iHr] / INTE;

    INTE = 0.0;
    for (iHr = 1; iHr <= nTMAX; iHr++) { INTE = INTE + UHM[iHr]; iHrLast = iHr; if (INTE > thr1) break; }
    INTE = 0.0;
    for (iHr = nTMAX; iHr >= 1; iHr--) { INTE = INTE + UHM[iHr]; iHrStrt = iHr; if (INTE > thr1) break; }

    INTE = 0.0;
    for (jHr = 1; jHr <= nTMAX; jHr++) UHQ[jHr] = 0.0;
    for (jHr = 1; jHr <= nTMAX; jHr++) {
        UHQ0 = 0.0;
        for (iHr = iHrStrt; iHr <= iHrLast; iHr++) {
            if ((jHr - iHr) > 0) {
                if (jHr - iHr <= nTsub) UHQ0 = UHQ0 + fr[jHr - iHr] * UHM[iHr];
            } else break;
        }
        UHQ[jHr] = UHQ0;
        INTE = INTE + UHQ0;
    }
    if (INTE > 0.0) for (jHr = 1; jHr <= nTMAX; jHr++) UHQ[jHr] = UHQ[jHr] / INTE;

    INTE = 0.0;
    for (iHr = 1; iHr <= nTMAX; iHr++) { INTE = INTE + UHQ[iHr]; iHrLast = iHr; if (INTE > thr2) break; }
    for (iHr = 1; iHr <= nTMAX; iHr++) UHQ[iHr] = UHQ[iHr] / INTE;

    ntdh = (iHrLast + nTsub - 1) / nTsub;
    for (k = 0; k < ntdh; k++) out[k] = 0.0;
    for (jHr = 1; jHr <= iHrLast; jHr++) {
        iTagg = (jHr + nTsub - 1) / nTsub;
        out[iTagg - 1] = out[iTagg - 1] + UHQ[jHr];
    }
    return ntdh;
}

/* ------------------------------------------------------------------------------------------ */
/* topology helpers                                                                            */
/* ------------------------------------------------------------------------------------------ */
typedef struct { int id, ix; } idix_t;
static int cmp_idix(const void *a, const void *b)
{
    const idix_t *x = (const idix_t *)a, *y = (const idix_t *)b;
    if (x->id != y->id) return (x->id < y->id) ? -1 : 1;
    return (x->ix < y->ix) ? -1 : (x->ix > y->ix);
}
/* network_topo.f90:362-464 downReachIndex: index of the segment whose id == downId (ids <= 0 => none) */
static void down_index(int nUp, int nSeg, const int *segId, const int *downId, int *out)
{
    idix_t *s = (idix_t *)malloc(sizeof(idix_t) * (nSeg > 0 ? nSeg : 1));
    int i;
    for (i = 0; i < nSeg; i++) { s[i].id = segId[i]; s[i].ix = i; }
    qsort(s, nSeg, sizeof(idix_t), cmp_idix);
    for (i = 0; i < nUp; i++) {
        int lo = 0, hi = nSeg - 1, f = -1, id = downId[i];
        out[i] = -1;
        if (id <= 0) continue;
        while (lo <= hi) { int m = (lo + hi) / 2; if (s[m].id < id) lo = m + 1; else if (s[m].id > id) hi = m - 1; else { f = m; hi = m - 1; } }
        if (f >= 0) out[i] = s[f].ix;
    }
    free(s);
}

#define ALLOC(p, n) do { (p) = calloc((size_t)((n) > 0 ? (n) : 1), sizeof(*(p))); } while (0)

void mro_destroy(mro_t *h);
void mro_set_channel(mro_t *h, int floodplain, double dscale, double floodplainSlope);
static const char *LP_NAMES[] = {"HYP_E_emr", "HYP_E_lim", "HYP_E_min", "HYP_E_zero", "HYP_Qrate_emr", "HYP_Erate_emr", "HYP_Qrate_prim",
                                 "HYP_Qrate_amp", "HYP_Qrate_phs", "HYP_prim_F", "HYP_A_avg", "HYP_Qsim_mode",
                                 "H06_Smax", "H06_alpha", "H06_envfact", "H06_S_ini", "H06_c1", "H06_c2", "H06_exponent", "H06_denominator",
                                 "H06_c_compare", "H06_frac_Sdead", "H06_E_rel_ini",
                                 "H06_I_Jan", "H06_I_Feb", "H06_I_Mar", "H06_I_Apr", "H06_I_May", "H06_I_Jun", "H06_I_Jul", "H06_I_Aug", "H06_I_Sep",
                                 "H06_I_Oct", "H06_I_Nov", "H06_I_Dec",
                                 "H06_D_Jan", "H06_D_Feb", "H06_D_Mar", "H06_D_Apr", "H06_D_May", "H06_D_Jun", "H06_D_Jul", "H06_D_Aug", "H06_D_Sep",
                                 "H06_D_Oct", "H06_D_Nov", "H06_D_Dec",
                                 "H06_purpose", "H06_I_mem_F", "H06_D_mem_F", "H06_I_mem_L", "H06_D_mem_L", "LakeTargVol", NULL};
enum { LP_HYP_E_emr, LP_HYP_E_lim, LP_HYP_E_min, LP_HYP_E_zero, LP_HYP_Qrate_emr, LP_HYP_Erate_emr, LP_HYP_Qrate_prim,
       LP_HYP_Qrate_amp, LP_HYP_Qrate_phs, LP_HYP_prim_F, LP_HYP_A_avg, LP_HYP_Qsim_mode, LP_COUNT,
       LP_H06_Smax = LP_COUNT, LP_H06_alpha, LP_H06_envfact, LP_H06_S_ini, LP_H06_c1, LP_H06_c2, LP_H06_exponent, LP_H06_denominator,
       LP_H06_c_compare, LP_H06_frac_Sdead, LP_H06_E_rel_ini, LP_H06_I_Jan, LP_H06_D_Jan = LP_H06_I_Jan + 12,
       LP_H06_purpose = LP_H06_D_Jan + 12, LP_H06_I_mem_F, LP_H06_D_mem_F, LP_H06_I_mem_L, LP_H06_D_mem_L, LP_END,
       LP_LakeTargVol = LP_END /* NETOPO%LakeTargVol: the lake follows the target volume REACH_WM_VOL (lake_route.f90:196-203) */ };

/* ------------------------------------------------------------------------------------------ */
/* create: read_streamSeg.f90 inputs -> augment_ntopo (process_ntopo.f90:39-266) -> put_data_struct */
/* ------------------------------------------------------------------------------------------ */
mro_t *mro_create(int nRch, int nHRU,
                  const int *segId, const int *downSegId,
                  const int *hruSegId, const double *hruArea,
                  const double *length, const double *slope,
                  const double *width_in,      /* NULL -> wscale*sqrt(totalArea) */
                  const double *man_n_in,      /* NULL -> mann_n */
                  const int *islake_in,        /* NULL -> none */
                  const int *lakeModelType_in, /* NULL -> Doll */
                  const double *D03_MaxStorage, const double *D03_Coefficient,
                  const double *D03_Power, const double *D03_S0,
                  double dt, const char *route_opt,
                  int doesBasinRoute, int hw_drain_point, double min_length_route,
                  int is_lake_sim, int lakeRegulate, int LakeInputOption,
                  double runoffMin, double time_conv, double length_conv,
                  double fshape, double tscale, double velo, double diff, double mann_n, double wscale,
                  int nThreads)
{
    mro_t *h = (mro_t *)calloc(1, sizeof(mro_t));
    int i, k, m, *cnt, *hruSegIx, *indeg, *queue, qh, qt, *lev;
    const char *c;
    h->nRch = nRch; h->nHRU = nHRU; h->dt = dt;
    h->doesBasinRoute = doesBasinRoute; h->hw_drain_point = hw_drain_point;
    h->min_length_route = min_length_route; h->is_lake_sim = is_lake_sim; h->lakeRegulate = lakeRegulate;
    h->LakeInputOption = LakeInputOption; h->runoffMin = runoffMin;
    h->time_conv = time_conv; h->length_conv = length_conv;
    h->fshape = fshape; h->tscale = tscale; h->velo = velo; h->diff = diff; h->mann_n = mann_n; h->wscale = wscale;
    h->nThreads = nThreads > 0 ? nThreads : 1;
    h->iTime = 1;
    /* read_control.f90:583-597: digits of route_opt, in order */
    h->nRoutes = 0;
    for (c = route_opt; *c; c++) {
        int id = *c - '0', mm = -1;
        if (id >= 0 && id < N_METHOD) mm = id;
        if (mm < 0 || h->onRoute[mm]) { free(h); return NULL; }
        h->onRoute[mm] = 1; h->routeOrder[h->nRoutes++] = mm;
    }

    ALLOC(h->segId, nRch); ALLOC(h->downSegId, nRch); ALLOC(h->downIndex, nRch);
    memcpy(h->segId, segId, sizeof(int) * nRch); memcpy(h->downSegId, downSegId, sizeof(int) * nRch);
    down_index(nRch, nRch, segId, downSegId, h->downIndex);

    /* up2downSegment (network_topo.f90:202-311): upstream lists filled in reach-index order */
    ALLOC(h->up_ptr, nRch + 1); ALLOC(cnt, nRch);
    for (i = 0; i < nRch; i++) if (h->downIndex[i] >= 0) h->up_ptr[h->downIndex[i] + 1]++;
    for (i = 0; i < nRch; i++) h->up_ptr[i + 1] += h->up_ptr[i];
    ALLOC(h->up_idx, h->up_ptr[nRch]); ALLOC(h->goodBas, h->up_ptr[nRch]); ALLOC(h->nGood, nRch);
    for (i = 0; i < nRch; i++) { int d = h->downIndex[i]; if (d >= 0) h->up_idx[h->up_ptr[d] + cnt[d]++] = i; }
    free(cnt);

    /* hru2segment (network_topo.f90:46-196): contributing HRUs in HRU-index order, weight = area/sum(area) */
    ALLOC(hruSegIx, nHRU);
    down_index(nHRU, nRch, segId, hruSegId, hruSegIx);
    ALLOC(h->hru_ptr, nRch + 1); ALLOC(cnt, nRch);
    for (i = 0; i < nHRU; i++) if (hruSegIx[i] >= 0) h->hru_ptr[hruSegIx[i] + 1]++;
    for (i = 0; i < nRch; i++) h->hru_ptr[i + 1] += h->hru_ptr[i];
    ALLOC(h->hru_idx, h->hru_ptr[nRch]); ALLOC(h->hru_wgt, h->hru_ptr[nRch]);
    for (i = 0; i < nHRU; i++) { int s = hruSegIx[i]; if (s >= 0) h->hru_idx[h->hru_ptr[s] + cnt[s]++] = i; }
    free(cnt); free(hruSegIx);

    /* a topological order (results do not depend on which valid order is used) + level sets */
    ALLOC(h->order, nRch); ALLOC(indeg, nRch); ALLOC(queue, nRch); ALLOC(lev, nRch);
    for (i = 0; i < nRch; i++) indeg[i] = h->up_ptr[i + 1] - h->up_ptr[i];
    qh = qt = 0;
    for (i = 0; i < nRch; i++) if (indeg[i] == 0) { queue[qt++] = i; lev[i] = 0; }
    h->nLevel = 0;
    while (qh < qt) {
        int r = queue[qh++], d = h->downIndex[r];
        if (lev[r] + 1 > h->nLevel) h->nLevel = lev[r] + 1;
        if (d >= 0) { if (lev[r] + 1 > lev[d]) lev[d] = lev[r] + 1; if (--indeg[d] == 0) queue[qt++] = d; }
    }
    if (qt != nRch) { snprintf(h->message, 256, "mro_create/network has a cycle"); }
    memcpy(h->order, queue, sizeof(int) * nRch);
    ALLOC(h->lev_ptr, h->nLevel + 1); ALLOC(h->lev_idx, nRch);
    for (i = 0; i < qt; i++) h->lev_ptr[lev[queue[i]] + 1]++;
    for (k = 0; k < h->nLevel; k++) h->lev_ptr[k + 1] += h->lev_ptr[k];
    ALLOC(cnt, h->nLevel);
    for (i = 0; i < qt; i++) { int r = queue[i]; h->lev_idx[h->lev_ptr[lev[r]] + cnt[lev[r]]++] = r; }
    free(cnt); free(indeg); free(lev);

    /* reach_list (network_topo.f90:637-779): basArea, upsArea, totalArea, goodBas */
    ALLOC(h->RLENGTH, nRch); ALLOC(h->R_SLOPE, nRch); ALLOC(h->R_WIDTH, nRch); ALLOC(h->R_MAN_N, nRch);
    ALLOC(h->BASAREA, nRch); ALLOC(h->UPSAREA, nRch); ALLOC(h->TOTAREA, nRch);
    for (k = 0; k < qt; k++) {
        int r = queue[k]; double ups = 0.0, bas = 0.0;
        for (m = h->up_ptr[r]; m < h->up_ptr[r + 1]; m++) ups = ups + h->TOTAREA[h->up_idx[m]];
        for (m = h->hru_ptr[r]; m < h->hru_ptr[r + 1]; m++) bas += hruArea[h->hru_idx[m]];
        h->UPSAREA[r] = ups; h->BASAREA[r] = bas; h->TOTAREA[r] = bas + ups;
        for (m = h->hru_ptr[r]; m < h->hru_ptr[r + 1]; m++) h->hru_wgt[m] = hruArea[h->hru_idx[m]] / bas;
        h->nGood[r] = 0;
        for (m = h->up_ptr[r]; m < h->up_ptr[r + 1]; m++) { h->goodBas[m] = (h->TOTAREA[r] > VERYSMALL); h->nGood[r] += h->goodBas[m]; }
    }
    free(queue);

    /* geometry (process_ntopo.f90:176-187) and put_data_struct (process_ntopo.f90:359-366) */
    for (i = 0; i < nRch; i++) {
        h->RLENGTH[i] = length[i];
        h->R_SLOPE[i] = fmax(slope[i], MIN_SLOPE);
        h->R_WIDTH[i] = width_in ? width_in[i] : wscale * sqrt(h->TOTAREA[i]);
        h->R_MAN_N[i] = man_n_in ? man_n_in[i] : mann_n;
    }
    ALLOC(h->R_DEPTH, nRch); ALLOC(h->SIDE_SLOPE, nRch); ALLOC(h->FLDP_SLOPE, nRch); ALLOC(h->R_STORAGE, nRch);
    mro_set_channel(h, 0, (double)0.000045f, 1000.0);   /* default-real literal, globalData.f90:187 */
    /* lakes (process_ntopo.f90:476-485; network_topo.f90:958-985) */
    ALLOC(h->isLake, nRch); ALLOC(h->lakeInlet, nRch); ALLOC(h->lakeModelType, nRch);
    ALLOC(h->D03_MaxStorage, nRch); ALLOC(h->D03_Coefficient, nRch); ALLOC(h->D03_Power, nRch); ALLOC(h->D03_S0, nRch);
    if (is_lake_sim) {
        for (i = 0; i < nRch; i++) {
            h->isLake[i] = islake_in ? (islake_in[i] == 1) : 0;
            h->lakeModelType[i] = (!lakeRegulate || !lakeModelType_in) ? LAKE_DOLL03 : lakeModelType_in[i];
            if (D03_MaxStorage) h->D03_MaxStorage[i] = D03_MaxStorage[i];
            if (D03_Coefficient) h->D03_Coefficient[i] = D03_Coefficient[i];
            if (D03_Power) h->D03_Power[i] = D03_Power[i];
            if (D03_S0) h->D03_S0[i] = D03_S0[i];
        }
        for (i = 0; i < nRch; i++) { int d = h->downIndex[i]; h->lakeInlet[i] = (d >= 0 && h->isLake[d]); }
    }

    /* unit hydrographs */
    if (make_basin_uh(dt, fshape, tscale, &h->ntdh_bas, &h->FRAC_FUTURE) != 0) { mro_destroy(h); return NULL; }
    ALLOC(h->uh_ptr, nRch + 1);
    if (h->onRoute[M_IRF]) {
        /* make_uh is init-time work, independent per reach: built with all host threads (values do not depend on it) */
        int tot = 0, mx = 0;
        double *all = (double *)malloc(sizeof(double) * 240 * (size_t)(nRch > 0 ? nRch : 1));
#pragma omp parallel for schedule(dynamic, 1024) reduction(max : mx)
        for (i = 0; i < nRch; i++) {
            int n = make_uh_one(length[i], dt, velo, diff, all + (size_t)i * 240), kk;
            /* process_ntopo.f90:496-499: lake UH is an impulse (islake is only set when is_lake_sim) */
            if (h->isLake[i]) { for (kk = 0; kk < n; kk++) all[(size_t)i * 240 + kk] = 0.0; all[(size_t)i * 240] = 1.0; }
            h->uh_ptr[i + 1] = n;
            if (n > mx) mx = n;
        }
        h->maxtdh = mx;
        for (i = 0; i < nRch; i++) { tot += h->uh_ptr[i + 1]; h->uh_ptr[i + 1] += h->uh_ptr[i]; }
        ALLOC(h->uh_val, tot); ALLOC(h->QFUTURE_IRF, tot);
        for (i = 0; i < nRch; i++) memcpy(h->uh_val + h->uh_ptr[i], all + (size_t)i * 240, sizeof(double) * (h->uh_ptr[i + 1] - h->uh_ptr[i]));
        free(all);
    }

    /* cold start (init_model_data.f90:399-463,600) */
    ALLOC(h->BASIN_QI, nRch); ALLOC(h->BASIN_QR0, nRch); ALLOC(h->BASIN_QR1, nRch);
    ALLOC(h->QFUTURE, (size_t)nRch * h->ntdh_bas); ALLOC(h->qfuture_alloc, nRch);
    ALLOC(h->reachRunoff, nRch); ALLOC(h->reachEvapo, nRch); ALLOC(h->reachPrecip, nRch);
    ALLOC(h->Qobs, nRch); ALLOC(h->Qelapsed, nRch);
    for (m = 0; m < N_METHOD; m++) {
        ALLOC(h->REACH_Q[m], nRch); ALLOC(h->REACH_VOL0[m], nRch); ALLOC(h->REACH_VOL1[m], nRch);
        ALLOC(h->REACH_INFLOW[m], nRch); ALLOC(h->WB[m], nRch);
        ALLOC(h->FLOOD_VOL1[m], nRch); ALLOC(h->REACH_ELE[m], nRch); ALLOC(h->WM_ACTUAL[m], nRch); ALLOC(h->Qerror[m], nRch);
        if (N_MOLECULE[m] > 0 && h->onRoute[m]) ALLOC(h->MOL[m], (size_t)nRch * N_MOLECULE[m]);   /* molecule%Q(:) = 0, init_model_data.f90:463-497 */
    }
    ALLOC(h->KW, nRch);
    if (h->onRoute[M_KWT] && is_lake_sim)
        for (i = 0; i < nRch; i++) if (h->isLake[i]) {
            kwave_t *w = &h->KW[i]; w->n = 1; w->QF[0] = -9999; w->TI[0] = -9999; w->TR[0] = -9999; w->RF[0] = 0;
        }
    return h;
}

void mro_destroy(mro_t *h)
{
    int m;
    if (!h) return;
    free(h->segId); free(h->downSegId); free(h->downIndex); free(h->up_ptr); free(h->up_idx); free(h->goodBas); free(h->nGood);
    free(h->hru_ptr); free(h->hru_idx); free(h->hru_wgt); free(h->order); free(h->lev_ptr); free(h->lev_idx);
    free(h->RLENGTH); free(h->R_SLOPE); free(h->R_WIDTH); free(h->R_MAN_N); free(h->BASAREA); free(h->UPSAREA); free(h->TOTAREA);
    free(h->isLake); free(h->lakeInlet); free(h->lakeModelType);
    free(h->D03_MaxStorage); free(h->D03_Coefficient); free(h->D03_Power); free(h->D03_S0);
    free(h->FRAC_FUTURE); free(h->uh_ptr); free(h->uh_val);
    free(h->BASIN_QI); free(h->BASIN_QR0); free(h->BASIN_QR1); free(h->QFUTURE); free(h->qfuture_alloc); free(h->reachRunoff); free(h->reachEvapo); free(h->reachPrecip);
    for (m = 0; m < N_METHOD; m++) { free(h->REACH_Q[m]); free(h->REACH_VOL0[m]); free(h->REACH_VOL1[m]); free(h->REACH_INFLOW[m]); free(h->WB[m]);
                                     free(h->FLOOD_VOL1[m]); free(h->REACH_ELE[m]); free(h->MOL[m]); free(h->WM_ACTUAL[m]); free(h->Qerror[m]); }
    free(h->Qobs); free(h->Qelapsed);
    { int k; for (k = 0; k < 64; k++) free(h->LP[k]); }
    if (h->h06Mem) { int k; for (k = 0; k < h->nRch; k++) free(h->h06Mem[k]); free(h->h06Mem); free(h->h06Len); }
    free(h->R_DEPTH); free(h->SIDE_SLOPE); free(h->FLDP_SLOPE); free(h->R_STORAGE);
    free(h->QFUTURE_IRF); free(h->KW);
    free(h);
}

/* ------------------------------------------------------------------------------------------ */
/* process_remap.f90:319-422 basin2reach                                                       */
/* ------------------------------------------------------------------------------------------ */
static int basin2reach(mro_t *h, const double *basinRunoff, double *reachRunoff, int limit)
{
    int j, ierr = 0;
#pragma omp parallel for schedule(static) num_threads(h->nThreads) reduction(max : ierr)
    for (j = 0; j < h->nRch; j++) {
        int m, nContrib = h->hru_ptr[j + 1] - h->hru_ptr[j];
        if (nContrib > 0) {
            double r = 0.0;
            for (m = h->hru_ptr[j]; m < h->hru_ptr[j + 1]; m++) {
                double ro = basinRunoff[h->hru_idx[m]];
                if (limit && ro < NEG_RUNOFF_TOL) ierr = 20;
                r = r + h->hru_wgt[m] * ro * h->time_conv * h->length_conv;
            }
            if (limit && r < h->runoffMin) r = h->runoffMin;
            reachRunoff[j] = r * h->BASAREA[j];
        } else {
            if (limit) reachRunoff[j] = h->runoffMin;
        }
    }
    if (ierr) snprintf(h->message, 256, "basin2reach/exceeded negative runoff tolerance");
    return ierr;
}

/* ------------------------------------------------------------------------------------------ */
/* basinUH.f90:70-178 hru_irf + irf_conv                                                       */
/* ------------------------------------------------------------------------------------------ */
static void hru_irf(mro_t *h, int j)
{
    int n = h->ntdh_bas, k; double *qf = h->QFUTURE + (size_t)j * n; double inq = h->BASIN_QI[j];
    int lake = (h->isLake[j] && h->is_lake_sim);
    h->BASIN_QR0[j] = h->BASIN_QR1[j];
    for (k = 0; k < n; k++) {
        double uh = lake ? (k == 0 ? 1.0 : 0.0) : h->FRAC_FUTURE[k];
        qf[k] = qf[k] + uh * inq;
    }
    h->BASIN_QR1[j] = qf[0];
    for (k = 1; k < n; k++) qf[k - 1] = qf[k];
    qf[n - 1] = 0.0;
}

/* water_balance.f90:61-87 (REACH_WM_FLUX = 0; precip/evap = 0 unless LakeInputOption uses them) */
static void comp_reach_wb_lake(mro_t *h, int m, int j, double Qupstream, double Qlat, int lakeFlag)
{
    double dt = h->dt;
    double dVol = h->REACH_VOL1[m][j] - h->REACH_VOL0[m][j];
    double Qin = Qupstream * dt, Qlateral = Qlat * dt;
    double precip = (lakeFlag && h->hasEP) ? h->reachPrecip[j] * dt : 0.0;
    double evapo = (lakeFlag && h->hasEP) ? -1.0 * h->reachEvapo[j] * dt : 0.0;
    double Qout = -1.0 * h->REACH_Q[m][j] * dt;
    double Qtake_actual = -1.0 * (h->wmFlux ? h->WM_ACTUAL[m][j] : 0.0) * dt;
    h->WB[m][j] = dVol - (Qin + Qlateral + precip + Qtake_actual + Qout + evapo);
}

/* Water abstraction (+) / injection (-) of a reach, shared by irf_rch, kw_rch, mc_rch and dfw_rch (irf_route.f90:114-142 =
   kwe_route.f90:118-146 = mc_route.f90:118-146 = dfw_route.f90:122-150): taken from the storage first, then from the upstream
   inflow, then from the lateral flow.  REACH_WM_FLUX_actual starts as the demand, missing value included. */
static void wm_cascade(mro_t *h, int M, int j, double q_upstream, double *q_upstream_mod, double *Qlat)
{
    double dt = h->dt, Qabs;
    *q_upstream_mod = q_upstream;
    if (!h->wmFlux) { h->WM_ACTUAL[M][j] = 0.0; return; }
    Qabs = h->wmFlux[j];
    h->WM_ACTUAL[M][j] = h->wmFlux[j];
    if (h->wmFlux[j] == -9999.0) return;                   /* realMissing: no water management at this reach */
    if (Qabs > 0) {
        if (h->REACH_VOL1[M][j] / dt > Qabs) {
            h->REACH_VOL1[M][j] = h->REACH_VOL1[M][j] - Qabs * dt;
        } else {
            Qabs = Qabs - h->REACH_VOL1[M][j] / dt;
            h->REACH_VOL1[M][j] = 0.0;
            if (q_upstream > Qabs) {
                *q_upstream_mod = q_upstream - Qabs;
            } else {
                Qabs = Qabs - q_upstream;
                *q_upstream_mod = 0.0;
                if (*Qlat > Qabs) {
                    *Qlat = *Qlat - Qabs;
                } else {
                    Qabs = Qabs - *Qlat;
                    *Qlat = 0.0;
                    h->WM_ACTUAL[M][j] = h->wmFlux[j] - Qabs;
                }
            }
        }
    } else {
        *Qlat = *Qlat - Qabs;
    }
}

/* data_assimilation.f90:23-97: pull REACH_Q towards the last observation, the correction fading over qBlendPeriod steps */
static int direct_insertion(mro_t *h, int M, int j)
{
    double *Qerror = &h->Qerror[M][j], Qcorrect, k, x0, y0;
    const int Qelapsed = h->Qelapsed[j], blend = h->qBlendPeriod;
    if (h->Qobs[j] > 0.0) *Qerror = h->REACH_Q[M][j] - h->Qobs[j];
    if (Qelapsed > blend) *Qerror = 0.0;
    if (Qelapsed <= blend) {
        switch (h->QerrTrend) {
            case 1: Qcorrect = *Qerror; break;
            case 2: Qcorrect = *Qerror * (1.0 - (double)Qelapsed / (double)blend); break;
            case 3:
                x0 = 0.25; y0 = (double)0.90f;       /* single-precision literals, data_assimilation.f90:76 */
                k = log(1.0 / y0 - 1.0) / (blend / 2.0 - blend * x0);
                Qcorrect = *Qerror / (1.0 + exp(-k * (1.0 * Qelapsed - blend / 2.0)));
                break;
            case 4:
                if (*Qerror != 0.0) { k = log(0.1 / fabs(*Qerror)) / (1.0 * blend); Qcorrect = *Qerror * exp(k * Qelapsed); }
                else Qcorrect = 0.0;
                break;
            default: snprintf(h->message, 256, "direct_insertion/discharge error trend model must be 1(const),2(liear), or 3(logistic)"); return 81;
        }
    } else Qcorrect = 0.0;
    h->REACH_Q[M][j] = fmax(h->REACH_Q[M][j] - Qcorrect, 0.0);
    return 0;
}

/* river reaches: direct insertion when qmodOption = 1, else the water balance (irf_route.f90:188-202 and the Euler schemes) */
static int finish_reach(mro_t *h, int m, int j, double Qupstream, double Qlat)
{
    if (h->qmodOption == 1) return direct_insertion(h, m, j);
    comp_reach_wb_lake(h, m, j, Qupstream, Qlat, 0);
    return 0;
}
static void comp_reach_wb(mro_t *h, int m, int j, double Qupstream, double Qlat) { comp_reach_wb_lake(h, m, j, Qupstream, Qlat, 0); }

/* accum_runoff.f90:60-75 */
static int accum_inst_runoff(mro_t *h, int j)
{
    int m; double q_upstream = 0.0;
    h->REACH_Q[M_SUM][j] = h->BASIN_QR1[j];
    if (h->up_ptr[j + 1] > h->up_ptr[j]) {
        for (m = h->up_ptr[j]; m < h->up_ptr[j + 1]; m++) q_upstream = q_upstream + h->REACH_Q[M_SUM][h->up_idx[m]];
        h->REACH_Q[M_SUM][j] = h->REACH_Q[M_SUM][j] + q_upstream;
    }
    return 0;
}

/* irf_route.f90:40-264 */
static int irf_rch(mro_t *h, int j)
{
    const int M = M_IRF;
    int nUps = h->nGood[j], m, k, ntdh; double q_upstream = 0.0, Qlat = 0.0, dt = h->dt, q_mod, q_in;
    double *QF = h->QFUTURE_IRF + h->uh_ptr[j]; const double *UH = h->uh_val + h->uh_ptr[j];
    h->REACH_VOL0[M][j] = h->REACH_VOL1[M][j];
    if (nUps > 0) {
        for (k = 0; k < nUps; k++) {
            m = h->up_ptr[j] + k;
            if (!h->goodBas[m]) continue;
            q_upstream = q_upstream + h->REACH_Q[M][h->up_idx[m]];
        }
        Qlat = h->BASIN_QR1[j];
    } else {
        if (h->hw_drain_point == 1) { q_upstream = q_upstream + h->BASIN_QR1[j]; Qlat = 0.0; }
        else if (h->hw_drain_point == 2) { Qlat = h->BASIN_QR1[j]; }
    }
    h->REACH_INFLOW[M][j] = q_upstream;
    wm_cascade(h, M, j, q_upstream, &q_mod, &Qlat);
    q_in = q_upstream; q_upstream = q_mod;          /* conv_upsbas_qr sees the inflow left after the abstraction */
    ntdh = h->uh_ptr[j + 1] - h->uh_ptr[j];
    if (h->RLENGTH[j] > h->min_length_route) {
        for (k = 0; k < ntdh; k++) QF[k] = QF[k] + UH[k] * q_upstream;
        /* "*0.999" is a default-real (single precision) literal in irf_route.f90:245 and the reference is built without
           -fdefault-real-8 (route/build/Makefile:101): the factor is float32(0.999) = 0.99900001287460327 */
        QF[0] = fmin((fmax(0.0, h->REACH_VOL1[M][j]) / dt + q_upstream) * (double)0.999f, QF[0]);
        h->REACH_VOL1[M][j] = h->REACH_VOL1[M][j] - (QF[0] - q_upstream) * dt;
        h->REACH_Q[M][j] = QF[0] + Qlat;
        for (k = 1; k < ntdh; k++) QF[k - 1] = QF[k];
        QF[ntdh - 1] = 0.0;
    } else {
        for (k = 0; k < ntdh; k++) QF[k] = 0.0;
        QF[0] = q_upstream;
        h->REACH_Q[M][j] = QF[0] + Qlat;
        h->REACH_VOL0[M][j] = 0.0; h->REACH_VOL1[M][j] = 0.0;
    }
    return finish_reach(h, M, j, q_in, Qlat);
}

/* ------------------------------------------------------------------------------------------ */
/* hydraulic.f90: trapezoidal main channel (b, zc) + floodplain (zf) above bankDepth.          */
/* Integer powers are the multiplication chains a compiler expands x**n to.                    */
/* ------------------------------------------------------------------------------------------ */
static double hy_Btop(double y, double b, double zc, double zf, double bd)      /* hydraulic.f90:46-77 */
{
    double B;
    if (y <= bd) return b + 2 * y * zc;
    B = b + 2 * bd * zc;
    B = B + zf * (y - bd) * 2;
    return B;
}
static double hy_Pwet(double y, double b, double zc, double zf, double bd)      /* hydraulic.f90:82-113 */
{
    double P;
    if (y <= bd) return b + 2 * y * sqrt(1 + zc * zc);
    P = b + 2 * bd * sqrt(1 + zc * zc);
    P = P + 2 * (y - bd) * sqrt(1 + zf * zf);
    return P;
}
static double hy_area(double y, double b, double zc, double zf, double bd)      /* flow_area, hydraulic.f90:118-152 */
{
    double A, Bt, Bb;
    if (y <= bd) return y * (b + zc * y);
    A = bd * (b + zc * bd);
    Bt = hy_Btop(y, b, zc, zf, bd); Bb = hy_Btop(bd, b, zc, zf, bd);
    return A + (y - bd) * (Bt + Bb) / 2.0;
}
static double hy_water_height(double area, double b, double zc, double zf, double bd)   /* hydraulic.f90:157-202 */
{
    double A_bank = hy_area(bd, b, zc, zf, bd);
    if (area > A_bank) {
        double Bb = hy_Btop(bd, b, zc, zf, bd);
        double disc = Bb * Bb - 4.0 * zf * (A_bank - area);
        return bd + (-Bb + sqrt(disc)) / (2.0 * zf);
    }
    if (zc == 0) return area / b;
    return (-b + sqrt(b * b + 4.0 * area * zc)) / (2.0 * zc);
}
static double hy_flow_depth(double Q, double b, double zc, double S, double n, double zf, double bd)   /* hydraulic.f90:299-420 */
{
    const double c13 = 1.0 / 3.0, c23 = 2.0 / 3.0, c53 = 5.0 / 3.0, c103 = 10.0 / 3.0, err_thresh = 0.005, Qmin = 1.e-50;
    double error = 100.0, depth = 0.0, y0, Coef1, Coef2, A, P, Bt, hh, dhdy, t;
    if (!(Q > Qmin)) return 0.0;
    {   /* bankDepth is always passed by the routing schemes: floodplain = .true. */
        double Abf = hy_area(bd, b, zc, zf, bd), Pbf = hy_Pwet(bd, b, zc, zf, bd), Bbf = hy_Btop(bd, b, zc, zf, bd);
        double Qbf = Abf * pow(Abf / Pbf, c23) * sqrt(S) / n;
        if (Q < Qbf) {
            t = sqrt(S) / n / Q; Coef1 = t * t * t;
            Coef2 = 2 * sqrt(zc * zc + 1.0);
            y0 = pow(1.0 / Coef1 / (b * b * b), 1.0 / 5.0);
            while (error > err_thresh) {
                double A2, A4, A5;
                A = hy_area(y0, b, zc, zf, bd); Bt = hy_Btop(y0, b, zc, zf, bd); P = hy_Pwet(y0, b, zc, zf, bd);
                A2 = A * A; A4 = A2 * A2; A5 = A4 * A;
                hh = Coef1 * A5 / (P * P) - 1.0;
                dhdy = Coef1 * (5 * A4 * Bt * P - 2 * Coef2 * A5) / (P * P * P);
                depth = y0 - hh / dhdy;
                error = fabs((depth - y0) / depth);
                y0 = depth;
            }
        } else {
            y0 = bd + 2.0;
            Coef1 = sqrt(S) / n / pow(Pbf, c23);
            Coef2 = 2 * pow(zf / 2, c53) * sqrt(S) / n / pow(zf * zf + 1.0, c13);
            while (error > err_thresh) {
                double ye = y0 - bd;
                hh = Coef1 * pow(Abf + Bbf * ye, c53) + Coef2 * pow(ye, c103) / pow(ye, c23) - Q;
                dhdy = Coef1 * c53 * Bbf * pow(Abf + Bbf * ye, c23) + Coef2 * (c103 - c23) * pow(ye, c53);
                depth = y0 - hh / dhdy;
                error = fabs((depth - y0) / depth);
                y0 = depth;
            }
        }
    }
    return depth;
}
static double hy_friction_slope(double Q, double y, double b, double zc, double n, double zf, double bd)
{
    double A = hy_area(y, b, zc, zf, bd), P = hy_Pwet(y, b, zc, zf, bd);
    double t = Q * n / A / pow(A / P, 2.0 / 3.0);
    return t * t;
}
static double hy_celerity(double Q, double y, double b, double zc, double S, double n, double zf, double bd)   /* hydraulic.f90:425-471 */
{
    double Bt, Sf;
    (void)S;
    if (!(y > 0.0)) return 0.0;
    Bt = hy_Btop(y, b, zc, zf, bd);
    Sf = hy_friction_slope(Q, y, b, zc, n, zf, bd);          /* useFrictionSlope = .true. */
    return 5.0 / 3.0 * pow(Sf, 0.3) * pow(Q, 0.4) / pow(Bt, 0.4) / pow(n, 0.6);
}
static double hy_diffusivity(double Q, double y, double b, double zc, double S, double n, double zf, double bd) /* hydraulic.f90:476-522 */
{
    double Bt, Sf;
    (void)S;
    if (!(y > 0.0)) return 0.0;
    Bt = hy_Btop(y, b, zc, zf, bd);
    Sf = hy_friction_slope(Q, y, b, zc, n, zf, bd);
    return fabs(Q) / Sf / Bt / 2.0;
}

/* process_ntopo.f90:174-203: bankfull depth, side / floodplain slopes, bankfull storage */
void mro_set_channel(mro_t *h, int floodplain, double dscale, double floodplainSlope)
{
    int i;
    h->floodplain = floodplain; h->dscale = dscale; h->floodplainSlope = floodplainSlope;
    for (i = 0; i < h->nRch; i++) {
        h->R_DEPTH[i] = floodplain ? dscale * sqrt(h->TOTAREA[i]) : HIGH_DEPTH;
        h->SIDE_SLOPE[i] = 0.0;
        h->FLDP_SLOPE[i] = floodplainSlope;
        h->R_STORAGE[i] = hy_area(h->R_DEPTH[i], h->R_WIDTH[i], h->SIDE_SLOPE[i], h->FLDP_SLOPE[i], h->R_DEPTH[i]) * h->RLENGTH[i];
    }
}

/* upstream discharge and lateral flow of a reach, shared by irf_rch / kw_rch / mc_rch / dfw_rch
   (kwe_route.f90:82-113, mc_route.f90:80-110, dfw_route.f90:86-117) */
static int euler_inflow(mro_t *h, int M, int j, double *q_upstream, double *Qlat)
{
    int nUps = h->nGood[j], m, k, isHW = 1; double qup = 0.0;
    h->REACH_VOL0[M][j] = h->REACH_VOL1[M][j];
    *Qlat = 0.0;
    if (nUps > 0) {
        isHW = 0;
        for (k = 0; k < nUps; k++) {
            m = h->up_ptr[j] + k;
            if (!h->goodBas[m]) continue;
            qup = qup + h->REACH_Q[M][h->up_idx[m]];
        }
        *Qlat = h->BASIN_QR1[j];
    } else {
        if (h->hw_drain_point == 1) { qup = qup + h->BASIN_QR1[j]; *Qlat = 0.0; }
        else if (h->hw_drain_point == 2) { *Qlat = h->BASIN_QR1[j]; }
    }
    h->REACH_INFLOW[M][j] = qup;
    *q_upstream = qup;
    return isHW;
}

/* flood volume and water surface height after the volume update (kwe_route.f90:319-326 and siblings) */
static void euler_stage(mro_t *h, int M, int j)
{
    if (h->REACH_VOL1[M][j] > h->R_STORAGE[j]) h->FLOOD_VOL1[M][j] = h->REACH_VOL1[M][j] - h->R_STORAGE[j];
    else h->FLOOD_VOL1[M][j] = 0.0;
    h->REACH_ELE[M][j] = hy_water_height(h->REACH_VOL1[M][j] / h->RLENGTH[j], h->R_WIDTH[j], h->SIDE_SLOPE[j], h->FLDP_SLOPE[j], h->R_DEPTH[j]);
}

/* advection_diffusion.f90:19-262: solve_ade with its defaults (central difference, Neumann downstream B.C.,
   wck = wdk = 1) + TDMA.  prev/cur hold nMol nodes; node nMol-1 (1-based) is the reach outlet. */
static void solve_ade(double L, int nMol, double dtl, double FluxUp, double ck, double dk, const double *prev, double *cur)
{
    const double wck = 1.0, wdk = 1.0;
    double up[MAX_MOLECULE], di[MAX_MOLECULE], lo[MAX_MOLECULE], b[MAX_MOLECULE], D[MAX_MOLECULE], b1[MAX_MOLECULE];
    const int Nx = nMol - 1;
    const double dx = L / (Nx - 1);
    const double Cd = dk * dtl / (dx * dx), Ca = ck * dtl / dx;
    int i;
    for (i = 0; i < nMol; i++) { up[i] = 0.0; lo[i] = 0.0; }
    di[0] = 1.0;
    for (i = 1; i < nMol - 1; i++) di[i] = 2.0 + 4 * wdk * Cd;
    di[nMol - 1] = 1.0;
    for (i = 2; i < nMol; i++) up[i] = wck * Ca - 2.0 * wdk * Cd;
    for (i = 0; i < nMol - 2; i++) lo[i] = -wck * Ca - 2.0 * wdk * Cd;
    lo[nMol - 2] = -1.0;
    b[0] = FluxUp;
    b[nMol - 1] = prev[nMol - 1] - prev[nMol - 2];
    for (i = 1; i < nMol - 1; i++)
        b[i] = ((1.0 - wck) * Ca + 2.0 * (1.0 - wdk) * Cd) * prev[i - 1] + (2.0 - 4.0 * (1.0 - wdk) * Cd) * prev[i]
             - ((1.0 - wck) * Ca - 2.0 * (1.0 - wdk) * Cd) * prev[i + 1];
    /* TDMA: up[i] = A(i-1,i), lo[i] = A(i+1,i) */
    for (i = 0; i < nMol; i++) { D[i] = di[i]; b1[i] = b[i]; }
    for (i = 1; i < nMol; i++) {
        double coef = lo[i - 1] / D[i - 1];
        D[i] = D[i] - coef * up[i];
        b1[i] = b1[i] - coef * b1[i - 1];
    }
    cur[nMol - 1] = b1[nMol - 1] / D[nMol - 1];
    for (i = nMol - 2; i >= 0; i--) cur[i] = (b1[i] - up[i + 1] * cur[i + 1]) / D[i];
}

/* kwe_route.f90:40-365 (M_KW, dk = 0) and dfw_route.f90:43-372 (M_DW): one implicit Euler step of the linearised
   advection(-diffusion) equation on the reach's molecule */
static int kw_dw_rch(mro_t *h, int M, int j)
{
    const int nMol = N_MOLECULE[M];
    double q_upstream, q_in, Qlat, dt = h->dt, *mol = h->MOL[M] + (size_t)j * nMol, cur[MAX_MOLECULE];
    const int isHW = euler_inflow(h, M, j, &q_in, &Qlat);
    const double S = h->R_SLOPE[j], n = h->R_MAN_N[j], bt = h->R_WIDTH[j], bd = h->R_DEPTH[j], zc = h->SIDE_SLOPE[j], zf = h->FLDP_SLOPE[j], L = h->RLENGTH[j];
    int i;
    wm_cascade(h, M, j, q_in, &q_upstream, &Qlat);      /* the solver sees Qupstream_mod, the water balance Qupstream */
    if (!isHW || h->hw_drain_point == 1) {
        if (L > h->min_length_route) {
            double Qbar = (q_upstream + mol[0] + mol[nMol - 2]) / 3.0;
            double depth = hy_flow_depth(fabs(Qbar), bt, zc, S, n, zf, bd);
            double ck = hy_celerity(fabs(Qbar), depth, bt, zc, S, n, zf, bd);
            double dk = M == M_DW ? hy_diffusivity(fabs(Qbar), depth, bt, zc, S, n, zf, bd) : 0.0;
            solve_ade(L, nMol, dt, q_upstream, ck, dk, mol, cur);
            if (fabs(cur[nMol - 2]) > 0.0) {
                double volTmp = fmax(0.0, h->REACH_VOL1[M][j]);
                double qoutTmp = cur[nMol - 2] * dt;
                double pcntReduc = fmin((volTmp + dt * q_upstream) * 0.999 / qoutTmp, 1.0);
                for (i = 1; i < nMol; i++) cur[i] = cur[i] * pcntReduc;
            }
            h->REACH_VOL1[M][j] = h->REACH_VOL1[M][j] + (q_upstream - cur[nMol - 2]) * dt;
            euler_stage(h, M, j);
            h->REACH_Q[M][j] = cur[nMol - 2] + Qlat;
            for (i = 0; i < nMol; i++) mol[i] = cur[i];
        } else {
            h->REACH_Q[M][j] = q_upstream + Qlat;
            for (i = 0; i < nMol; i++) mol[i] = 0.0;
            mol[nMol - 1] = h->REACH_Q[M][j];
            h->REACH_VOL0[M][j] = 0.0; h->REACH_VOL1[M][j] = 0.0; h->FLOOD_VOL1[M][j] = 0.0; h->REACH_ELE[M][j] = 0.0;
        }
    } else {
        h->REACH_Q[M][j] = Qlat;
        h->REACH_VOL0[M][j] = 0.0; h->REACH_VOL1[M][j] = 0.0; h->FLOOD_VOL1[M][j] = 0.0; h->REACH_ELE[M][j] = 0.0;
        for (i = 0; i < nMol; i++) mol[i] = 0.0;
        mol[nMol - 1] = h->REACH_Q[M][j];
    }
    return finish_reach(h, M, j, q_in, Qlat);
}

/* mc_route.f90:45-418: Muskingum-Cunge with sub-stepping when the Courant number exceeds one */
static int mc_rch(mro_t *h, int j)
{
    const int M = M_MC;
    const double Y = 0.5, Qmin = 1.e-50;
    double q_upstream, q_in, Qlat, dt = h->dt, *mol = h->MOL[M] + (size_t)j * 2;
    const int isHW = euler_inflow(h, M, j, &q_in, &Qlat);
    const double S = h->R_SLOPE[j], n = h->R_MAN_N[j], bt = h->R_WIDTH[j], bd = h->R_DEPTH[j], zc = h->SIDE_SLOPE[j], zf = h->FLDP_SLOPE[j], L = h->RLENGTH[j];
    double Q00 = mol[0], Q01 = mol[1], Q10, Q11;
    wm_cascade(h, M, j, q_in, &q_upstream, &Qlat);
    if (!isHW || h->hw_drain_point == 1) {
        if (L > h->min_length_route) {
            double theta = dt / L, Qbar;
            Q10 = q_upstream;
            Qbar = (Q00 + Q10 + Q01) / 3.0;
            if (Qbar > Qmin) {
                double depth = hy_flow_depth(fabs(Qbar), bt, zc, S, n, zf, bd);
                double ck = hy_celerity(fabs(Qbar), depth, bt, zc, S, n, zf, bd);
                double Cn = ck * theta, dTsub = dt, QinPrev, QoutPrev, sum = 0.0;
                int ntSub = 1, ix;
                if (Cn > 1.0) { ntSub = (int)ceil(dt / L * ck); dTsub = dt / ntSub; }
                QinPrev = Q00; QoutPrev = Q01;
                for (ix = 1; ix <= ntSub; ix++) {
                    double Qin = Q10, Qout;
                    Qbar = (Qin + QinPrev + QoutPrev) / 3.0;
                    if (Qbar > Qmin) {
                        double topWidth, X, C0, C1, C2;
                        depth = hy_flow_depth(fabs(Qbar), bt, zc, S, n, zf, bd);
                        topWidth = hy_Btop(depth, bt, zc, zf, bd);
                        ck = hy_celerity(fabs(Qbar), depth, bt, zc, S, n, zf, bd);
                        X = 0.5 * (1.0 - Qbar / (topWidth * S * ck * L));
                        Cn = ck * dTsub / L;
                        C0 = (-X + Cn * (1 - Y)) / (1 - X + Cn * (1 - Y));
                        C1 = (X + Cn * Y) / (1 - X + Cn * (1 - Y));
                        C2 = (1 - X - Cn * Y) / (1 - X + Cn * (1 - Y));
                        Qout = C0 * Qin + C1 * QinPrev + C2 * QoutPrev;
                        Qout = fmax(0.0, Qout);
                    } else Qout = 0.0;
                    sum = sum + Qout;
                    QinPrev = Qin; QoutPrev = Qout;
                }
                Q11 = sum / (double)ntSub;
                if (fabs(Q11) > 0.0) {
                    /* "*0.999" is a single-precision literal here (mc_route.f90:352), unlike kwe/dfw_route */
                    double pcntReduc = fmin((h->REACH_VOL1[M][j] / dt + Q10) * (double)0.999f / Q11, 1.0);
                    Q11 = Q11 * pcntReduc;
                }
                h->REACH_VOL1[M][j] = h->REACH_VOL1[M][j] + (Q10 - Q11) * dt;
                euler_stage(h, M, j);
                h->REACH_Q[M][j] = Q11 + Qlat;
            } else {
                Q11 = 0.0;
                h->REACH_Q[M][j] = Q11 + Qlat;
                h->REACH_VOL1[M][j] = h->REACH_VOL1[M][j] + (Q10 - Q11) * dt;
                euler_stage(h, M, j);
            }
        } else {
            Q10 = q_upstream; Q11 = q_upstream;
            h->REACH_Q[M][j] = q_upstream + Qlat;
            h->REACH_VOL0[M][j] = 0.0; h->REACH_VOL1[M][j] = 0.0; h->FLOOD_VOL1[M][j] = 0.0; h->REACH_ELE[M][j] = 0.0;
        }
    } else {
        Q10 = 0.0; Q11 = 0.0;
        h->REACH_Q[M][j] = Qlat;
        h->REACH_VOL0[M][j] = 0.0; h->REACH_VOL1[M][j] = 0.0; h->FLOOD_VOL1[M][j] = 0.0; h->REACH_ELE[M][j] = 0.0;
    }
    mol[0] = Q10; mol[1] = Q11;
    return finish_reach(h, M, j, q_in, Qlat);
}

/* ------------------------------------------------------------------------------------------ */
/* datetime_data.f90: month, day and day-of-year of simDatetime(1) = start + (iTime-1)*dt      */
/* ------------------------------------------------------------------------------------------ */
static long long cal_days_from_civil(long long y, int m, int d, int noleap)
{
    static const int cum[12] = {0, 31, 59, 90, 120, 151, 181, 212, 243, 273, 304, 334};
    long long era; unsigned yoe, doy, doe;
    if (noleap) return (y - 1970) * 365 + cum[m - 1] + (d - 1);
    y -= m <= 2;
    era = (y >= 0 ? y : y - 399) / 400;
    yoe = (unsigned)(y - era * 400); doy = (153u * (unsigned)(m + (m > 2 ? -3 : 9)) + 2) / 5 + (unsigned)d - 1; doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
    return era * 146097 + (long long)doe - 719468;
}
static void cal_civil_from_days(long long days, int noleap, int *y, int *m, int *d)
{
    if (noleap) {
        static const int ml[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
        long long yy = days >= 0 ? days / 365 : -((-days + 364) / 365); int doy = (int)(days - yy * 365), k = 0;
        while (doy >= ml[k]) doy -= ml[k++];
        *y = 1970 + (int)yy; *m = k + 1; *d = doy + 1;
    } else {
        long long z = days + 719468, era = (z >= 0 ? z : z - 146096) / 146097;
        unsigned doe = (unsigned)(z - era * 146097), yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
        unsigned doy = doe - (365 * yoe + yoe / 4 - yoe / 100), mp = (5 * doy + 2) / 153;
        *d = (int)(doy - (153 * mp + 2) / 5 + 1); *m = (int)(mp < 10 ? mp + 3 : mp - 9); *y = (int)(yoe + era * 400 + (*m <= 2));
    }
}
/* calendar of the step being routed; returns 0 when no start datetime was given */
static int step_calendar(const mro_t *h, int *month, int *day, int *doy)
{
    double t; long long days; int y;
    if (!h->hasStart) return 0;
    t = (double)cal_days_from_civil(h->startYear, h->startMonth, h->startDay, h->noleap) * SECPRDAY + h->startSec + (double)(h->iTime - 1) * h->dt;
    days = (long long)floor((t + 1.e-6) / SECPRDAY);
    cal_civil_from_days(days, h->noleap, &y, month, day);
    *doy = (int)(days - cal_days_from_civil(y, 1, 1, h->noleap)) + 1;        /* fn_dayofyear, datetime_data.f90:209-218 */
    return 1;
}
void mro_set_sim_start(mro_t *h, int year, int month, int day, double sec, int noleap)
{
    h->hasStart = 1; h->startYear = year; h->startMonth = month; h->startDay = day; h->startSec = sec; h->noleap = noleap;
}
int mro_set_lake_param(mro_t *h, const char *name, const double *values)
{
    int k;
    for (k = 0; LP_NAMES[k]; k++) if (!strcmp(LP_NAMES[k], name)) {
        free(h->LP[k]); h->LP[k] = (double *)malloc(sizeof(double) * (size_t)(h->nRch > 0 ? h->nRch : 1));
        memcpy(h->LP[k], values, sizeof(double) * (size_t)h->nRch);
        return 0;
    }
    return 1;
}

/* lake_route.f90:28-229,466-470 (endorheic and Doll03; LakeTargVol / WM / H06 / HYPE not restated) */
static int lake_route(mro_t *h, int j, int M)
{
    int m; double q_upstream = 0.0, dt = h->dt, *V1 = &h->REACH_VOL1[M][j], *Q = &h->REACH_Q[M][j];
    for (m = h->up_ptr[j]; m < h->up_ptr[j + 1]; m++) q_upstream = q_upstream + h->REACH_Q[M][h->up_idx[m]];
    const int targVol = h->LP[LP_LakeTargVol] && h->LP[LP_LakeTargVol][j] != 0.0;
    const double wmVolJ = h->wmVol ? h->wmVol[j] : 0.0;        /* REACH_WM_VOL (0 when is_vol_wm is off, main_route.f90:117-123) */
    if (h->iTime == 1 && h->volJumpStart && targVol) {          /* lake_route.f90:137-139 */
        *V1 = wmVolJ;
    } else if (h->iTime == 1) {   /* cold start (isColdStart = T) */
        switch (h->lakeModelType[j]) {
            case LAKE_ENDORHEIC: *V1 = h->D03_S0[j]; break;
            case LAKE_DOLL03:    *V1 = h->D03_MaxStorage[j]; break;
            case LAKE_H06:
                if (!h->LP[LP_H06_Smax]) { snprintf(h->message, 256, "lake_route/Hanasaki parameters are not set"); return 20; }
                *V1 = h->LP[LP_H06_Smax][j]; break;
            case LAKE_HYPE:
                if (!h->LP[LP_HYP_E_emr] || !h->LP[LP_HYP_E_zero] || !h->LP[LP_HYP_A_avg]) { snprintf(h->message, 256, "lake_route/HYPE parameters are not set"); return 20; }
                *V1 = (h->LP[LP_HYP_E_emr][j] - h->LP[LP_HYP_E_zero][j]) * h->LP[LP_HYP_A_avg][j]; break;
            default: snprintf(h->message, 256, "lake_route/lake model type not restated in oracle"); return 20;
        }
    }
    h->REACH_VOL0[M][j] = *V1;
    *V1 = *V1 + q_upstream * dt;
    if (h->LakeInputOption == 1 || h->LakeInputOption == 2) *V1 = *V1 + h->BASIN_QR1[j] * dt;
    if (h->LakeInputOption == 0 || h->LakeInputOption == 2) {   /* lake_route.f90:166-174 */
        const double pr = h->hasEP ? h->reachPrecip[j] : 0.0, ev = h->hasEP ? h->reachEvapo[j] : 0.0;
        *V1 = *V1 + pr * dt;
        if (*V1 > ev * dt) *V1 = *V1 - ev * dt;
        else { if (h->hasEP) h->reachEvapo[j] = *V1 / dt; *V1 = 0.0; }   /* basinevapo is shared by the routing methods of a step */
    }
    /* water abstraction / injection from the lake (lake_route.f90:176-193) */
    h->WM_ACTUAL[M][j] = h->wmFlux ? h->wmFlux[j] : 0.0;
    if (h->wmFlux && h->wmFlux[j] != -9999.0) {
        const double f = h->wmFlux[j];
        if (f <= 0) { *V1 = *V1 - f * dt; h->WM_ACTUAL[M][j] = f; }
        else if (f * dt <= *V1) { *V1 = *V1 - f * dt; h->WM_ACTUAL[M][j] = f; }
        else { h->WM_ACTUAL[M][j] = *V1 / dt; *V1 = 0.0; }
    }
    if (targVol) {                                             /* lake_route.f90:196-203 */
        if (*V1 < wmVolJ) *Q = 0;
        else { *Q = (*V1 - wmVolJ) / dt; *V1 = wmVolJ; }
    } else
    switch (h->lakeModelType[j]) {
        case LAKE_ENDORHEIC: *Q = 0.0; break;
        case LAKE_DOLL03:
            if ((*V1 - h->D03_S0[j]) > 0)
                *Q = h->D03_Coefficient[j] * (*V1 - h->D03_S0[j]) *
                     pow((*V1 - h->D03_S0[j]) / (h->D03_MaxStorage[j] - h->D03_S0[j]), h->D03_Power[j]);
            else *Q = 0;
            *Q = *Q / SECPRDAY;
            *Q = fmin(*Q, *V1 / dt);
            *V1 = *V1 - *Q * dt;
            break;
        case LAKE_H06: {           /* lake_route.f90:231-396 (no water-management demand: is_flux_wm = F) */
            int k, i, month, day, doy, start_month = 0;
            double I_months[12], D_months[12], I_yearly, D_yearly, c, target_r, sI = 0.0, sD = 0.0, ratio;
            double **P = h->LP;
            for (k = LP_H06_Smax; k < LP_END; k++) if (!P[k]) { snprintf(h->message, 256, "lake_route/Hanasaki parameter %s is not set", LP_NAMES[k]); return 20; }
            if (!step_calendar(h, &month, &day, &doy)) { snprintf(h->message, 256, "lake_route/Hanasaki needs the simulation start datetime"); return 20; }
            if (P[LP_H06_I_mem_F][j] != 0.0) {      /* memory of the upstream inflow, one row per month */
                static const int ndays31[7] = {0, 2, 4, 6, 7, 9, 11}, ndays30[3] = {3, 5, 8};
                const double memL = (double)(int)P[LP_H06_I_mem_L][j];
                int L31 = (int)floor(memL * 31 * SECPRDAY / dt), L30 = (int)floor(memL * 30 * SECPRDAY / dt), LF, L, m2;
                double *mem;
                if (!h->h06Mem) { h->h06Mem = (double **)calloc((size_t)h->nRch, sizeof(double *)); h->h06Len = (int *)calloc((size_t)h->nRch, sizeof(int)); }
                if (!h->h06Mem[j]) {                /* first call: filled with the monthly parameters, nothing inserted */
                    h->h06Len[j] = L31;
                    h->h06Mem[j] = (double *)malloc(sizeof(double) * 12 * (size_t)(L31 > 0 ? L31 : 1));
                    for (k = 0; k < 12; k++) for (i = 0; i < L31; i++) h->h06Mem[j][(size_t)k * L31 + i] = P[LP_H06_I_Jan + k][j];
                } else {                            /* shift the current month's row and insert the inflow of this call */
                    L = h->h06Len[j]; mem = h->h06Mem[j] + (size_t)(month - 1) * L;
                    for (i = L - 1; i >= 1; i--) mem[i] = mem[i - 1];
                    mem[0] = q_upstream;
                }
                L = h->h06Len[j];
                for (k = 0; k < 7; k++) { m2 = ndays31[k]; mem = h->h06Mem[j] + (size_t)m2 * L; sI = 0.0; for (i = 0; i < L31; i++) sI += mem[i]; P[LP_H06_I_Jan + m2][j] = sI / L31; }
                for (k = 0; k < 3; k++) { m2 = ndays30[k]; mem = h->h06Mem[j] + (size_t)m2 * L; sI = 0.0; for (i = 0; i < L30; i++) sI += mem[i]; P[LP_H06_I_Jan + m2][j] = sI / L30; }
                /* November is not updated by the reference (lake_route.f90:258-272) */
                LF = h->noleap ? (int)floor(memL * 28 * SECPRDAY / dt) : (int)floor(memL * 28.25 * SECPRDAY / dt);
                mem = h->h06Mem[j] + (size_t)1 * L; sI = 0.0; for (i = 0; i < LF; i++) sI += mem[i]; P[LP_H06_I_Jan + 1][j] = sI / LF;
            }
            sI = 0.0; sD = 0.0;
            for (k = 0; k < 12; k++) { I_months[k] = P[LP_H06_I_Jan + k][j]; D_months[k] = P[LP_H06_D_Jan + k][j]; sI += I_months[k]; sD += D_months[k]; }
            I_yearly = sI / 12; D_yearly = sD / 12;
            c = P[LP_H06_Smax][j] / (I_yearly * 365 * SECPRDAY);
            for (i = 1; i <= 12; i++) if (I_yearly <= I_months[i - 1]) start_month = i + 1;
            if (month == start_month && day == 1) P[LP_H06_E_rel_ini][j] = *V1 / (P[LP_H06_alpha][j] * P[LP_H06_Smax][j]);
            if ((int)P[LP_H06_purpose][j] == 1) {
                if (P[LP_H06_envfact][j] * I_yearly <= D_yearly)
                    target_r = I_months[month - 1] * P[LP_H06_c1][j] + I_yearly * P[LP_H06_c2][j] * (D_months[month - 1] / D_yearly);
                else target_r = I_yearly + D_months[month - 1] - D_yearly;
            } else target_r = I_yearly;
            if (c >= P[LP_H06_c_compare][j]) *Q = target_r * P[LP_H06_E_rel_ini][j];
            else if (0 <= c && c < P[LP_H06_c_compare][j]) {
                ratio = pow(c / P[LP_H06_denominator][j], P[LP_H06_exponent][j]);
                *Q = P[LP_H06_E_rel_ini][j] * target_r * ratio + q_upstream * (1 - pow(c / P[LP_H06_denominator][j], P[LP_H06_exponent][j]));
            }                                       /* else (c < 0 or NaN): REACH_Q keeps its previous value */
            if (*V1 < (P[LP_H06_Smax][j] * P[LP_H06_frac_Sdead][j])) {
                *Q = *Q - (P[LP_H06_Smax][j] * P[LP_H06_frac_Sdead][j] - *V1) / dt;
                if (*Q < 0) *Q = 0;
            } else if (*V1 > P[LP_H06_Smax][j]) {
                *Q = *Q + (*V1 - P[LP_H06_Smax][j]) / dt;
            }
            *V1 = *V1 - *Q * dt;
            break; }
        case LAKE_HYPE: {          /* lake_route.f90:398-438 */
            int k, month, day, doy;
            double ELE, F_sin, F_lin, Q_prim, Q_spill, Q_sim; int F_prim;
            for (k = 0; k < LP_COUNT; k++) if (!h->LP[k]) { snprintf(h->message, 256, "lake_route/HYPE parameter %s is not set", LP_NAMES[k]); return 20; }
            if (!step_calendar(h, &month, &day, &doy)) { snprintf(h->message, 256, "lake_route/HYPE needs the simulation start datetime"); return 20; }
            ELE = *V1 / h->LP[LP_HYP_A_avg][j] + h->LP[LP_HYP_E_zero][j];
            F_sin = fmax(0.0, (1 + h->LP[LP_HYP_Qrate_amp][j] * sin(2 * PI_MR * (doy + (int)h->LP[LP_HYP_Qrate_phs][j]) / 365)));
            F_lin = fmin(fmax((ELE - h->LP[LP_HYP_E_min][j]) / (h->LP[LP_HYP_E_lim][j] - h->LP[LP_HYP_E_min][j]), 0.0), 1.0);
            F_prim = h->LP[LP_HYP_prim_F][j] != 0.0 ? 1 : 0;
            Q_prim = F_sin * F_lin * F_prim * h->LP[LP_HYP_Qrate_prim][j];
            Q_spill = 0.0;
            if (ELE > h->LP[LP_HYP_E_emr][j]) Q_spill = h->LP[LP_HYP_Qrate_emr][j] * pow(ELE - h->LP[LP_HYP_E_emr][j], h->LP[LP_HYP_Erate_emr][j]);
            if (h->LP[LP_HYP_Qsim_mode][j] != 0.0) Q_sim = Q_prim + Q_spill; else Q_sim = fmax(Q_prim, Q_spill);
            *Q = fmin(Q_sim, fmax(0.0, (ELE - h->LP[LP_HYP_E_min][j]) * h->LP[LP_HYP_A_avg][j]) / dt);
            *V1 = *V1 - *Q * dt;
            break; }
        default: snprintf(h->message, 256, "lake_route/lake model type not restated in oracle"); return 20;
    }
    /* lake_route does not touch REACH_INFLOW */
    comp_reach_wb_lake(h, M, j, q_upstream, h->BASIN_QR1[j], 1);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* kwt_route.f90                                                                               */
/* ------------------------------------------------------------------------------------------ */
typedef struct { double *Q, *T, *X; int n, cap; } wave_buf_t;   /* Q_JRCH, TENTRY, T_EXIT (0:n-1) */

static void wb_reserve(wave_buf_t *b, int n)
{
    if (n > b->cap) {
        int c = b->cap ? b->cap : 64; while (c < n) c *= 2;
        b->Q = (double *)realloc(b->Q, sizeof(double) * c); b->T = (double *)realloc(b->T, sizeof(double) * c);
        b->X = (double *)realloc(b->X, sizeof(double) * c); b->cap = c;
    }
}

/* kwt_route.f90:1444-1622, one output interval [T0,T1]; TOLD/QOLD are 1-based (NOLD entries) */
static int interp_rch(const double *TOLD1, const double *QOLD1, int NOLD, double T0, double T1, double *QNEW)
{
    const double *TOLD = TOLD1 - 1, *QOLD = QOLD1 - 1;   /* 1-based views */
    int IBEG, IEND, IMID, i; double AREAB = 0.0, AREAE = 0.0, AREAM = 0.0, SLOPE, QEST0, QEST1;
    if (TOLD[1] > T0 || TOLD[NOLD] < T1) return 1;
    IBEG = 1;
    for (i = 2; i <= NOLD; i++) if (T0 <= TOLD[i]) { IBEG = i; break; }
    IEND = 1;
    for (i = 1; i <= NOLD; i++) if (T1 <= TOLD[i]) { IEND = i; break; }
    if (T1 < TOLD[IBEG]) {
        SLOPE = (QOLD[IBEG] - QOLD[IBEG - 1]) / (TOLD[IBEG] - TOLD[IBEG - 1]);
        QEST0 = SLOPE * (T0 - TOLD[IBEG - 1]) + QOLD[IBEG - 1];
        QEST1 = SLOPE * (T1 - TOLD[IBEG - 1]) + QOLD[IBEG - 1];
        *QNEW = 0.5 * (QEST0 + QEST1);
        return 0;
    }
    if (T0 < TOLD[IBEG]) {
        SLOPE = (QOLD[IBEG] - QOLD[IBEG - 1]) / (TOLD[IBEG] - TOLD[IBEG - 1]);
        QEST0 = SLOPE * (T0 - TOLD[IBEG - 1]) + QOLD[IBEG - 1];
        AREAB = (TOLD[IBEG] - T0) * 0.5 * (QEST0 + QOLD[IBEG]);
    }
    if (T1 < TOLD[IEND]) {
        SLOPE = (QOLD[IEND] - QOLD[IEND - 1]) / (TOLD[IEND] - TOLD[IEND - 1]);
        QEST1 = SLOPE * (T1 - TOLD[IEND - 1]) + QOLD[IEND - 1];
        AREAE = (T1 - TOLD[IEND - 1]) * 0.5 * (QOLD[IEND - 1] + QEST1);
    }
    if (IBEG < IEND) {
        for (IMID = IBEG + 1; IMID <= IEND; IMID++) {
            if (IMID < IEND || (IMID == IEND && T1 == TOLD[IEND] && T0 < TOLD[IEND - 1]))
                AREAM = AREAM + (TOLD[IMID] - TOLD[IMID - 1]) * 0.5 * (QOLD[IMID - 1] + QOLD[IMID]);
        }
    }
    *QNEW = (AREAB + AREAE + AREAM) / (T1 - T0);
    return 0;
}

/* kwt_route.f90:999-1123 */
static double rm_interp(double T0, double Q1, double Q2, double T1, double T2)
{
    return Q1 + ((Q2 - Q1) / (T2 - T1)) * (T0 - T1);
}
static int remove_rch(wave_buf_t *b)
{
    int NPRT = b->n - 1, IPRT, MPRT, i, k, ISEL;
    double *Q, *T, *Z, *ABSERR; unsigned char *PARFLG; int *INDEX1;
#pragma omp atomic
    g_cnt[0]++;
    if (b->n > g_cnt[4]) g_cnt[4] = b->n;
    Q = (double *)malloc(sizeof(double) * b->n); T = (double *)malloc(sizeof(double) * b->n);
    Z = (double *)malloc(sizeof(double) * b->n); ABSERR = (double *)malloc(sizeof(double) * b->n);
    PARFLG = (unsigned char *)malloc(b->n);
    INDEX1 = (int *)malloc(sizeof(int) * b->n);
    memcpy(Q, b->Q, sizeof(double) * b->n); memcpy(T, b->T, sizeof(double) * b->n); memcpy(Z, b->X, sizeof(double) * b->n);
    for (i = 0; i <= NPRT; i++) { PARFLG[i] = 1; ABSERR[i] = HUGE_DP; }
    for (IPRT = 1; IPRT <= NPRT - 1; IPRT++)
        ABSERR[IPRT] = fabs(rm_interp(T[IPRT], Q[IPRT - 1], Q[IPRT + 1], T[IPRT - 1], T[IPRT + 1]) - Q[IPRT]);
    for (;;) {
        double emin;
        MPRT = -1;
        for (i = 0; i <= NPRT; i++) if (PARFLG[i]) INDEX1[++MPRT] = i;   /* INDEX1(0:MPRT) = pack(INDEX0,PARFLG) */
        if (MPRT < MAXQPAR) break;
        ISEL = 0; emin = ABSERR[INDEX1[0]];
        for (k = 1; k <= MPRT; k++) if (ABSERR[INDEX1[k]] < emin) { emin = ABSERR[INDEX1[k]]; ISEL = k; }   /* minloc: first minimum */
        if (INDEX1[ISEL - 1] > 0) {
            int INEG = INDEX1[ISEL - 2], IMID = INDEX1[ISEL - 1], IPOS = INDEX1[ISEL + 1];
            ABSERR[IMID] = fabs(rm_interp(T[IMID], Q[INEG], Q[IPOS], T[INEG], T[IPOS]) - Q[IMID]);
        }
        if (INDEX1[ISEL + 1] < NPRT) {
            int INEG = INDEX1[ISEL - 1], IMID = INDEX1[ISEL + 1], IPOS = INDEX1[ISEL + 2];
            ABSERR[IMID] = fabs(rm_interp(T[IMID], Q[INEG], Q[IPOS], T[INEG], T[IPOS]) - Q[IMID]);
        }
        PARFLG[INDEX1[ISEL]] = 0;
    }
    for (k = 0; k <= MPRT; k++) { b->Q[k] = Q[INDEX1[k]]; b->T[k] = T[INDEX1[k]]; b->X[k] = Z[INDEX1[k]]; }
    b->n = MPRT + 1;
    free(Q); free(T); free(Z); free(ABSERR); free(PARFLG); free(INDEX1);
    return 0;
}

/* kwt_route.f90:1130-1439.  Arrays are the 1-based slices Q_JRCH(1:NQ1) etc. */
static int kinwav_rch(mro_t *h, int JRCH, double T_START, double T_END,
                      double *Q_JRCH1, double *TENTRY1, double *T_EXIT1, unsigned char *FROUTE1, int NQ1, int *NQ2)
{
    double *Q_JRCH = Q_JRCH1 - 1, *TENTRY = TENTRY1 - 1, *T_EXIT = T_EXIT1 - 1; unsigned char *FROUTE = FROUTE1 - 1;
    enum { CAP = 64 };
    int IX[CAP], MF[CAP]; double T0[CAP], T1[CAP], Q0[CAP], Q1[CAP], Q2[CAP], WC[CAP];
    double ALFA, K, XMX, X, XB, WDIFF, XXB, A1, A2, CM, TEXIT, TNEXT = 0.0, TEXIT2;
    int NN, NI, IW, JW, IXB = 0, JXB, IROUTE, JROUTE, ICOUNT, i;
    double p1, p2;

    *NQ2 = 0;
    if (NQ1 + 2 > CAP) { snprintf(h->message, 256, "kinwav_rch/oracle scratch too small"); return 60; }
    ALFA = 5.0 / 3.0;
    K = sqrt(h->R_SLOPE[JRCH]) / h->R_MAN_N[JRCH];
    XMX = h->RLENGTH[JRCH];
    NN = NQ1; NI = NN;
    if (NN == 0) return 0;
    for (i = 1; i <= NI; i++) { MF[i] = i; IX[i] = i; Q0[i] = Q1[i] = Q2[i] = Q_JRCH[i]; T0[i] = T1[i] = TENTRY[i]; }
    p1 = 1.0 / ALFA; p2 = (ALFA - 1.0) / ALFA;
    for (i = 1; i <= NN; i++) WC[i] = ALFA * pow(K, p1) * pow(Q1[i], p2);

    if (NN > 1) {
        X = 0.0;
        for (;;) {
            XB = XMX;
            for (IW = 2; IW <= NN; IW++) {
                JW = IW - 1;
                if (WC[IW] == 0.0 || WC[JW] == 0.0) continue;
                WDIFF = 1.0 / WC[JW] - 1.0 / WC[IW];
                if (WDIFF == 0.0) continue;
                if (WC[IW] == WC[JW]) continue;
                XXB = (T1[IW] - T1[JW]) / WDIFF;
                if (XXB < X || XXB > XB) continue;
                XB = XXB; IXB = IW;
            }
            if (XB == XMX) break;
#pragma omp atomic
            g_cnt[1]++;
            NN = NN - 1;
            JXB = IXB - 1;
            Q2[JXB] = fmax(Q2[JXB], Q2[IXB]);
            Q1[JXB] = fmin(Q1[JXB], Q1[IXB]);
            A2 = pow(Q2[JXB] / K, p1);
            A1 = pow(Q1[JXB] / K, p1);
            CM = (Q2[JXB] - Q1[JXB]) / (A2 - A1);
            T1[JXB] = T1[JXB] + XB / WC[JXB] - XB / CM;
            WC[JXB] = CM;
            for (i = IX[IXB]; i <= NI; i++) MF[i] = MF[i] - 1;
            for (i = IXB; i <= NN; i++) { IX[i] = IX[i + 1]; T1[i] = T1[i + 1]; WC[i] = WC[i + 1]; Q1[i] = Q1[i + 1]; Q2[i] = Q2[i + 1]; }
            X = XB;
        }
    }

    ICOUNT = 0;
#define RUPDATE(QNEW_, TOLD_, TNEW_) do { \
        ICOUNT = ICOUNT + 1; \
        if (ICOUNT > NQ1) { snprintf(h->message, 256, "kinwav_rch/RUPDATE/array bounds exceeded"); return 60; } \
        Q_JRCH[ICOUNT] = (QNEW_); TENTRY[ICOUNT] = (TOLD_); T_EXIT[ICOUNT] = (TNEW_); \
        if (ICOUNT > 1) { if (T_EXIT[ICOUNT] <= T_EXIT[ICOUNT - 1]) T_EXIT[ICOUNT] = T_EXIT[ICOUNT - 1] + 1.0; } \
        if (ICOUNT == 1 && T_EXIT[ICOUNT] <= T_START) T_EXIT[ICOUNT] = T_START + 1.0; \
        if (T_EXIT[ICOUNT] < T_END) FROUTE[ICOUNT] = 1; \
    } while (0)

    for (IROUTE = 1; IROUTE <= NN; IROUTE++) {
        if (WC[IROUTE] < VERYSMALL) { snprintf(h->message, 256, "kinwav_rch/zero flow for reach id %d", h->segId[JRCH]); return 20; }
        TEXIT = fmin(XMX / WC[IROUTE] + T1[IROUTE], HUGE_DP);
        if (IROUTE < NN) TNEXT = fmin(XMX / WC[IROUTE + 1] + T1[IROUTE + 1], HUGE_DP);
        if (IROUTE == NN) TNEXT = HUGE_DP;
        if (Q1[IROUTE] != Q2[IROUTE]) {
            if (TEXIT < T_END) {
                TEXIT2 = fmin(TEXIT + 1.0, TEXIT + 0.5 * (fmin(TNEXT, T_END) - TEXIT));
                if (TEXIT2 == TEXIT) { snprintf(h->message, 256, "kinwav_rch/TEXIT equals TEXIT2 in kinwav"); return 30; }
#pragma omp atomic
                g_cnt[2]++;
                RUPDATE(Q1[IROUTE], T1[IROUTE], TEXIT);
                RUPDATE(Q2[IROUTE], T1[IROUTE], TEXIT2);
            } else {
#pragma omp atomic
                g_cnt[3]++;
                for (JROUTE = 1; JROUTE <= NI; JROUTE++)
                    if (MF[JROUTE] == IROUTE) RUPDATE(Q0[JROUTE], T0[JROUTE], TEXIT);
            }
        } else {
            RUPDATE(Q1[IROUTE], T1[IROUTE], TEXIT);
        }
    }
#undef RUPDATE
    *NQ2 = ICOUNT;
    return 0;
}

/* kwt_route.f90:619-993 */
typedef struct { const double *QF, *TR; const unsigned char *RF; int n; double basQF[2], basTR[2]; unsigned char basRF[2]; } series_t;

static int qexmul_rch(mro_t *h, int JRCH, double T0, double T1, wave_buf_t *out, int base /* write QD,TD at out[base..] */, int *ND)
{
    int NUPB, NUPR = 0, NUPS, IUPS, INDX, MUPR, IR, IUPR, IMAX, IPRT, JUPS, JUPS_OLD, ITIM_OLD, IWAV, IBEG, IEND, nDone, i;
    double DT = T1 - T0, TIME_OLD, Q_AGG, SCFAC, SFLOW, SLOPE, PREDV;
    series_t *US; double *UWIDTH, *CTIME; int *ITIM; unsigned char *MFLG;
    kwave_t *snap;   /* copies of the upstream waves taken before they are stripped (USFLOW) */

    *ND = 0;
    if (h->nGood[JRCH] == 0) return 0;
    NUPB = h->up_ptr[JRCH + 1] - h->up_ptr[JRCH];
    for (IUPS = 0; IUPS < NUPB; IUPS++) { INDX = h->up_idx[h->up_ptr[JRCH] + IUPS]; MUPR = h->nGood[INDX]; if (MUPR > 0) NUPR++; }
    NUPS = NUPB + NUPR;

    if (NUPS == 1) {   /* one upstream basin that is a headwater */
        IR = h->up_idx[h->up_ptr[JRCH]];
        wb_reserve(out, base + 1);
        out->Q[base] = h->BASIN_QR1[IR] / h->R_WIDTH[JRCH];
        out->T[base] = T1;
        *ND = 1;
        return 0;
    }

    US = (series_t *)calloc(NUPS, sizeof(series_t)); UWIDTH = (double *)malloc(sizeof(double) * NUPS);
    CTIME = (double *)malloc(sizeof(double) * NUPS); ITIM = (int *)malloc(sizeof(int) * NUPS);
    MFLG = (unsigned char *)calloc(NUPS, 1); snap = (kwave_t *)malloc(sizeof(kwave_t) * (NUPR > 0 ? NUPR : 1));
    IMAX = NUPB;
    for (IUPS = 0; IUPS < NUPB; IUPS++) {
        series_t *s = &US[IUPS];
        IR = h->up_idx[h->up_ptr[JRCH] + IUPS];
        s->basQF[0] = h->BASIN_QR0[IR]; s->basQF[1] = h->BASIN_QR1[IR];
        s->basTR[0] = T0; s->basTR[1] = T1; s->basRF[0] = s->basRF[1] = 1;
        s->QF = s->basQF; s->TR = s->basTR; s->RF = s->basRF; s->n = 2;
        UWIDTH[IUPS] = 1.0;
        CTIME[IUPS] = s->TR[1];
    }
    IUPR = 0;
    for (IUPS = 0; IUPS < NUPB; IUPS++) {
        INDX = h->up_idx[h->up_ptr[JRCH] + IUPS];
        MUPR = h->nGood[INDX];
        if (MUPR > 0) {
            kwave_t *w = &h->KW[INDX]; int NS, NR = 0, NQ; series_t *s;
            IUPR = IUPR + 1;
            IR = INDX;
            NS = w->n;
            if (NS == 0) { snprintf(h->message, 256, "qexmul_rch/RCHSTA_out%%LKW_ROUTE%%KWAVE is not associated"); free(US); free(UWIDTH); free(CTIME); free(ITIM); free(MFLG); free(snap); return 20; }
            for (i = 0; i < NS; i++) NR += w->RF[i];
            NQ = (NR + 1 < NS) ? NR + 1 : NS;
            snap[IUPR - 1] = *w;
            s = &US[NUPB + IUPR - 1];
            s->QF = snap[IUPR - 1].QF; s->TR = snap[IUPR - 1].TR; s->RF = snap[IUPR - 1].RF; s->n = NQ;
            /* remove the routed particles from the upstream reach: KWAVE(0:NS-NR) = NEW_WAVE(NR-1:NS-1) */
            {
                int nn = NS - NR + 1, k;
                if (NR - 1 < 0) { snprintf(h->message, 256, "qexmul_rch/upstream wave has no routed element"); free(US); free(UWIDTH); free(CTIME); free(ITIM); free(MFLG); free(snap); return 20; }
                for (k = 0; k < nn; k++) { w->QF[k] = snap[IUPR - 1].QF[NR - 1 + k]; w->TI[k] = snap[IUPR - 1].TI[NR - 1 + k]; w->TR[k] = snap[IUPR - 1].TR[NR - 1 + k]; w->RF[k] = snap[IUPR - 1].RF[NR - 1 + k]; }
                w->n = nn;
            }
            UWIDTH[NUPB + IUPR - 1] = h->R_WIDTH[IR];
            CTIME[NUPB + IUPR - 1] = s->TR[1];
            IMAX = IMAX + (NR - 1);
        }
    }

    wb_reserve(out, base + IMAX + 1);
    IPRT = 0;
    for (i = 0; i < NUPS; i++) ITIM[i] = 1;
    JUPS_OLD = 0x7fffffff; ITIM_OLD = 0x7fffffff;
    for (;;) {
        JUPS = 0; for (i = 1; i < NUPS; i++) if (CTIME[i] < CTIME[JUPS]) JUPS = i;   /* MINLOC: first minimum */
        if (JUPS == JUPS_OLD && ITIM[JUPS] == ITIM_OLD) { snprintf(h->message, 256, "qexmul_rch/stuck in the continuous do-loop"); free(US); free(UWIDTH); free(CTIME); free(ITIM); free(MFLG); free(snap); return 20; }
        JUPS_OLD = JUPS; ITIM_OLD = ITIM[JUPS];
        if (!MFLG[JUPS]) {
            if (!US[JUPS].RF[ITIM[JUPS]]) {
                MFLG[JUPS] = 1; CTIME[JUPS] = HUGE_DP;
            } else {
                if (IPRT >= 1) TIME_OLD = out->T[base + IPRT - 1]; else TIME_OLD = -HUGE_DP;
                if (CTIME[JUPS] < TIME_OLD) { snprintf(h->message, 256, "qexmul_rch/expect process in order of time"); free(US); free(UWIDTH); free(CTIME); free(ITIM); free(MFLG); free(snap); return 30; }
                if (CTIME[JUPS] != TIME_OLD) {
                    Q_AGG = 0.0;
                    for (IUPS = 0; IUPS < NUPS; IUPS++) {
                        const series_t *s = &US[IUPS];
                        IWAV = ITIM[IUPS];
                        SCFAC = UWIDTH[IUPS] / h->R_WIDTH[JRCH];
                        if (IUPS == JUPS) {
                            SFLOW = s->QF[IWAV] * SCFAC;
                        } else {
                            IBEG = IWAV; if (s->TR[IBEG] >= CTIME[JUPS]) IBEG = IWAV - 1;
                            IEND = IBEG + 1;
                            if (IBEG < 0 || IEND >= s->n) { snprintf(h->message, 256, "qexmul_rch/bracket beyond series"); free(US); free(UWIDTH); free(CTIME); free(ITIM); free(MFLG); free(snap); return 40; }
                            if (s->TR[IEND] < CTIME[JUPS] || s->TR[IBEG] > CTIME[JUPS]) { snprintf(h->message, 256, "qexmul_rch/the times are not ordered as we assume"); free(US); free(UWIDTH); free(CTIME); free(ITIM); free(MFLG); free(snap); return 40; }
                            SLOPE = (s->QF[IEND] - s->QF[IBEG]) / (s->TR[IEND] - s->TR[IBEG]);
                            PREDV = s->QF[IBEG] + SLOPE * (CTIME[JUPS] - s->TR[IBEG]);
                            SFLOW = PREDV * SCFAC;
                        }
                        Q_AGG = Q_AGG + SFLOW;
                    }
                    IPRT = IPRT + 1;
                    if (IPRT > IMAX) { snprintf(h->message, 256, "qexmul_rch/QD_TEMP bounds exceeded"); free(US); free(UWIDTH); free(CTIME); free(ITIM); free(MFLG); free(snap); return 60; }
                    out->Q[base + IPRT - 1] = Q_AGG;
                    out->T[base + IPRT - 1] = CTIME[JUPS];
                } else {
#pragma omp atomic
                    g_cnt[5]++;
                }
                if (ITIM[JUPS] == US[JUPS].n - 1) { MFLG[JUPS] = 1; CTIME[JUPS] = HUGE_DP; }
                else { ITIM[JUPS] = ITIM[JUPS] + 1; CTIME[JUPS] = US[JUPS].TR[ITIM[JUPS]]; }
            }
        }
        nDone = 0; for (i = 0; i < NUPS; i++) nDone += MFLG[i];
        if (nDone == NUPS) break;
    }
    free(US); free(UWIDTH); free(CTIME); free(ITIM); free(MFLG); free(snap);
    *ND = IPRT;
    (void)DT;
    return 0;
}

/* kwt_route.f90:461-613 */
static int getusq_rch(mro_t *h, int JRCH, double T0, double T1, wave_buf_t *b)
{
    double DT = T1 - T0; int ND = 0, NJ, ierr, i; kwave_t *w = &h->KW[JRCH];
    /* own entries go first, so reserve their place after we know the (pre-existing) size; a cold
       start has exactly one own entry that is defined from QD(1) */
    int nOwn = w->n > 0 ? w->n : 1;
    int isUpLake = 0, iUp = -1, nUps = h->up_ptr[JRCH + 1] - h->up_ptr[JRCH];
    wb_reserve(b, nOwn + 1);
    if (h->is_lake_sim) {
        for (i = 0; i < nUps; i++) { int u = h->up_idx[h->up_ptr[JRCH] + i]; if (h->isLake[u]) { isUpLake = 1; iUp = u; } }
        if (isUpLake && nUps > 1) { snprintf(h->message, 256, "getusq_rch/lake outlet reach should have one upstream lake"); return 10; }
    }
    if (isUpLake) {
        ND = 1;
        b->Q[nOwn] = h->REACH_Q[M_KWT][iUp] / h->R_WIDTH[JRCH];
        b->T[nOwn] = T1;
    } else {
        ierr = qexmul_rch(h, JRCH, T0, T1, b, nOwn, &ND);
        if (ierr) return ierr;
    }
    if (w->n == 0) {   /* cold start, kwt_route.f90:587-596 */
        w->n = 1; w->QF[0] = b->Q[nOwn]; w->TI[0] = T0 - DT; w->TR[0] = T0; w->RF[0] = 1;
    }
    NJ = w->n - 1;
    for (i = 0; i <= NJ; i++) { b->Q[i] = w->QF[i]; b->T[i] = w->TI[i]; b->X[i] = w->TR[i]; }
    for (i = 0; i < ND; i++) b->X[NJ + 1 + i] = -9999.0;
    b->n = NJ + 1 + ND;
    return 0;
}

/* kwt_route.f90:36-346 */
static int kwt_rch(mro_t *h, int j, double T0, double T1, wave_buf_t *b)
{
    const int M = M_KWT;
    int NUPS = h->nGood[j], ierr, i, k, NQ1, NQ2, NR, NN; kwave_t *w = &h->KW[j];
    double q_upstream, T_START, T_END, QNEW, Q_END, TIMEI; unsigned char FROUTE[64];
    if (NUPS > 0) {
        ierr = getusq_rch(h, j, T0, T1, b);
        if (ierr) return ierr;
        for (i = 0; i < b->n; i++) if (b->Q[i] < 0.0) { snprintf(h->message, 256, "kwt_rch/negative flow extracted from upstream reach"); return 20; }
        q_upstream = 0.0;
        for (k = 0; k < NUPS; k++) {
            int m = h->up_ptr[j] + k;
            if (!h->goodBas[m]) continue;
            q_upstream = q_upstream + h->REACH_Q[M][h->up_idx[m]];
        }
        h->REACH_INFLOW[M][j] = q_upstream;
    } else {
        h->REACH_INFLOW[M][j] = 0.0;
        h->REACH_Q[M][j] = h->BASIN_QR1[j];
        w->n = 1; w->QF[0] = -9999; w->TI[0] = -9999; w->TR[0] = -9999; w->RF[0] = 0;
        return 0;
    }
    if (b->n > MAXQPAR) { ierr = remove_rch(b); if (ierr) return ierr; }
    NQ1 = b->n - 1;
    T_START = T0; T_END = T1;     /* RSTEP = 0 */
    if (h->wmFlux && h->wmFlux[j] != -9999.0) {   /* extract_from_rch, kwt_route.f90:351-455 (note its sign: > 0 adds water) */
        const double Qtake = h->wmFlux[j], alfa = 5.0 / 3.0, K = sqrt(h->R_SLOPE[j]) / h->R_MAN_N[j];
        const int NRw = b->n;
        double Qavg, totQ, Qfrac;
        ierr = interp_rch(b->T, b->Q, NRw, T_START, T_END, &Qavg);
        if (ierr) { snprintf(h->message, 256, "extract_from_rch/interp_rch"); return ierr; }
        totQ = Qavg * h->R_WIDTH[j];
        if (Qtake > 0.0) { Qfrac = Qtake / totQ; for (i = 1; i < NRw; i++) b->Q[i] = b->Q[i] * (1.0 + Qfrac); }
        else if (Qtake < 0.0 && fabs(Qtake) < totQ) { Qfrac = fabs(Qtake) / totQ; for (i = 1; i < NRw; i++) b->Q[i] = b->Q[i] * (1.0 - Qfrac); }
        else { for (i = 0; i < NRw; i++) b->Q[i] = 0.0; }           /* RPARAM%MINFLOW: minFlow of the network file, 0 here */
        for (i = 1; i < NRw; i++) {
            double wc = alfa * pow(K, 1.0 / alfa) * pow(b->Q[i], (alfa - 1.0) / alfa);
            b->X[i] = fmin(h->RLENGTH[j] / wc + b->T[i], HUGE_DP);
        }
    }
    FROUTE[0] = 1; for (i = 1; i <= NQ1; i++) FROUTE[i] = 0;
    ierr = kinwav_rch(h, j, T_START, T_END, b->Q + 1, b->T + 1, b->X + 1, FROUTE + 1, NQ1, &NQ2);
    if (ierr) return ierr;
    NR = 0; for (i = 0; i <= NQ1; i++) NR += FROUTE[i]; NR -= 1;
    NN = NQ2 - NR;
    if (NR + 1 > NQ2) { snprintf(h->message, 256, "kwt_rch/no non-routed particle left (NR+1>NQ2)"); return 21; }
    ierr = interp_rch(b->X, b->Q, NR + 2, T_START, T_END, &QNEW);
    if (ierr) { snprintf(h->message, 256, "kwt_rch/interp_rch/bad bounds"); return ierr; }
    h->REACH_Q[M][j] = QNEW * h->R_WIDTH[j] + h->BASIN_QR1[j];
    Q_END = b->Q[NR] + ((b->Q[NR + 1] - b->Q[NR]) / (b->X[NR + 1] - b->X[NR])) * (T_END - b->X[NR]);
    TIMEI = b->T[NR] + ((b->T[NR + 1] - b->T[NR]) / (b->X[NR + 1] - b->X[NR])) * (T_END - b->X[NR]);
    if (NQ2 + 2 > KW_CAP) { snprintf(h->message, 256, "kwt_rch/KWAVE capacity exceeded"); return 60; }
    w->n = NQ2 + 2;
    w->QF[NR + 1] = Q_END; w->TI[NR + 1] = TIMEI; w->TR[NR + 1] = T_END; w->RF[NR + 1] = 1;
    for (i = 0; i <= NR; i++) { w->QF[i] = b->Q[i]; w->TI[i] = b->T[i]; w->TR[i] = b->X[i]; w->RF[i] = FROUTE[i]; }
    for (i = NR + 1; i <= NQ2; i++) { w->QF[i + 1] = b->Q[i]; w->TI[i + 1] = b->T[i]; w->TR[i + 1] = b->X[i]; w->RF[i + 1] = FROUTE[i]; }
    if (h->downSegId[j] <= 0 || (h->is_lake_sim && h->lakeInlet[j])) {
        for (i = 0; i <= NN; i++) { w->QF[i] = w->QF[NR + 1 + i]; w->TI[i] = w->TI[NR + 1 + i]; w->TR[i] = w->TR[NR + 1 + i]; w->RF[i] = w->RF[NR + 1 + i]; }
        w->n = NN + 1;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* main_route.f90:29-268 + route_network :273-409                                             */
/* ------------------------------------------------------------------------------------------ */
static int route_one(mro_t *h, int M, int j, double T0, double T1, wave_buf_t *b)
{
    if (h->isLake[j] && h->is_lake_sim && M != M_SUM) return lake_route(h, j, M);
    if (M == M_SUM) return accum_inst_runoff(h, j);
    if (M == M_IRF) return irf_rch(h, j);
    if (M == M_KW || M == M_DW) return kw_dw_rch(h, M, j);
    if (M == M_MC) return mc_rch(h, j);
    return kwt_rch(h, j, T0, T1, b);
}

int mro_step_ep(mro_t *h, double T0, double T1, const double *basinRunoff, const double *basinEvapo, const double *basinPrecip);
int mro_step(mro_t *h, double T0, double T1, const double *basinRunoff) { return mro_step_ep(h, T0, T1, basinRunoff, NULL, NULL); }

/* data assimilation (public_var.f90:189-191 qmodOption, qBlendPeriod, QerrTrend), and the gauge observations of the NEXT step
   only: obs [nRch] in the caller's reach order, NaN or < 0 = no gauge value there; NULL = the gauge file has no record at that
   time (gage_obs_data%time_ix = integerMissing) */
void mro_set_da(mro_t *h, int qmodOption, int qBlendPeriod, int QerrTrend) { h->qmodOption = qmodOption; h->qBlendPeriod = qBlendPeriod; h->QerrTrend = QerrTrend; }
void mro_set_obs(mro_t *h, const double *obs) { h->obsNow = obs != NULL; h->obsRow = obs; }

/* water management forcing of the NEXT steps (until changed): flux_wm / vol_wm [nRch] in the caller's reach order, NULL = off */
void mro_set_wm(mro_t *h, const double *flux_wm, const double *vol_wm, int volJumpStart)
{
    h->wmFlux = flux_wm; h->wmVol = (vol_wm && h->is_lake_sim) ? vol_wm : NULL; h->volJumpStart = volJumpStart;
}

/* basinEvapo / basinPrecip [nHRU] in the units of the runoff (NULL = no lake forcing: exactly zero in lake_route) */
int mro_step_ep(mro_t *h, double T0, double T1, const double *basinRunoff, const double *basinEvapo, const double *basinPrecip)
{
    int j, r, ierr;
    if (h->qmodOption == 1) {             /* main_route.f90:125-148: before the runoff is mapped */
        if (h->obsNow) {
            for (j = 0; j < h->nRch; j++) {
                const double qobs = h->obsRow[j];
                if ((qobs != qobs) || (qobs < 0)) continue;
                h->Qobs[j] = qobs; h->Qelapsed[j] = 0;
            }
        } else {
            for (j = 0; j < h->nRch; j++) h->Qelapsed[j] = h->Qelapsed[j] + 1;
        }
        h->obsNow = 0; h->obsRow = NULL;
    } else if (h->qmodOption != 0) { snprintf(h->message, 256, "main_route/Error: qmodOption invalid"); return 1; }
    ierr = basin2reach(h, basinRunoff, h->reachRunoff, 1);
    if (ierr) return ierr;
    h->hasEP = (h->is_lake_sim && basinEvapo && basinPrecip);
    if (h->hasEP) {                       /* main_route.f90:174-199: the same basin2reach, limitRunoff absent = .true. */
        ierr = basin2reach(h, basinEvapo, h->reachEvapo, 1);
        if (ierr) return ierr;
        ierr = basin2reach(h, basinPrecip, h->reachPrecip, 1);
        if (ierr) return ierr;
    }
    if (h->doesBasinRoute == 1) {
#pragma omp parallel for schedule(static) num_threads(h->nThreads)
        for (j = 0; j < h->nRch; j++) { h->BASIN_QI[j] = h->reachRunoff[j]; hru_irf(h, j); }
    } else {
        for (j = 0; j < h->nRch; j++) { h->BASIN_QR0[j] = h->BASIN_QR1[j]; h->BASIN_QR1[j] = h->reachRunoff[j]; }
    }
    for (r = 0; r < h->nRoutes; r++) {
        int M = h->routeOrder[r];
        if (h->nThreads <= 1) {
            wave_buf_t b = {0, 0, 0, 0, 0};
            for (j = 0; j < h->nRch; j++) { ierr = route_one(h, M, h->order[j], T0, T1, &b); if (ierr) break; }
            free(b.Q); free(b.T); free(b.X);
            if (ierr) return ierr;
        } else {
            int err = 0;
#pragma omp parallel num_threads(h->nThreads)
            {
                wave_buf_t b = {0, 0, 0, 0, 0};
                int lev;
                for (lev = 0; lev < h->nLevel; lev++) {
                    int k;
#pragma omp for schedule(dynamic, 64)
                    for (k = h->lev_ptr[lev]; k < h->lev_ptr[lev + 1]; k++) {
                        int e = route_one(h, M, h->lev_idx[k], T0, T1, &b);
                        if (e) {
#pragma omp atomic write
                            err = e;
                        }
                    }
                }
                free(b.Q); free(b.T); free(b.X);
            }
            if (err) return err;
        }
    }
    h->iTime++;
    return 0;
}

/* run nSteps with T advancing as init_model_data.f90:311-312; outputs [method slot r][step][reach] */
int mro_run(mro_t *h, int nSteps, double T0, const double *runoff /* [nSteps][nHRU] */, double *q_out /* [nRoutes][nSteps][nRch] or NULL */)
{
    int t, r, ierr; double t0 = T0, t1 = T0 + h->dt;
    for (t = 0; t < nSteps; t++) {
        ierr = mro_step(h, t0, t1, runoff + (size_t)t * h->nHRU);
        if (ierr) return ierr;
        if (q_out) for (r = 0; r < h->nRoutes; r++)
            memcpy(q_out + ((size_t)r * nSteps + t) * h->nRch, h->REACH_Q[h->routeOrder[r]], sizeof(double) * h->nRch);
        t0 = t1; t1 = t0 + h->dt;
    }
    return 0;
}

int mro_run_ep(mro_t *h, int nSteps, double T0, const double *runoff, const double *evapo, const double *precip, double *q_out)
{
    int t, r, ierr; double t0 = T0, t1 = T0 + h->dt;
    for (t = 0; t < nSteps; t++) {
        ierr = mro_step_ep(h, t0, t1, runoff + (size_t)t * h->nHRU, evapo ? evapo + (size_t)t * h->nHRU : NULL, precip ? precip + (size_t)t * h->nHRU : NULL);
        if (ierr) return ierr;
        if (q_out) for (r = 0; r < h->nRoutes; r++)
            memcpy(q_out + ((size_t)r * nSteps + t) * h->nRch, h->REACH_Q[h->routeOrder[r]], sizeof(double) * h->nRch);
        t0 = t1; t1 = t0 + h->dt;
    }
    return 0;
}
void mro_get_lake_forcing(mro_t *h, double *evapo, double *precip) { memcpy(evapo, h->reachEvapo, sizeof(double) * h->nRch); memcpy(precip, h->reachPrecip, sizeof(double) * h->nRch); }

/* ------------------------------------------------------------------------------------------ */
/* accessors                                                                                   */
/* ------------------------------------------------------------------------------------------ */
enum { F_REACH_Q = 0, F_REACH_VOL1 = 1, F_REACH_INFLOW = 2, F_WB = 3, F_BASIN_QI = 4, F_BASIN_QR1 = 5, F_BASIN_QR0 = 6, F_REACH_VOL0 = 7,
       F_QERROR = 8, F_QOBS = 9,
       F_WIDTH = 10, F_TOTAREA = 11, F_BASAREA = 12, F_SLOPE = 13 };

int mro_get(mro_t *h, int method, int field, double *out)
{
    const double *src = NULL;
    switch (field) {
        case F_REACH_Q: src = h->REACH_Q[method]; break;
        case F_REACH_VOL1: src = h->REACH_VOL1[method]; break;
        case F_REACH_VOL0: src = h->REACH_VOL0[method]; break;
        case F_REACH_INFLOW: src = h->REACH_INFLOW[method]; break;
        case F_WB: src = h->WB[method]; break;
        case F_BASIN_QI: src = h->BASIN_QI; break;
        case F_BASIN_QR1: src = h->BASIN_QR1; break;
        case F_BASIN_QR0: src = h->BASIN_QR0; break;
        case F_QERROR: src = h->Qerror[method]; break;
        case F_QOBS: src = h->Qobs; break;
        case F_WIDTH: src = h->R_WIDTH; break;
        case F_TOTAREA: src = h->TOTAREA; break;
        case F_BASAREA: src = h->BASAREA; break;
        case F_SLOPE: src = h->R_SLOPE; break;
        default: return 1;
    }
    memcpy(out, src, sizeof(double) * h->nRch);
    return 0;
}
const char *mro_message(mro_t *h) { return h->message; }
int mro_ntdh_bas(mro_t *h) { return h->ntdh_bas; }
int mro_maxtdh(mro_t *h) { return h->maxtdh; }
int mro_nlevel(mro_t *h) { return h->nLevel; }
void mro_get_frac_future(mro_t *h, double *out) { memcpy(out, h->FRAC_FUTURE, sizeof(double) * h->ntdh_bas); }
void mro_get_uh_ptr(mro_t *h, int *out) { memcpy(out, h->uh_ptr, sizeof(int) * (h->nRch + 1)); }
void mro_get_uh_val(mro_t *h, double *out) { if (h->uh_val) memcpy(out, h->uh_val, sizeof(double) * h->uh_ptr[h->nRch]); }
void mro_get_down_index(mro_t *h, int *out) { memcpy(out, h->downIndex, sizeof(int) * h->nRch); }
void mro_get_qfuture(mro_t *h, double *out) { memcpy(out, h->QFUTURE, sizeof(double) * (size_t)h->nRch * h->ntdh_bas); }
void mro_set_qfuture(mro_t *h, const double *in) { memcpy(h->QFUTURE, in, sizeof(double) * (size_t)h->nRch * h->ntdh_bas); }
void mro_get_qfuture_irf(mro_t *h, double *out) { if (h->QFUTURE_IRF) memcpy(out, h->QFUTURE_IRF, sizeof(double) * h->uh_ptr[h->nRch]); }
void mro_set_qfuture_irf(mro_t *h, const double *in) { if (h->QFUTURE_IRF) memcpy(h->QFUTURE_IRF, in, sizeof(double) * h->uh_ptr[h->nRch]); }
int mro_set(mro_t *h, int method, int field, const double *in)
{
    double *dst = NULL;
    switch (field) {
        case F_REACH_Q: dst = h->REACH_Q[method]; break;
        case F_REACH_VOL1: dst = h->REACH_VOL1[method]; break;
        case F_REACH_VOL0: dst = h->REACH_VOL0[method]; break;
        case F_BASIN_QR1: dst = h->BASIN_QR1; break;
        case F_BASIN_QR0: dst = h->BASIN_QR0; break;
        case F_QERROR: dst = h->Qerror[method]; break;
        case F_QOBS: dst = h->Qobs; break;
        default: return 1;
    }
    memcpy(dst, in, sizeof(double) * h->nRch);
    return 0;
}
int mro_n_molecule(int method) { return method >= 0 && method < N_METHOD ? N_MOLECULE[method] : 0; }
void mro_get_molecule(mro_t *h, int method, double *out) { if (h->MOL[method]) memcpy(out, h->MOL[method], sizeof(double) * (size_t)h->nRch * N_MOLECULE[method]); }
void mro_set_molecule(mro_t *h, int method, const double *in) { if (h->MOL[method]) memcpy(h->MOL[method], in, sizeof(double) * (size_t)h->nRch * N_MOLECULE[method]); }
void mro_set_itime(mro_t *h, long it) { h->iTime = it; }
void mro_get_qelapsed(mro_t *h, int *out) { memcpy(out, h->Qelapsed, sizeof(int) * h->nRch); }
void mro_set_qelapsed(mro_t *h, const int *in) { memcpy(h->Qelapsed, in, sizeof(int) * h->nRch); }
void mro_set_threads(mro_t *h, int n) { h->nThreads = n > 0 ? n : 1; }
/* KWT state in the restart layout [seg][wave] (write_restart_pio.f90:1039-1134), wave dimension = cap */
void mro_get_kwt_state(mro_t *h, int cap, int *numWaves, double *qf, double *ti, double *tr, unsigned char *rf)
{
    int i, k;
    for (i = 0; i < h->nRch; i++) {
        numWaves[i] = h->KW[i].n;
        for (k = 0; k < cap; k++) {
            int ok = k < h->KW[i].n;
            qf[(size_t)i * cap + k] = ok ? h->KW[i].QF[k] : -9999.0;
            ti[(size_t)i * cap + k] = ok ? h->KW[i].TI[k] : -9999.0;
            tr[(size_t)i * cap + k] = ok ? h->KW[i].TR[k] : -9999.0;
            rf[(size_t)i * cap + k] = ok ? h->KW[i].RF[k] : 0;
        }
    }
}
void mro_set_kwt_state(mro_t *h, int cap, const int *numWaves, const double *qf, const double *ti, const double *tr, const unsigned char *rf)
{
    int i, k;
    for (i = 0; i < h->nRch; i++) {
        h->KW[i].n = numWaves[i];
        for (k = 0; k < numWaves[i] && k < KW_CAP; k++) {
            h->KW[i].QF[k] = qf[(size_t)i * cap + k]; h->KW[i].TI[k] = ti[(size_t)i * cap + k];
            h->KW[i].TR[k] = tr[(size_t)i * cap + k]; h->KW[i].RF[k] = rf[(size_t)i * cap + k];
        }
    }
}
/* stand-alone helpers exposed for unit tests */
double mro_gammp(double a, double x) { return gammp(a, x); }
int mro_make_uh_one(double len, double dt, double velo, double diff, double *out) { return make_uh_one(len, dt, velo, diff, out); }
int mro_interp_rch(const double *T, const double *Q, int n, double T0, double T1, double *out) { return interp_rch(T, Q, n, T0, T1, out); }
int mro_remove_rch(double *Q, double *T, double *X, int n)
{
    wave_buf_t b = {0, 0, 0, 0, 0}; int i;
    wb_reserve(&b, n); memcpy(b.Q, Q, sizeof(double) * n); memcpy(b.T, T, sizeof(double) * n); memcpy(b.X, X, sizeof(double) * n); b.n = n;
    remove_rch(&b);
    for (i = 0; i < b.n; i++) { Q[i] = b.Q[i]; T[i] = b.T[i]; X[i] = b.X[i]; }
    n = b.n; free(b.Q); free(b.T); free(b.X);
    return n;
}

/* ------------------------------------------------------------------------------------------ */
/* remap_1D_runoff, process_remap.f90:164-262 (one time step).  hru_ix / qhru_ix are 0-based,  */
/* -1 = integerMissing; basinRunoff is zeroed by the caller (get_basin_runoff.f90:73).         */
/* ------------------------------------------------------------------------------------------ */
void mro_remap_1d(int nMap, const int *hru_ix, const int *num_qhru, const int *qhru_ix, const double *weight,
                  const double *sim, double *basinRunoff)
{
    const double xTol = 1.e-6;
    int ixOverlap = 0, iHRU, ixPoly;
    for (iHRU = 0; iHRU < nMap; iHRU++) {
        int jHRU = hru_ix[iHRU];
        double sumWeights;
        if (jHRU < 0) { ixOverlap = ixOverlap + num_qhru[iHRU]; continue; }
        sumWeights = 0.0;
        basinRunoff[jHRU] = 0.0;
        for (ixPoly = 0; ixPoly < num_qhru[iHRU]; ixPoly++) {
            int ixRunoff;
            if (qhru_ix[ixOverlap] < 0) { ixOverlap = ixOverlap + 1; continue; }
            ixRunoff = qhru_ix[ixOverlap];
            if (sim[ixRunoff] > -xTol) {
                sumWeights = sumWeights + weight[ixOverlap];
                basinRunoff[jHRU] = basinRunoff[jHRU] + weight[ixOverlap] * sim[ixRunoff];
            }
            ixOverlap = ixOverlap + 1;
        }
        if (sumWeights > xTol) {
            if (fabs(1.0 - sumWeights) > xTol) basinRunoff[jHRU] = basinRunoff[jHRU] / sumWeights;
        }
    }
}
