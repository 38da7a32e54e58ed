/* TEST INFRASTRUCTURE -- never shipped, never on the product path.
 *
 * A stand-in for libmizuroute_b200.so that implements the part of the C ABI (include/mizuroute_b200.h) the
 * stand-alone host route_runoff.cpp calls, on top of the CPU oracle (oracle/mr_oracle.c).  tests/test_host_stub.py
 * links a second copy of the host against it (tests/stub/_build/, git-ignored) so that the host's own logic -- control
 * file, time map, forcing ingest, history files, aggregation, restart files -- is checked end to end without a GPU.
 * The numbers the stub produces are the oracle's, so these tests say nothing about the CUDA path; the `-m gpu` tests
 * run the real host against the real library.
 */
#include "../../oracle/mr_oracle.c"
#include "../../include/mizuroute_b200.h"

struct mr_handle_s {
    mr_options o;
    mro_t *m;
    int nRch, nHRU;
    /* remap (mr_set_remap) */
    int nForcing, nMap, *mapHru, *numQ, *qIx; double *wgt;
    /* named lake parameters and the simulation start, applied to the oracle in mr_set_network */
    char lpName[64][32]; double *lpVal[64]; int nLp;
    int hasStart, sy, sm, sd, noleap; double ssec;
    /* water management of the next batch (mr_upload_wm) */
    double *wmF, *wmV; int wmSteps, wmJump;
    /* forcing ingest (mr_set_ingest / mr_ingest_records): rows left resident for mr_route_resident, its REACH_Q for mr_download_q */
    int *ingSrc; int ingCols, ingRescale; double ingA, ingB, ingFill; double *rows; int rowSteps; double *qLast;
    /* gauge observations of the next batch (mr_upload_obs) */
    double *obs; int *obsHas; int obsSteps, qmod;
    /* lake forcing of the next batch (mr_upload_lake_forcing) */
    double *ev, *pr; int epSteps;
    /* BASIN_QR(1) of the steps of the last batch */
    double *qr; int qrSteps;
    /* REACH_Q of the last batch [route][step][reach] and the open history period (mr_history_means) */
    double *qKeep; int qKeepSteps; double *histAcc; int histCount;
};

static void say(char *message, const char *txt)
{
    if (message) { memset(message, 0, MR_STRLEN); strncpy(message, txt, MR_STRLEN - 1); }
}

int mr_create(const mr_options *opts, mr_handle *out, char *message)
{
    struct mr_handle_s *h = (struct mr_handle_s *)calloc(1, sizeof *h);
    h->o = *opts;
    *out = h;
    say(message, "");
    return 0;
}

int mr_set_network(mr_handle h, int nRch, int nHRU, const int *segId, const int *downSegId, const int *hruSegId, const double *hruArea,
                   const double *length, const double *slope, const double *width, const double *man_n, const int *islake,
                   const int *lakeModelType, const double *D03_MaxStorage, const double *D03_Coefficient, const double *D03_Power,
                   const double *D03_S0, char *message)
{
    char ropt[16]; int r;
    for (r = 0; r < h->o.n_routes; r++) ropt[r] = (char)('0' + h->o.route_methods[r]);
    ropt[h->o.n_routes] = 0;
    h->nRch = nRch; h->nHRU = nHRU;
    h->m = mro_create(nRch, nHRU, segId, downSegId, hruSegId, hruArea, length, slope, width, man_n, islake, lakeModelType, D03_MaxStorage,
                      D03_Coefficient, D03_Power, D03_S0, h->o.dt, ropt, h->o.doesBasinRoute, h->o.hw_drain_point, h->o.min_length_route,
                      h->o.is_lake_sim, h->o.lakeRegulate, h->o.LakeInputOption, h->o.runoffMin, h->o.time_conv, h->o.length_conv,
                      h->o.fshape, h->o.tscale, h->o.velo, h->o.diff, h->o.mann_n, h->o.wscale, 1);
    if (!h->m) { say(message, "mr_set_network/oracle refused the network"); return 20; }
    { int k; for (k = 0; k < h->nLp; k++) if (mro_set_lake_param(h->m, h->lpName[k], h->lpVal[k])) { say(message, "mr_set_network/unknown lake parameter"); return 20; } }
    if (h->hasStart) mro_set_sim_start(h->m, h->sy, h->sm, h->sd, h->ssec, h->noleap);
    if (h->o.floodplain) mro_set_channel(h->m, 1, h->o.dscale > 0.0 ? h->o.dscale : (double)0.000045f, h->o.floodplainSlope > 0.0 ? h->o.floodplainSlope : 1000.0);
    say(message, "");
    return 0;
}

int mr_set_remap(mr_handle h, int nForcing, int nMap, const int *mapHruIndex, const int *numQhru, const int *qhruIndex, const double *weight, char *message)
{
    int i, tot = 0;
    for (i = 0; i < nMap; i++) tot += numQhru[i];
    h->nForcing = nForcing; h->nMap = nMap;
    h->mapHru = (int *)malloc(sizeof(int) * (size_t)(nMap + 1)); memcpy(h->mapHru, mapHruIndex, sizeof(int) * (size_t)nMap);
    h->numQ = (int *)malloc(sizeof(int) * (size_t)(nMap + 1)); memcpy(h->numQ, numQhru, sizeof(int) * (size_t)nMap);
    h->qIx = (int *)malloc(sizeof(int) * (size_t)(tot + 1)); memcpy(h->qIx, qhruIndex, sizeof(int) * (size_t)tot);
    h->wgt = (double *)malloc(sizeof(double) * (size_t)(tot + 1)); memcpy(h->wgt, weight, sizeof(double) * (size_t)tot);
    say(message, "");
    return 0;
}

int mr_step_batch(mr_handle h, int nSteps, double T0, const double *runoff, double *q_out, char *message)
{
    int t, r, ierr; double t0 = T0, t1 = T0 + h->o.dt;
    if (h->epSteps && h->epSteps != nSteps) { say(message, "mr_step_batch/lake forcing was uploaded for a different number of steps"); return 1; }
    double *row = (double *)calloc((size_t)h->nHRU + 1, sizeof(double));
    free(h->qr); h->qr = (double *)malloc(sizeof(double) * (size_t)nSteps * (size_t)h->nRch); h->qrSteps = nSteps;
    free(h->qKeep); h->qKeep = (double *)malloc(sizeof(double) * ((size_t)h->o.n_routes * (size_t)nSteps * (size_t)h->nRch + 1)); h->qKeepSteps = nSteps;
    for (t = 0; t < nSteps; t++) {
        const double *in = runoff + (size_t)t * (size_t)(h->nMap ? h->nForcing : h->nHRU);
        if (h->nMap) { mro_remap_1d(h->nMap, h->mapHru, h->numQ, h->qIx, h->wgt, in, row); in = row; }
        if (h->wmSteps) mro_set_wm(h->m, h->wmF ? h->wmF + (size_t)t * h->nRch : NULL, h->wmV ? h->wmV + (size_t)t * h->nRch : NULL, h->wmJump);
        else mro_set_wm(h->m, NULL, NULL, 0);
        if (h->qmod == 1) mro_set_obs(h->m, (h->obsSteps && h->obsHas[t]) ? h->obs + (size_t)t * h->nRch : NULL);
        ierr = h->epSteps ? mro_step_ep(h->m, t0, t1, in, h->ev + (size_t)t * h->nHRU, h->pr + (size_t)t * h->nHRU) : mro_step(h->m, t0, t1, in);
        if (ierr) { char b[MR_STRLEN]; snprintf(b, sizeof b, "mr_step_batch/main_route/%s", mro_message(h->m)); say(message, b); free(row); return ierr; }
        for (r = 0; r < h->m->nRoutes; r++) {
            if (q_out) memcpy(q_out + ((size_t)r * nSteps + t) * h->nRch, h->m->REACH_Q[h->m->routeOrder[r]], sizeof(double) * (size_t)h->nRch);
            memcpy(h->qKeep + ((size_t)r * nSteps + t) * h->nRch, h->m->REACH_Q[h->m->routeOrder[r]], sizeof(double) * (size_t)h->nRch);
        }
        memcpy(h->qr + (size_t)t * h->nRch, h->m->BASIN_QR1, sizeof(double) * (size_t)h->nRch);
        t0 = t1; t1 = t0 + h->o.dt;
    }
    free(row);
    h->epSteps = 0; h->wmSteps = 0; h->obsSteps = 0;
    say(message, "");
    return 0;
}

int mr_set_ingest(mr_handle h, int nForcing, const int *forcingOfHru, double scale, double offset, double fill, char *message)
{
    free(h->ingSrc); h->ingSrc = (int *)malloc(sizeof(int) * (size_t)(h->nHRU + 1)); memcpy(h->ingSrc, forcingOfHru, sizeof(int) * (size_t)h->nHRU);
    h->ingCols = nForcing; h->ingRescale = (scale != -9999.0 || offset != -9999.0);
    h->ingA = scale == -9999.0 ? 1.0 : scale; h->ingB = offset == -9999.0 ? 0.0 : offset; h->ingFill = fill;
    say(message, "");
    return 0;
}

/* plain restatement of read_1D_forcing's weighted mean + scale_forcing + sort_flux (the device code is mr_ingest.h) */
int mr_ingest_records(mr_handle h, int nSteps, int nRec, const double *records, const int *recPtr, const int *recIdx, const double *recFrac, char *message)
{
    int t, i, j; (void)nRec;
    free(h->rows); h->rows = (double *)malloc(sizeof(double) * ((size_t)nSteps * (size_t)h->nHRU + 1)); h->rowSteps = nSteps;
    for (t = 0; t < nSteps; t++) for (i = 0; i < h->nHRU; i++) {
        const int src = h->ingSrc[i]; double v = 0.0;
        if (src >= 0) {
            if (!recFrac) v = records[(size_t)recIdx[recPtr[t]] * h->ingCols + src];
            else {
                double wsum = 0.0, wtot = 0.0;
                for (j = recPtr[t]; j < recPtr[t + 1]; j++) { const double x = records[(size_t)recIdx[j] * h->ingCols + src]; if (x != h->ingFill) { wsum += x * recFrac[j]; wtot += recFrac[j]; } }
                v = wtot == 0.0 ? h->ingFill : (wtot < 1.0 ? wsum / wtot : wsum);
            }
            if (h->ingRescale && v != h->ingFill && v != -9999.0) v = h->ingA * v + h->ingB;
            if (v == h->ingFill || v < 0.0) v = 0.0;
        }
        h->rows[(size_t)t * h->nHRU + i] = v;
    }
    say(message, "");
    return 0;
}

int mr_route_resident(mr_handle h, int nSteps, double T0, char *message)
{
    if (!h->rows || h->rowSteps != nSteps) { say(message, "mr_route_resident/no resident rows for that many steps"); return 1; }
    free(h->qLast); h->qLast = (double *)malloc(sizeof(double) * ((size_t)h->o.n_routes * (size_t)nSteps * (size_t)h->nRch + 1));
    return mr_step_batch(h, nSteps, T0, h->rows, h->qLast, message);
}

int mr_download_q(mr_handle h, int nSteps, double *q_out, char *message)
{
    memcpy(q_out, h->qLast, sizeof(double) * (size_t)h->o.n_routes * (size_t)nSteps * (size_t)h->nRch);
    say(message, "");
    return 0;
}

int mr_set_lake_param(mr_handle h, const char *name, int n, const double *values, char *message)
{
    if (h->nLp >= 64) { say(message, "mr_set_lake_param/too many parameters for the stub"); return 1; }
    strncpy(h->lpName[h->nLp], name, 31); h->lpName[h->nLp][31] = 0;
    h->lpVal[h->nLp] = (double *)malloc(sizeof(double) * (size_t)n); memcpy(h->lpVal[h->nLp], values, sizeof(double) * (size_t)n);
    h->nLp++;
    say(message, "");
    return 0;
}

int mr_set_sim_start(mr_handle h, int year, int month, int day, double secOfDay, int noleap, char *message)
{
    h->hasStart = 1; h->sy = year; h->sm = month; h->sd = day; h->ssec = secOfDay; h->noleap = noleap;
    if (h->m) mro_set_sim_start(h->m, year, month, day, secOfDay, noleap);
    say(message, "");
    return 0;
}

int mr_set_da(mr_handle h, int qmodOption, int qBlendPeriod, int QerrTrend, char *message)
{
    h->qmod = qmodOption; mro_set_da(h->m, qmodOption, qBlendPeriod, QerrTrend);
    say(message, "");
    return 0;
}

int mr_upload_obs(mr_handle h, int nSteps, const int *hasRecord, const double *obs, char *message)
{
    const size_t n = (size_t)nSteps * (size_t)h->nRch; int t;
    free(h->obs); free(h->obsHas);
    h->obs = (double *)malloc(sizeof(double) * (n + 1)); memcpy(h->obs, obs, sizeof(double) * n);
    h->obsHas = (int *)malloc(sizeof(int) * (size_t)(nSteps + 1));
    for (t = 0; t < nSteps; t++) h->obsHas[t] = hasRecord ? hasRecord[t] : 1;
    h->obsSteps = nSteps;
    say(message, "");
    return 0;
}

int mr_upload_wm(mr_handle h, int nSteps, const double *flux_wm, const double *vol_wm, int volJumpStart, char *message)
{
    const size_t n = (size_t)nSteps * (size_t)h->nRch;
    free(h->wmF); free(h->wmV); h->wmF = h->wmV = NULL;
    if (flux_wm) { h->wmF = (double *)malloc(sizeof(double) * (n + 1)); memcpy(h->wmF, flux_wm, sizeof(double) * n); }
    if (vol_wm) { h->wmV = (double *)malloc(sizeof(double) * (n + 1)); memcpy(h->wmV, vol_wm, sizeof(double) * n); }
    h->wmSteps = nSteps; h->wmJump = volJumpStart;
    say(message, "");
    return 0;
}

int mr_upload_lake_forcing(mr_handle h, int nSteps, const double *basinEvapo, const double *basinPrecip, char *message)
{
    const size_t n = (size_t)nSteps * (size_t)h->nHRU;
    free(h->ev); free(h->pr);
    h->ev = (double *)malloc(sizeof(double) * (n + 1)); h->pr = (double *)malloc(sizeof(double) * (n + 1));
    memcpy(h->ev, basinEvapo, sizeof(double) * n); memcpy(h->pr, basinPrecip, sizeof(double) * n);
    h->epSteps = nSteps;
    say(message, "");
    return 0;
}

int mr_download_basin_q(mr_handle h, int nSteps, double *qr_out, char *message)
{
    if (nSteps > h->qrSteps) { say(message, "mr_download_basin_q/more steps than the last batch routed"); return 1; }
    memcpy(qr_out, h->qr, sizeof(double) * (size_t)nSteps * (size_t)h->nRch);
    say(message, "");
    return 0;
}

/* histVars_data.f90:154-246: sums in step order in double precision, mean, rounded to float32 */
int mr_history_means(mr_handle h, int nSteps, int nAgg, int wantDlay, int flush, int maxPeriods, float *out, int *nPeriods, char *message)
{
    const size_t N = (size_t)h->nRch; const int nr = h->o.n_routes, nSeries = nr + (wantDlay ? 1 : 0);
    int s, t, per = 0, c = 0; size_t i;
    if (nSteps > h->qKeepSteps) { say(message, "mr_history_means/more steps requested than the last batch routed"); return 1; }
    if (!h->histAcc) h->histAcc = (double *)calloc((size_t)(nr + 1) * N, sizeof(double));
    if ((h->histCount + nSteps) / nAgg + ((flush && (h->histCount + nSteps) % nAgg) ? 1 : 0) > maxPeriods) { say(message, "mr_history_means/more periods complete than the output buffer holds"); return 1; }
    for (s = 0; s < nSeries; s++) {
        double *acc = h->histAcc + (size_t)s * N;
        per = 0; c = h->histCount;
        for (t = 0; t < nSteps; t++) {
            const double *row = s < nr ? h->qKeep + ((size_t)s * h->qKeepSteps + t) * N : h->qr + (size_t)t * N;
            if (c == 0) for (i = 0; i < N; i++) acc[i] = 0.0;
            for (i = 0; i < N; i++) acc[i] = acc[i] + row[i];
            if (++c == nAgg) { for (i = 0; i < N; i++) out[((size_t)per * nSeries + s) * N + i] = (float)(acc[i] / (double)c); per++; c = 0; }
        }
        if (flush && c > 0) { for (i = 0; i < N; i++) out[((size_t)per * nSeries + s) * N + i] = (float)(acc[i] / (double)c); per++; c = 0; }
    }
    h->histCount = c;
    *nPeriods = per;
    say(message, "");
    return 0;
}

long mr_get_info(mr_handle h, int what)
{
    switch (what) {
        case MR_INFO_NRCH: return h->nRch;
        case MR_INFO_NHRU: return h->nHRU;
        case MR_INFO_NTDH_BAS: return h->m->ntdh_bas;
        case MR_INFO_MAXTDH: return h->m->maxtdh;
        case MR_INFO_STEPS_DONE: return h->m->iTime - 1;
        case MR_INFO_NFORCING: return h->nMap ? h->nForcing : h->nHRU;
        default: return -1;
    }
}

int mr_get_reach_uh(mr_handle h, int *ntdh, double *uh, char *message)
{
    int i, k; const int mx = h->m->maxtdh;
    for (i = 0; i < h->nRch; i++) {
        const int a = h->m->uh_ptr[i], n = h->m->uh_ptr[i + 1] - a;
        ntdh[i] = n;
        for (k = 0; k < mx; k++) uh[(size_t)i * mx + k] = k < n ? h->m->uh_val[a + k] : 0.0;
    }
    say(message, "");
    return 0;
}

static int method_of(mr_handle h, int code) { int r; for (r = 0; r < h->o.n_routes; r++) if (h->o.route_methods[r] == code) return 1; return 0; }

/* KWT state pieces go through the oracle's whole-state accessors */
typedef struct { int *n; double *qf, *ti, *tr; unsigned char *rf; } kw_t;
static kw_t kw_get(mr_handle h)
{
    kw_t s; const size_t N = (size_t)h->nRch, W = MR_KW_SLOTS;
    s.n = (int *)malloc(sizeof(int) * N); s.qf = (double *)malloc(8 * N * W); s.ti = (double *)malloc(8 * N * W); s.tr = (double *)malloc(8 * N * W);
    s.rf = (unsigned char *)malloc(N * W);
    mro_get_kwt_state(h->m, MR_KW_SLOTS, s.n, s.qf, s.ti, s.tr, s.rf);
    return s;
}
static void kw_free(kw_t s) { free(s.n); free(s.qf); free(s.ti); free(s.tr); free(s.rf); }

int mr_get_flux(mr_handle h, int method, int field, double *out, char *message)
{
    say(message, "");
    if (mro_get(h->m, method, field, out)) { say(message, "mr_get_flux/unknown field"); return 1; }     /* the field ids are the oracle's */
    return 0;
}

int mr_get_state(mr_handle h, int var, void *buf, long nbytes, char *message)
{
    const size_t N = (size_t)h->nRch, W = MR_KW_SLOTS; size_t i, k; int r;
    double *d = (double *)buf; int *ip = (int *)buf;
    (void)nbytes;
    say(message, "");
    switch (var) {
        case MR_ST_BASIN_QFUTURE: mro_get_qfuture(h->m, d); return 0;
        case MR_ST_BASIN_QR: for (i = 0; i < N; i++) { d[2 * i] = h->m->BASIN_QR0[i]; d[2 * i + 1] = h->m->BASIN_QR1[i]; } return 0;
        case MR_ST_IRF_QFUTURE: {
            const int mx = h->m->maxtdh;
            for (i = 0; i < N; i++) { const int a = h->m->uh_ptr[i], n = h->m->uh_ptr[i + 1] - a;
                                      for (k = 0; k < (size_t)mx; k++) d[i * mx + k] = (int)k < n ? h->m->QFUTURE_IRF[a + k] : 0.0; }
            return 0; }
        case MR_ST_IRF_VOL: memcpy(d, h->m->REACH_VOL1[M_IRF], 8 * N); return 0;
        case MR_ST_LAKE_VOL: for (r = 0; r < h->o.n_routes; r++) memcpy(d + (size_t)r * N, h->m->REACH_VOL1[h->o.route_methods[r]], 8 * N); return 0;
        case MR_ST_KWT_NWAVE: case MR_ST_KWT_QWAVE: case MR_ST_KWT_TENTRY: case MR_ST_KWT_TEXIT: case MR_ST_KWT_ROUTED: {
            kw_t s;
            if (!method_of(h, MR_KINEMATIC_WAVE_TRACKING)) { say(message, "mr_get_state/KWT is not active"); return 1; }
            s = kw_get(h);
            if (var == MR_ST_KWT_NWAVE) memcpy(ip, s.n, sizeof(int) * N);
            else if (var == MR_ST_KWT_QWAVE) memcpy(d, s.qf, 8 * N * W);
            else if (var == MR_ST_KWT_TENTRY) memcpy(d, s.ti, 8 * N * W);
            else if (var == MR_ST_KWT_TEXIT) memcpy(d, s.tr, 8 * N * W);
            else for (i = 0; i < N * W; i++) ip[i] = s.rf[i];
            kw_free(s);
            return 0; }
        case MR_ST_MOLECULE_KW: case MR_ST_MOLECULE_MC: case MR_ST_MOLECULE_DW: mro_get_molecule(h->m, M_KW + (var - MR_ST_MOLECULE_KW), d); return 0;
        case MR_ST_QERROR: for (r = 0; r < h->o.n_routes; r++) memcpy(d + (size_t)r * N, h->m->Qerror[h->o.route_methods[r]], 8 * N); return 0;
        case MR_ST_DA_QOBS: memcpy(d, h->m->Qobs, 8 * N); return 0;
        case MR_ST_DA_QELAPSED: memcpy(ip, h->m->Qelapsed, 4 * N); return 0;
        default: say(message, "mr_get_state/unknown state variable"); return 1;
    }
}

int mr_set_state(mr_handle h, int var, const void *buf, long nbytes, char *message)
{
    const size_t N = (size_t)h->nRch, W = MR_KW_SLOTS; size_t i, k; int r;
    const double *d = (const double *)buf; const int *ip = (const int *)buf;
    (void)nbytes;
    say(message, "");
    switch (var) {
        case MR_ST_BASIN_QFUTURE: mro_set_qfuture(h->m, d); return 0;
        case MR_ST_BASIN_QR: for (i = 0; i < N; i++) { h->m->BASIN_QR0[i] = d[2 * i]; h->m->BASIN_QR1[i] = d[2 * i + 1]; } return 0;
        case MR_ST_IRF_QFUTURE: {
            const int mx = h->m->maxtdh;
            for (i = 0; i < N; i++) { const int a = h->m->uh_ptr[i], n = h->m->uh_ptr[i + 1] - a; for (k = 0; (int)k < n; k++) h->m->QFUTURE_IRF[a + k] = d[i * mx + k]; }
            return 0; }
        case MR_ST_IRF_VOL: memcpy(h->m->REACH_VOL1[M_IRF], d, 8 * N); return 0;
        case MR_ST_LAKE_VOL:
            for (r = 0; r < h->o.n_routes; r++) if (h->o.route_methods[r] != MR_ACCUM_RUNOFF) memcpy(h->m->REACH_VOL1[h->o.route_methods[r]], d + (size_t)r * N, 8 * N);
            return 0;
        case MR_ST_KWT_NWAVE: case MR_ST_KWT_QWAVE: case MR_ST_KWT_TENTRY: case MR_ST_KWT_TEXIT: case MR_ST_KWT_ROUTED: {
            kw_t s = kw_get(h);
            if (var == MR_ST_KWT_NWAVE) memcpy(s.n, ip, sizeof(int) * N);
            else if (var == MR_ST_KWT_QWAVE) memcpy(s.qf, d, 8 * N * W);
            else if (var == MR_ST_KWT_TENTRY) memcpy(s.ti, d, 8 * N * W);
            else if (var == MR_ST_KWT_TEXIT) memcpy(s.tr, d, 8 * N * W);
            else for (i = 0; i < N * W; i++) s.rf[i] = (unsigned char)(ip[i] != 0);
            mro_set_kwt_state(h->m, MR_KW_SLOTS, s.n, s.qf, s.ti, s.tr, s.rf);
            kw_free(s);
            return 0; }
        case MR_ST_MOLECULE_KW: case MR_ST_MOLECULE_MC: case MR_ST_MOLECULE_DW: mro_set_molecule(h->m, M_KW + (var - MR_ST_MOLECULE_KW), d); return 0;
        case MR_ST_QERROR: for (r = 0; r < h->o.n_routes; r++) memcpy(h->m->Qerror[h->o.route_methods[r]], d + (size_t)r * N, 8 * N); return 0;
        case MR_ST_DA_QOBS: memcpy(h->m->Qobs, d, 8 * N); return 0;
        case MR_ST_DA_QELAPSED: memcpy(h->m->Qelapsed, ip, 4 * N); return 0;
        default: say(message, "mr_set_state/unknown state variable"); return 1;
    }
}

int mr_set_steps_done(mr_handle h, long steps, char *message)
{
    mro_set_itime(h->m, steps + 1);
    say(message, "");
    return 0;
}

void mr_destroy(mr_handle h)
{
    if (!h) return;
    if (h->m) mro_destroy(h->m);
    free(h->mapHru); free(h->numQ); free(h->qIx); free(h->wgt); free(h->qr); free(h->ev); free(h->pr); free(h->wmF); free(h->wmV); free(h->qKeep); free(h->histAcc);
    free(h);
}
