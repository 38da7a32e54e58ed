// Lake reaches: lake_route (lake_route.f90:87-229,398-438,466-470) for the endorheic, Doll-2003 and HYPE types, with the lake forcing
// of main_route.f90:174-199,243-249 -- reach-level evaporation and precipitation, produced by the same basin2reach as the
// runoff -- when the caller supplies it (mr_upload_lake_forcing); without it both are exactly zero.
// Compiles for the device (called from route_reach / kwt_task) and for the host (tests/emul), where it must match the
// oracle bit for bit.
#pragma once
#include <cmath>
#include "mr_dev.h"

#if defined(__CUDACC__)
#define MR_LAKE_FN __device__
#else
#define MR_LAKE_FN inline
#endif

namespace mr {

// basin2reach (process_remap.f90:372-420, limitRunoff absent = .true.) of one reach for one HRU-level flux row
MR_LAKE_FN double lake_basin2reach(const DevNet &d, int p, const double *flux) {
    const int h0 = d.hruPtr[p], h1 = d.hruPtr[p + 1];
    if (h1 <= h0) return d.runoffMin;
    double r = 0.0;
    for (int m = h0; m < h1; ++m) {
        const double ro = flux[d.hruIdx[m]];
        if (ro < -1.e-3) raise(d.err, 20, p, E_NEG_RUNOFF);        // negRunoffTol, public_var.f90:31
        r = r + d.hruWgt[m] * ro * d.tconv * d.lconv;
    }
    if (r < d.runoffMin) r = d.runoffMin;
    return r * d.basArea[p];
}

// comp_reach_wb with lakeFlag (water_balance.f90:61-87): precipitation and evaporation enter the balance
MR_LAKE_FN double lake_wb(double v1, double v0, double qup, double qlat, double q, double dt, double pr, double ev, bool ep, double took = 0.0) {
    const double dVol = v1 - v0;
    const double Qin = qup * dt, Qlateral = qlat * dt;
    const double precip = ep ? pr * dt : 0.0;
    const double Qout = -1.0 * q * dt;
    const double Qtake = -1.0 * took * dt;
    const double evapo = ep ? -1.0 * ev * dt : 0.0;
    return dVol - (Qin + Qlateral + precip + Qtake + Qout + evapo);
}

// HYPE reservoir outflow (lake_route.f90:398-438) of a lake holding volume v1 on day-of-year doy.  Out of line and fed
// through a pointer into HBM (a DevNet reference would force the kernel's parameter block into local memory, a dozen scalar
// arguments would raise the register pressure at the call site), so that sin() / pow() do not weigh on the wavefront
// kernels that merely may meet such a lake.
MR_DEV_NOINLINE double hype_outflow(const HypeParams *hp, int doy, double v1, double dt) {
    const HypeParams h = *hp;
    const double ELE = v1 / h.A_avg + h.E_zero;
    const double F_sin = fmax(0.0, (1 + h.Qrate_amp * sin(2 * 3.14159265359 * (doy + (int)h.Qrate_phs) / 365)));   // pi, public_var.f90:16
    const double F_lin = fmin(fmax((ELE - h.E_min) / (h.E_lim - h.E_min), 0.0), 1.0);
    const int F_prim = h.prim_F != 0.0 ? 1 : 0;
    const double Q_prim = F_sin * F_lin * F_prim * h.Qrate_prim;
    double Q_spill = 0.0;
    if (ELE > h.E_emr) Q_spill = h.Qrate_emr * pow(ELE - h.E_emr, h.Erate_emr);
    const double Q_sim = h.Qsim_mode != 0.0 ? Q_prim + Q_spill : fmax(Q_prim, Q_spill);
    return fmin(Q_sim, fmax(0.0, (ELE - h.E_min) * h.A_avg) / dt);
}

// Hanasaki-2006 release (lake_route.f90:231-396, no water-management demand) of the lake L holding volume v1; updates the
// lake's inflow memory, monthly mean inflows and release coefficient as the reference updates RCHFLX / RPARAM.  `prevQ` is
// REACH_Q of the previous step (kept when the storage ratio c is negative).  Out of line, pointer arguments only.
MR_DEV_NOINLINE double h06_release(H06Lake *Lk, double *memAll, int month, int day, int noleap, double v1, double qup, double prevQ, double dt) {
    H06Lake &L = *Lk;
    if (L.memF) {
        double *mem = memAll + L.memOff;
        if (!L.filled) {                              // first call: every row holds the monthly parameter, nothing is inserted
            for (int k = 0; k < 12; ++k) { for (int i = 0; i < L.L31; ++i) mem[(size_t)k * L.L31 + i] = L.I[k]; L.head[k] = 0; }
        } else {                                      // shift the month's row by one and insert the inflow: the ring steps back
            const int k = month - 1;
            int hd = L.head[k] - 1; if (hd < 0) hd += L.L31;
            mem[(size_t)k * L.L31 + hd] = qup;
            L.head[k] = hd;
        }
        // means, newest value first (sum(QPASTUP(m, 1:n))/n).  Rows other than the month's did not change since their mean
        // was last taken, so only that one is recomputed -- all of them on the first call.  November is never updated.
        for (int k = 0; k < 12; ++k) {
            if (k == 10 || (L.filled && k != month - 1)) continue;
            const int n = k == 1 ? (noleap ? L.LFnoleap : L.LF) : ((k == 3 || k == 5 || k == 8) ? L.L30 : L.L31);
            const double *row = mem + (size_t)k * L.L31;
            double sum = 0.0;
            int idx = L.head[k];
            for (int i = 0; i < n; ++i) { sum += row[idx]; if (++idx == L.L31) idx = 0; }
            L.I[k] = sum / n;
        }
        L.filled = 1;
    }
    double sI = 0.0, sD = 0.0;
    for (int k = 0; k < 12; ++k) { sI += L.I[k]; sD += L.D[k]; }
    const double I_yearly = sI / 12, D_yearly = sD / 12;
    const double c = L.Smax / (I_yearly * 365 * 86400.0);
    int start_month = 0;
    for (int i = 1; i <= 12; ++i) if (I_yearly <= L.I[i - 1]) start_month = i + 1;
    if (month == start_month && day == 1) L.E_rel_ini = v1 / (L.alpha * L.Smax);
    double target_r;
    if (L.purpose == 1) {
        if (L.envfact * I_yearly <= D_yearly) target_r = L.I[month - 1] * L.c1 + I_yearly * L.c2 * (L.D[month - 1] / D_yearly);
        else target_r = I_yearly + L.D[month - 1] - D_yearly;
    } else target_r = I_yearly;
    double q = prevQ;
    if (c >= L.c_compare) q = target_r * L.E_rel_ini;
    else if (0 <= c && c < L.c_compare) {
        const double ratio = pow(c / L.denominator, L.exponent);
        q = L.E_rel_ini * target_r * ratio + qup * (1 - ratio);
    }
    if (v1 < (L.Smax * L.frac_Sdead)) {
        q = q - (L.Smax * L.frac_Sdead - v1) / dt;
        if (q < 0) q = 0;
    } else if (v1 > L.Smax) {
        q = q + (v1 - L.Smax) / dt;
    }
    return q;
}

// HY: the domain holds HYPE or Hanasaki reservoirs.  The kernels are instantiated for both values and the launch picks one,
// so that a domain without them runs exactly the code it ran before those models existed (no extra registers, stack or spills).
template <int M, bool HY = false>
MR_LAKE_FN void lake_reach(const DevNet &d, int p, int t, long long tau) {
    const int N = d.nRch;
    double *Qs = d.qSer[M] + (size_t)t * N;
    const int u0 = d.upPtr[p], u1 = d.upPtr[p + 1];
    const double dt = d.dt;
    double qup = 0.0;
    for (int m = u0; m < u1; ++m) qup = qup + Qs[d.upIdx[m]];
    const int type = d.lakeType[p];
    double v1 = d.vol1[M][p];
    // water management (HY instantiation only): the lake follows a target volume / loses or gains a flux
    const bool follows = HY && d.lakeTargVol && d.lakeTargVol[p];
    const double target = (HY && d.wmVol) ? d.wmVol[(size_t)t * N + p] : 0.0;      // REACH_WM_VOL, 0 when is_vol_wm is off
    if (tau == 0 && follows && d.volJumpStart) {       // lake_route.f90:137-139
        v1 = target;
    } else
    if (tau == 0) {                                    // iTime==1 cold start, lake_route.f90:139-157
        if (type == MR_LAKE_ENDORHEIC) v1 = d.d03S0[p];
        else if (type == MR_LAKE_DOLL03) v1 = d.d03MaxS[p];
        else if (HY && type == MR_LAKE_HANASAKI06) {
            if (!d.h06) { raise(d.err, 20, p, E_LAKE_PARAM); return; }
            v1 = d.h06[d.lakeSlot[p]].Smax;
        }
        else if (HY && type == MR_LAKE_HYPE) {
            if (!d.hyp) { raise(d.err, 20, p, E_LAKE_PARAM); return; }
            const HypeParams &hp = d.hyp[d.lakeSlot[p]];
            v1 = (hp.E_emr - hp.E_zero) * hp.A_avg;
        }
        else { raise(d.err, 20, p, E_LAKE_TYPE); return; }
    }
    const double v0 = v1;
    const double qr1 = d.qrSer[(size_t)(t + 1) * N + p];
    // lake forcing of this step; the evaporation may have been cut back by a method routed earlier in the step
    const bool ep = HY && d.lakeEvap != nullptr;      // (forcing, like the parametric models, only in the HY instantiation)
    size_t ix = 0;
    double pr = 0.0, ev = 0.0;
    if (ep) { ix = (size_t)t * d.nLake + d.lakeSlot[p]; pr = d.lakePrecip[ix]; ev = d.lakeEvap[ix]; }
    v1 = v1 + qup * dt;
    if (d.lakeInputOption == 1 || d.lakeInputOption == 2) v1 = v1 + qr1 * dt;
    if (d.lakeInputOption == 0 || d.lakeInputOption == 2) {       // lake_route.f90:166-174
        v1 = v1 + pr * dt;
        if (v1 > ev * dt) v1 = v1 - ev * dt;
        else {                                         // not enough water to evaporate: basinevapo is updated
            if (ep) { ev = v1 / dt; d.lakeEvap[ix] = ev; }
            v1 = 0.0;
        }
    }
    double took = 0.0;                                 // REACH_WM_FLUX_actual
    if (HY && d.wmFlux) {                              // lake_route.f90:176-193
        const double f = d.wmFlux[(size_t)t * N + p];
        took = f;
        if (f != -9999.0) {
            if (f <= 0) v1 = v1 - f * dt;
            else if (f * dt <= v1) v1 = v1 - f * dt;
            else { took = v1 / dt; v1 = 0.0; }
        }
    }
    double q;
    if (follows) {                                     // lake_route.f90:196-203
        if (v1 < target) q = 0;
        else { q = (v1 - target) / dt; v1 = target; }
    } else
    if (type == MR_LAKE_ENDORHEIC) {
        q = 0.0;
    } else if (type == MR_LAKE_DOLL03) {
        const double s0 = d.d03S0[p];
        if ((v1 - s0) > 0) q = d.d03Coef[p] * (v1 - s0) * pow((v1 - s0) / (d.d03MaxS[p] - s0), d.d03Pow[p]);
        else q = 0;
        q = q / 86400.0;
        q = fmin(q, v1 / dt);
        v1 = v1 - q * dt;
    } else if (HY && type == MR_LAKE_HANASAKI06) {
        if constexpr (HY) {
            if (!d.h06) { raise(d.err, 20, p, E_LAKE_PARAM); return; }
            if (!d.stepMonth) { raise(d.err, 20, p, E_NO_CALENDAR); return; }
            const double prevQ = t > 0 ? d.qSer[M][(size_t)(t - 1) * N + p] : (d.lastK > 0 ? d.qSer[M][(size_t)(d.lastK - 1) * N + p] : 0.0);   // REACH_Q of the previous step
            q = h06_release(d.h06 + d.lakeSlot[p], d.h06Mem, d.stepMonth[t], d.stepDay[t], d.noleap, v1, qup, prevQ, dt);
            v1 = v1 - q * dt;
        } else q = 0.0;
    } else if (HY && type == MR_LAKE_HYPE) {
        if constexpr (HY) {
            if (!d.hyp) { raise(d.err, 20, p, E_LAKE_PARAM); return; }
            if (!d.stepDoy) { raise(d.err, 20, p, E_NO_CALENDAR); return; }
            q = hype_outflow(d.hyp + d.lakeSlot[p], d.stepDoy[t], v1, dt);
            v1 = v1 - q * dt;
        } else q = 0.0;
    } else { raise(d.err, 20, p, E_LAKE_TYPE); return; }
    Qs[p] = q;
    d.vol0[M][p] = v0; d.vol1[M][p] = v1;
    d.wb[M][p] = lake_wb(v1, v0, qup, qr1, q, dt, pr, ev, ep, took);
}

}  // namespace mr
