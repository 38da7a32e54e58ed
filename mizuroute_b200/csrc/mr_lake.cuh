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
MR_LAKE_FN double lake_wb(double v1, double v0, double qup, double qlat, double q, double dt, double pr, double ev, bool ep) {
    const double dVol = v1 - v0;
    const double Qin = qup * dt, Qlateral = qlat * dt;
    const double precip = ep ? pr * dt : 0.0;
    const double Qout = -1.0 * q * dt;
    const double Qtake = -1.0 * 0.0 * dt;
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

// HY: the domain holds HYPE reservoirs.  The kernels are instantiated for both values and the launch picks one, so that a
// domain without them runs exactly the code it ran before HYPE existed (no extra registers, stack or spills).
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
    if (tau == 0) {                                    // iTime==1 cold start, lake_route.f90:139-157
        if (type == MR_LAKE_ENDORHEIC) v1 = d.d03S0[p];
        else if (type == MR_LAKE_DOLL03) v1 = d.d03MaxS[p];
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
    const bool ep = d.lakeEvap != nullptr;
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
    double q;
    if (type == MR_LAKE_ENDORHEIC) {
        q = 0.0;
    } else if (type == MR_LAKE_DOLL03) {
        const double s0 = d.d03S0[p];
        if ((v1 - s0) > 0) q = d.d03Coef[p] * (v1 - s0) * pow((v1 - s0) / (d.d03MaxS[p] - s0), d.d03Pow[p]);
        else q = 0;
        q = q / 86400.0;
        q = fmin(q, v1 / dt);
        v1 = v1 - q * dt;
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
    d.wb[M][p] = lake_wb(v1, v0, qup, qr1, q, dt, pr, ev, ep);
}

}  // namespace mr
