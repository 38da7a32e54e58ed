// Per-reach bodies of runoff accumulation (accum_inst_runoff, accum_runoff.f90:60-75) and of the impulse-response-function
// routing (irf_rch + conv_upsbas_qr, irf_route.f90:82-150,235-262): one thread routes one (reach, step).  Compiles for the
// device (route_reach in mr_kernels.cuh) and, single-threaded, for the host (tests/emul), where it must match the oracle
// bit for bit.
#pragma once
#include <cmath>
#include "mr_dev.h"

namespace mr {

// water balance, water_balance.f90:67-87 (river reaches: no precipitation / evaporation); took = REACH_WM_FLUX_actual
MR_DEV double reach_wb(double v1, double v0, double qup, double qlat, double q, double dt, double took = 0.0) {
    const double dVol = v1 - v0;
    const double Qin = qup * dt, Qlateral = qlat * dt, precip = 0.0, evapo = 0.0;
    const double Qout = -1.0 * q * dt;
    const double Qtake = -1.0 * took * dt;
    return dVol - (Qin + Qlateral + precip + Qtake + Qout + evapo);
}

MR_DEV void sum_reach(const DevNet &d, int p, int t) {
    const int N = d.nRch;
    double *Qs = d.qSer[M_SUM] + (size_t)t * N;
    const int u0 = d.upPtr[p], u1 = d.upPtr[p + 1];
    double q = d.qrSer[(size_t)(t + 1) * N + p];
    if (u1 > u0) {
        double qup = 0.0;
        for (int m = u0; m < u1; ++m) qup = qup + Qs[d.upIdx[m]];
        q = q + qup;
    }
    Qs[p] = q;
}

// The future-flow series QFUTURE_IRF of a reach is a ring of ntdh slots, slot-major in HBM; logical slot k of step tau
// lives at physical slot (tau + k) mod ntdh, so the eoshift of the reference is the head moving on.
// EXT: the instantiation for domains with water management (see lake_reach in mr_lake.cuh for why it is a template flag)
template <bool EXT = false>
MR_DEV void irf_reach(const DevNet &d, int p, int t, long long tau) {
    const int N = d.nRch;
    double *Qs = d.qSer[M_IRF] + (size_t)t * N;
    const int nUps = d.nGood[p], u0 = d.upPtr[p];
    const double qr1 = d.qrSer[(size_t)(t + 1) * N + p], dt = d.dt;
    double v1 = d.vol1[M_IRF][p], v0 = v1;
    double qup = 0.0, qlat = 0.0;
    if (nUps > 0) {
        for (int m = 0; m < nUps; ++m) qup = qup + Qs[d.upIdx[u0 + m]];
        qlat = qr1;
    } else if (d.hwDrain == 1) { qup = qup + qr1; qlat = 0.0; }
    else if (d.hwDrain == 2) { qlat = qr1; }
    d.inflow[M_IRF][p] = qup;
    const double qin = qup;                            // the water balance sees the inflow before the abstraction
    double took = 0.0;
    if (EXT && d.wmFlux) took = wm_cascade(d.wmFlux[(size_t)t * N + p], dt, v1, qup, qlat);
    const int nt = d.ntdh[p];
    double *qf = d.qfutIrf + p;
    const double *uh = d.uh + p;
    double q;
    if (d.rlength[p] > d.minLengthRoute) {
        const int head = (int)(tau % nt);
        int s = head;
        for (int k = 0; k < nt; ++k) {
            qf[(size_t)s * N] = qf[(size_t)s * N] + uh[(size_t)k * N] * qup;
            if (++s == nt) s = 0;
        }
        double q1 = qf[(size_t)head * N];
        q1 = fmin((fmax(0.0, v1) / dt + qup) * (double)0.999f, q1);      // single-precision literal in irf_route.f90:245
        v1 = v1 - (q1 - qup) * dt;
        q = q1 + qlat;
        qf[(size_t)head * N] = 0.0;
    } else {                                       // pass-through, irf_route.f90:255-262
        const int nxt = (int)((tau + 1) % nt);
        for (int k = 0; k < nt; ++k) qf[(size_t)k * N] = 0.0;
        qf[(size_t)nxt * N] = qup;                 // logical slot 0 as seen by the next step / by mr_get_state
        q = qup + qlat;
        v0 = 0.0; v1 = 0.0;
    }
    if (EXT && d.daQobs) {                                                  // qmodOption 1: no water balance, irf_route.f90:188-202
        d.vol0[M_IRF][p] = v0; d.vol1[M_IRF][p] = v1;
        Qs[p] = direct_insertion(d, M_IRF, p, t, q);
        return;
    }
    Qs[p] = q;
    d.vol0[M_IRF][p] = v0; d.vol1[M_IRF][p] = v1;
    d.wb[M_IRF][p] = reach_wb(v1, v0, qin, qlat, q, dt, took);
}

}  // namespace mr
