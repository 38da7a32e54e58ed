// The Euler routing schemes of <route_opt> 3 / 4 / 5 -- kinematic wave (kwe_route.f90), Muskingum-Cunge (mc_route.f90)
// and diffusive wave (dfw_route.f90) -- with the channel hydraulics they share (hydraulic.f90) and the implicit
// advection-diffusion step (advection_diffusion.f90).  One thread routes one (reach, step); the reach's molecule
// (20 / 2 / 20 discharges) lives node-major in HBM.  Compiles for the device and, single-threaded, for the host
// (tests/emul), where it must match the oracle bit for bit.
//
// Operation order follows the Fortran; integer powers are the multiplication chains a compiler expands x**n to.
#pragma once
#include <cmath>
#include "mr_dev.h"

namespace mr {

// ---- hydraulic.f90: trapezoidal main channel (bottom width b, side slope zc) + floodplain (slope zf) above bankDepth bd
MR_DEV double hy_btop(double y, double b, double zc, double zf, double bd) {                  // Btop, :46-77
    if (y <= bd) return b + 2 * y * zc;
    double B = b + 2 * bd * zc;
    B = B + zf * (y - bd) * 2;
    return B;
}
MR_DEV double hy_pwet(double y, double b, double zc, double zf, double bd) {                  // Pwet, :82-113
    if (y <= bd) return b + 2 * y * sqrt(1 + zc * zc);
    double P = b + 2 * bd * sqrt(1 + zc * zc);
    P = P + 2 * (y - bd) * sqrt(1 + zf * zf);
    return P;
}
MR_DEV double hy_area(double y, double b, double zc, double zf, double bd) {                  // flow_area, :118-152
    if (y <= bd) return y * (b + zc * y);
    const double A = bd * (b + zc * bd);
    const double Bt = hy_btop(y, b, zc, zf, bd), Bb = hy_btop(bd, b, zc, zf, bd);
    return A + (y - bd) * (Bt + Bb) / 2.0;
}
MR_DEV double hy_water_height(double area, double b, double zc, double zf, double bd) {       // water_height, :157-202
    const double A_bank = hy_area(bd, b, zc, zf, bd);
    if (area > A_bank) {
        const double Bb = hy_btop(bd, b, zc, zf, bd);
        const double disc = Bb * Bb - 4.0 * zf * (A_bank - area);
        return bd + (-Bb + sqrt(disc)) / (2.0 * zf);
    }
    if (zc == 0) return area / b;
    return (-b + sqrt(b * b + 4.0 * area * zc)) / (2.0 * zc);
}
// What flow_depth / celerity evaluate from the channel alone, i.e. the same for every discharge: computed once per
// (reach, step) with the reference's expressions and reused by the sub-steps of Muskingum-Cunge (same values, fewer pow calls)
struct HyChannel { double Abf, Pbf, Bbf, Qbf, nPow06; };
MR_DEV HyChannel hy_channel(double b, double zc, double S, double n, double zf, double bd) {
    HyChannel c;
    c.Abf = hy_area(bd, b, zc, zf, bd); c.Pbf = hy_pwet(bd, b, zc, zf, bd); c.Bbf = hy_btop(bd, b, zc, zf, bd);
    c.Qbf = c.Abf * mr_pow(c.Abf / c.Pbf, 2.0 / 3.0) * sqrt(S) / n;          // bankfull uniform flow, hydraulic.f90:351
    c.nPow06 = mr_pow(n, 0.6);                                              // n**0.6 of celerity, hydraulic.f90:466
    return c;
}

// normal depth by Newton-Raphson to 0.5 % (flow_depth, :299-420; the schemes always pass bankDepth: floodplain = .true.)
MR_DEV_NOINLINE double hy_flow_depth(double Q, double b, double zc, double S, double n, double zf, double bd, double Abf, double Pbf, double Bbf, double Qbf) {
    const double c13 = 1.0 / 3.0, c23 = 2.0 / 3.0, c53 = 5.0 / 3.0, c103 = 10.0 / 3.0, err_thresh = 0.005, Qmin = 1.e-50;
    double error = 100.0, depth = 0.0, y0;
    if (!(Q > Qmin)) return 0.0;
    if (Q < Qbf) {
        const double t = sqrt(S) / n / Q, Coef1 = t * t * t;
        const double Coef2 = 2 * sqrt(zc * zc + 1.0);
        y0 = mr_pow(1.0 / Coef1 / (b * b * b), 1.0 / 5.0);
        MR_NOUNROLL
        while (error > err_thresh) {
            const double A = hy_area(y0, b, zc, zf, bd), Bt = hy_btop(y0, b, zc, zf, bd), P = hy_pwet(y0, b, zc, zf, bd);
            const double A2 = A * A, A4 = A2 * A2, A5 = A4 * A;
            const double hh = Coef1 * A5 / (P * P) - 1.0;
            const double dhdy = Coef1 * (5 * A4 * Bt * P - 2 * Coef2 * A5) / (P * P * P);
            depth = y0 - hh / dhdy;
            error = fabs((depth - y0) / depth);
            y0 = depth;
        }
    } else {
        y0 = bd + 2.0;
        const double Coef1 = sqrt(S) / n / mr_pow(Pbf, c23);
        const double Coef2 = 2 * mr_pow(zf / 2, c53) * sqrt(S) / n / mr_pow(zf * zf + 1.0, c13);
        MR_NOUNROLL
        while (error > err_thresh) {
            const double ye = y0 - bd;
            const double hh = Coef1 * mr_pow(Abf + Bbf * ye, c53) + Coef2 * mr_pow(ye, c103) / mr_pow(ye, c23) - Q;
            const double dhdy = Coef1 * c53 * Bbf * mr_pow(Abf + Bbf * ye, c23) + Coef2 * (c103 - c23) * mr_pow(ye, c53);
            depth = y0 - hh / dhdy;
            error = fabs((depth - y0) / depth);
            y0 = depth;
        }
    }
    return depth;
}
MR_DEV double hy_friction_slope(double Q, double y, double b, double zc, double n, double zf, double bd) {
    const double A = hy_area(y, b, zc, zf, bd), P = hy_pwet(y, b, zc, zf, bd);
    const double t = Q * n / A / mr_pow(A / P, 2.0 / 3.0);
    return t * t;
}
MR_DEV_NOINLINE double hy_celerity(double Q, double y, double b, double zc, double n, double zf, double bd, double nPow06) {   // celerity, :425-471
    if (!(y > 0.0)) return 0.0;
    const double Bt = hy_btop(y, b, zc, zf, bd);
    const double Sf = hy_friction_slope(Q, y, b, zc, n, zf, bd);       // useFrictionSlope = .true.
    return 5.0 / 3.0 * mr_pow(Sf, 0.3) * mr_pow(Q, 0.4) / mr_pow(Bt, 0.4) / nPow06;
}
MR_DEV_NOINLINE double hy_diffusivity(double Q, double y, double b, double zc, double n, double zf, double bd) { // diffusivity, :476-522
    if (!(y > 0.0)) return 0.0;
    const double Bt = hy_btop(y, b, zc, zf, bd);
    const double Sf = hy_friction_slope(Q, y, b, zc, n, zf, bd);
    return fabs(Q) / Sf / Bt / 2.0;
}

// upstream discharge, lateral flow and "is headwater" of a reach (kwe_route.f90:82-113 = mc_route.f90:80-110 = dfw_route.f90:86-117)
template <int M>
MR_DEV bool euler_inflow(const DevNet &d, int p, int t, double &qup, double &qlat) {
    const int N = d.nRch, nUps = d.nGood[p], u0 = d.upPtr[p];
    const double *Qs = d.qSer[M] + (size_t)t * N;
    const double qr1 = d.qrSer[(size_t)(t + 1) * N + p];
    qup = 0.0; qlat = 0.0;
    if (nUps > 0) {
        for (int m = 0; m < nUps; ++m) qup = qup + Qs[d.upIdx[u0 + m]];
        qlat = qr1;
        return false;
    }
    if (d.hwDrain == 1) { qup = qup + qr1; qlat = 0.0; }
    else if (d.hwDrain == 2) { qlat = qr1; }
    return true;
}

// reach_wb of mr_kernels.cuh (water_balance.f90:67-87), repeated here so that this header stands alone for the host build
MR_DEV double euler_wb(double v1, double v0, double qup, double qlat, double q, double dt, double took = 0.0) {
    const double dVol = v1 - v0;
    const double Qin = qup * dt, Qlateral = qlat * dt, precip = 0.0, evapo = 0.0;
    const double Qout = -1.0 * q * dt;
    const double Qtake = -1.0 * took * dt;
    return dVol - (Qin + Qlateral + precip + Qtake + Qout + evapo);
}

// kw_rch / dfw_rch + kinematic_wave / diffusive_wave (kwe_route.f90:40-365, dfw_route.f90:43-372): one implicit step of
// solve_ade (advection_diffusion.f90:19-262, central differences, Neumann outlet, wck = wdk = 1) by the Thomas algorithm.
// The tridiagonal coefficients are constants by row range, so only the forward-sweep pivots D and right-hand sides b1
// are kept; every element is computed with the reference's expression.
template <int M, bool EXT = false>
MR_DEV void kw_dw_reach(const DevNet &d, int p, int t) {
    static_assert(M == M_KW || M == M_DW, "kinematic / diffusive wave");
    constexpr int NM = NMOL<M>;
    const int N = d.nRch;
    const double dt = d.dt;
    double qup, qlat;
    const bool isHW = euler_inflow<M>(d, p, t, qup, qlat);
    double *mol = d.mol[M] + p;                                    // node k at mol[k * N]
    double v1 = d.vol1[M][p], v0 = v1, q, flood = 0.0, ele = 0.0;
    d.inflow[M][p] = qup;
    const double qin = qup;                                        // the solver sees Qupstream_mod, the water balance Qupstream
    double took = 0.0;
    if (EXT && d.wmFlux) took = wm_cascade(d.wmFlux[(size_t)t * N + p], dt, v1, qup, qlat);
    const double L = d.rlength[p];
    if (!isHW || d.hwDrain == 1) {
        if (L > d.minLengthRoute) {
            const double S = d.rslope[p], n = d.rmann[p], bt = d.rwidth[p], bd = d.rdepth[p], zc = d.sideSlope[p], zf = d.fldpSlope[p];
            const double Qbar = (qup + mol[0] + mol[(size_t)(NM - 2) * N]) / 3.0;
            const HyChannel hc = hy_channel(bt, zc, S, n, zf, bd);
            const double depth = hy_flow_depth(fabs(Qbar), bt, zc, S, n, zf, bd, hc.Abf, hc.Pbf, hc.Bbf, hc.Qbf);
            const double ck = hy_celerity(fabs(Qbar), depth, bt, zc, n, zf, bd, hc.nPow06);
            const double dk = M == M_DW ? hy_diffusivity(fabs(Qbar), depth, bt, zc, n, zf, bd) : 0.0;
            const double wck = 1.0, wdk = 1.0;
            const double dx = L / ((NM - 1) - 1);
            const double Cd = dk * dt / (dx * dx), Ca = ck * dt / dx;
            const double diMid = 2.0 + 4 * wdk * Cd, upV = wck * Ca - 2.0 * wdk * Cd, loV = -wck * Ca - 2.0 * wdk * Cd;
            const double cA = (1.0 - wck) * Ca + 2.0 * (1.0 - wdk) * Cd, cB = 2.0 - 4.0 * (1.0 - wdk) * Cd, cC = (1.0 - wck) * Ca - 2.0 * (1.0 - wdk) * Cd;
            double D[NM], b1[NM];
            // forward sweep: up[i] = A(i-1,i) (0 for i < 2), lo[i] = A(i+1,i) (loV for i <= NM-3, -1 at NM-2)
            double prevM = mol[0], prevC = mol[N], prevP;          // previous-step values at nodes i-1, i, i+1
            D[0] = 1.0; b1[0] = qup;
#pragma unroll
            for (int i = 1; i < NM; ++i) {
                double bi, di;
                if (i < NM - 1) {
                    prevP = mol[(size_t)(i + 1) * N];
                    bi = cA * prevM + cB * prevC - cC * prevP;
                    di = diMid;
                } else {
                    bi = prevC - prevM;                            // Sbc = prev(NM) - prev(NM-1)
                    di = 1.0;
                    prevP = 0.0;
                }
                const double lo = i - 1 <= NM - 3 ? loV : -1.0;
                const double up = i >= 2 ? upV : 0.0;
                const double coef = lo / D[i - 1];
                D[i] = di - coef * up;
                b1[i] = bi - coef * b1[i - 1];
                prevM = prevC; prevC = prevP;
            }
            // back substitution; the outlet is node NM-2 (0-based): its value fixes the low-storage reduction of nodes 1..NM-1
            const double cLast = b1[NM - 1] / D[NM - 1];
            double cur = (b1[NM - 2] - upV * cLast) / D[NM - 2];
            const double qout = cur;
            double pcnt = 1.0;
            const bool reduce = fabs(qout) > 0.0;
            if (reduce) {
                const double volTmp = fmax(0.0, v1);
                const double qoutTmp = qout * dt;
                pcnt = fmin((volTmp + dt * qup) * 0.999 / qoutTmp, 1.0);
            }
            mol[(size_t)(NM - 1) * N] = reduce ? cLast * pcnt : cLast;
            mol[(size_t)(NM - 2) * N] = reduce ? cur * pcnt : cur;
#pragma unroll
            for (int i = NM - 3; i >= 0; --i) {
                const double up = i + 1 >= 2 ? upV : 0.0;
                cur = (b1[i] - up * cur) / D[i];
                mol[(size_t)i * N] = (reduce && i >= 1) ? cur * pcnt : cur;
            }
            const double qo = reduce ? qout * pcnt : qout;
            v1 = v1 + (qup - qo) * dt;
            const double stor = d.rstorage[p];
            flood = v1 > stor ? v1 - stor : 0.0;
            ele = hy_water_height(v1 / L, bt, zc, zf, bd);
            q = qo + qlat;
        } else {                                                   // pass-through
            q = qup + qlat;
#pragma unroll
            for (int i = 0; i < NM - 1; ++i) mol[(size_t)i * N] = 0.0;
            mol[(size_t)(NM - 1) * N] = q;
            v0 = 0.0; v1 = 0.0;
        }
    } else {                                                       // headwater draining at the bottom of the reach
        q = qlat;
        v0 = 0.0; v1 = 0.0;
#pragma unroll
        for (int i = 0; i < NM - 1; ++i) mol[(size_t)i * N] = 0.0;
        mol[(size_t)(NM - 1) * N] = q;
    }
    if (EXT && d.daQobs) {                                         // qmodOption 1: direct insertion, no water balance
        d.vol0[M][p] = v0; d.vol1[M][p] = v1; d.floodVol[M][p] = flood; d.reachEle[M][p] = ele;
        d.qSer[M][(size_t)t * N + p] = direct_insertion(d, M, p, t, q);
        return;
    }
    d.qSer[M][(size_t)t * N + p] = q;
    d.vol0[M][p] = v0; d.vol1[M][p] = v1; d.floodVol[M][p] = flood; d.reachEle[M][p] = ele;
    d.wb[M][p] = euler_wb(v1, v0, qin, qlat, q, dt, took);
}

// mc_rch + muskingum_cunge (mc_route.f90:45-418), sub-stepping when the Courant number exceeds one
template <bool EXT = false>
MR_DEV void mc_reach(const DevNet &d, int p, int t) {
    constexpr int M = M_MC;
    const int N = d.nRch;
    const double dt = d.dt, Y = 0.5, Qmin = 1.e-50;
    double qup, qlat;
    const bool isHW = euler_inflow<M>(d, p, t, qup, qlat);
    double *mol = d.mol[M] + p;
    const double Q00 = mol[0], Q01 = mol[N];
    double Q10, Q11, q, v1 = d.vol1[M][p], v0 = v1, flood = 0.0, ele = 0.0;
    d.inflow[M][p] = qup;
    const double qin = qup;
    double took = 0.0;
    if (EXT && d.wmFlux) took = wm_cascade(d.wmFlux[(size_t)t * N + p], dt, v1, qup, qlat);
    const double L = d.rlength[p];
    if (!isHW || d.hwDrain == 1) {
        if (L > d.minLengthRoute) {
            const double S = d.rslope[p], n = d.rmann[p], bt = d.rwidth[p], bd = d.rdepth[p], zc = d.sideSlope[p], zf = d.fldpSlope[p];
            const double theta = dt / L;
            Q10 = qup;
            double Qbar = (Q00 + Q10 + Q01) / 3.0;
            if (Qbar > Qmin) {
                const HyChannel hc = hy_channel(bt, zc, S, n, zf, bd);
                double depth = hy_flow_depth(fabs(Qbar), bt, zc, S, n, zf, bd, hc.Abf, hc.Pbf, hc.Bbf, hc.Qbf);
                double ck = hy_celerity(fabs(Qbar), depth, bt, zc, n, zf, bd, hc.nPow06);
                double Cn = ck * theta, dTsub = dt;
                int ntSub = 1;
                if (Cn > 1.0) { ntSub = (int)ceil(dt / L * ck); dTsub = dt / ntSub; }
                double QinPrev = Q00, QoutPrev = Q01, sum = 0.0;
                MR_NOUNROLL
                for (int ix = 1; ix <= ntSub; ++ix) {
                    const double Qin = Q10;
                    double Qout;
                    Qbar = (Qin + QinPrev + QoutPrev) / 3.0;
                    if (Qbar > Qmin) {
                        depth = hy_flow_depth(fabs(Qbar), bt, zc, S, n, zf, bd, hc.Abf, hc.Pbf, hc.Bbf, hc.Qbf);
                        const double topWidth = hy_btop(depth, bt, zc, zf, bd);
                        ck = hy_celerity(fabs(Qbar), depth, bt, zc, n, zf, bd, hc.nPow06);
                        const double X = 0.5 * (1.0 - Qbar / (topWidth * S * ck * L));
                        Cn = ck * dTsub / L;
                        const double C0 = (-X + Cn * (1 - Y)) / (1 - X + Cn * (1 - Y));
                        const double C1 = (X + Cn * Y) / (1 - X + Cn * (1 - Y));
                        const double C2 = (1 - X - Cn * Y) / (1 - X + Cn * (1 - Y));
                        Qout = C0 * Qin + C1 * QinPrev + C2 * QoutPrev;
                        Qout = fmax(0.0, Qout);
                    } else Qout = 0.0;
                    sum = sum + Qout;
                    QinPrev = Qin; QoutPrev = Qout;
                }
                Q11 = sum / (double)ntSub;
                if (fabs(Q11) > 0.0) {                             // "*0.999" is a single-precision literal in mc_route.f90:352
                    const double pcnt = fmin((v1 / dt + Q10) * (double)0.999f / Q11, 1.0);
                    Q11 = Q11 * pcnt;
                }
                v1 = v1 + (Q10 - Q11) * dt;
                q = Q11 + qlat;
            } else {
                Q11 = 0.0;
                q = Q11 + qlat;
                v1 = v1 + (Q10 - Q11) * dt;
            }
            const double stor = d.rstorage[p];
            flood = v1 > stor ? v1 - stor : 0.0;
            ele = hy_water_height(v1 / L, bt, zc, zf, bd);
        } else {
            Q10 = qup; Q11 = qup;
            q = qup + qlat;
            v0 = 0.0; v1 = 0.0;
        }
    } else {
        Q10 = 0.0; Q11 = 0.0;
        q = qlat;
        v0 = 0.0; v1 = 0.0;
    }
    mol[0] = Q10; mol[N] = Q11;
    if (EXT && d.daQobs) {                                         // qmodOption 1: direct insertion, no water balance
        d.vol0[M][p] = v0; d.vol1[M][p] = v1; d.floodVol[M][p] = flood; d.reachEle[M][p] = ele;
        d.qSer[M][(size_t)t * N + p] = direct_insertion(d, M, p, t, q);
        return;
    }
    d.qSer[M][(size_t)t * N + p] = q;
    d.vol0[M][p] = v0; d.vol1[M][p] = v1; d.floodVol[M][p] = flood; d.reachEle[M][p] = ele;
    d.wb[M][p] = euler_wb(v1, v0, qin, qlat, q, dt, took);
}

}  // namespace mr
