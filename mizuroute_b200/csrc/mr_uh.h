// Init-time unit-hydrograph producers (host, double precision, evaluated one IEEE operation at a time).
//
//   incomplete gamma  P(a,x)     gamma_func.f90:16-121  (Numerical-Recipes gser/gcf with the 6-term Lanczos gammln)
//   hillslope UH      FRAC_FUTURE process_param.f90:13-92  (basinUH)
//   reach UH          UH(:)       process_param.f90:99-262 (make_uh, Saint-Venant impulse response)
//
// The routing results depend on these tables to the last bit, so the operation order of the reference
// is kept (including its single-precision literals 0.99999 / 0.9999, process_param.f90:205,211,242).
#pragma once
#include <cfloat>
#include <cmath>
#include <vector>

namespace mr {

struct GammaFn {
    static double lanczos_lngamma(double xx) {                       // gamma_func.f90:104-121
        static const double c[6] = {76.18009172947146, -86.50532032941677, 24.01409824083091,
                                    -1.231739572450155, 0.1208650973866179e-2, -0.5395239384953e-5};
        double t = xx + 5.5;
        t = (xx + 0.5) * std::log(t) - t;
        double den = xx + 1.0, series = 0.0;                         // arth(x+1,1,6): repeated addition
        for (double cj : c) { series += cj / den; den = den + 1.0; }
        return t + std::log(2.5066282746310005 * (1.000000000190015 + series) / xx);
    }
    static double series(double a, double x) {                       // gser, gamma_func.f90:30-60
        if (x == 0.0) return 0.0;
        double ap = a, term = 1.0 / a, total = term;
        for (int it = 0; it < 100; ++it) {
            ap = ap + 1.0;
            term = term * x / ap;
            total = total + term;
            if (std::fabs(term) < std::fabs(total) * DBL_EPSILON) break;
        }
        return total * std::exp(-x + a * std::log(x) - lanczos_lngamma(a));
    }
    static double contfrac(double a, double x) {                     // gcf, gamma_func.f90:65-99
        if (x == 0.0) return 1.0;
        const double fpmin = DBL_MIN / DBL_EPSILON;
        double b = x + 1.0 - a, c = 1.0 / fpmin, d = 1.0 / b, h = d;
        for (int i = 1; i <= 100; ++i) {
            double an = -i * (i - a);
            b = b + 2.0;
            d = an * d + b;
            if (std::fabs(d) < fpmin) d = fpmin;
            c = b + an / c;
            if (std::fabs(c) < fpmin) c = fpmin;
            d = 1.0 / d;
            double del = d * c;
            h = h * del;
            if (std::fabs(del - 1.0) <= DBL_EPSILON) break;
        }
        return std::exp(-x + a * std::log(x) - lanczos_lngamma(a)) * h;
    }
    static double P(double a, double x) {                            // gammp, gamma_func.f90:16-25
        return (x < a + 1.0) ? series(a, x) : 1.0 - contfrac(a, x);
    }
};

// process_param.f90:13-92.  Returns ierr (20: bisection for the number of bins failed).
inline int build_hillslope_uh(double dt, double fshape, double tscale, std::vector<double> &frac) {
    double trial;
    if (GammaFn::P(fshape, dt / tscale) > 0.999) {
        trial = 1.999;
    } else {
        double lo = 1.0, hi = 1000.0;
        trial = 0.5 * (lo + hi);
        bool found = false;
        for (int it = 1; it <= 100; ++it) {
            double cp = GammaFn::P(fshape, dt * trial / tscale);
            if (cp < 0.99) lo = trial;
            if (cp > 0.999) hi = trial;
            if (cp > 0.99 && cp < 0.999) { found = true; break; }
            trial = 0.5 * (lo + hi);
        }
        if (!found) return 20;
    }
    const int n = (int)std::ceil(trial);
    frac.assign(n, 0.0);
    double prev = 0.0;
    for (int j = 1; j <= n; ++j) {
        double cp = GammaFn::P(fshape, ((double)j * dt) / tscale);
        frac[j - 1] = std::fmax(0.0, cp - prev);
        prev = cp;
    }
    double total = 0.0;
    for (double v : frac) total += v;
    for (double &v : frac) v = v / total;
    return 0;
}

// process_param.f90:99-262 for one reach.  out must hold >= 240 values; returns ntdh.
inline int build_reach_uh(double length, double dt, double velo, double diff, double *out) {
    constexpr int NH = 240;                       // nTMAX = nHr = 240 hourly ordinates
    const double hour = 3600.0, pi_ref = 3.14159265359;          // public_var.f90:15
    const double cut_hi = (double)0.99999f, cut_lo = (double)0.9999f;
    double kern[NH + 1], conv[NH + 1], box[NH + 1];
    const int nsub = (int)std::ceil(dt / hour);
    for (int k = 1; k <= NH; ++k) box[k] = (k <= nsub) ? 1.0 / nsub : 0.0;

    double acc = 0.0, sec = 0.0;
    for (int i = 1; i <= NH; ++i) {
        sec = sec + hour;
        double hval = 0.0;
        if (velo > 0.0) {
            double dist = velo * sec - length;
            double pot = (dist * dist) / (4.0 * diff * sec);
            if (!(pot > 69.0)) hval = 1.0 / (2.0 * std::sqrt(pi_ref * diff * sec)) * length * std::exp(-pot);
        }
        kern[i] = hval;
        acc = acc + hval;
    }
    if (acc > 0.0) for (int i = 1; i <= NH; ++i) kern[i] = kern[i] / acc;

    int last = 1, first = 1;
    acc = 0.0;
    for (int i = 1; i <= NH; ++i) { acc = acc + kern[i]; last = i; if (acc > cut_hi) break; }
    acc = 0.0;
    for (int i = NH; i >= 1; --i) { acc = acc + kern[i]; first = i; if (acc > cut_hi) break; }

    acc = 0.0;
    for (int j = 1; j <= NH; ++j) {
        double s = 0.0;
        for (int i = first; i <= last; ++i) {
            const int lag = j - i;
            if (lag <= 0) break;
            if (lag <= nsub) s = s + box[lag] * kern[i];
        }
        conv[j] = s;
        acc = acc + s;
    }
    if (acc > 0.0) for (int j = 1; j <= NH; ++j) conv[j] = conv[j] / acc;

    acc = 0.0;
    for (int i = 1; i <= NH; ++i) { acc = acc + conv[i]; last = i; if (acc > cut_lo) break; }
    for (int i = 1; i <= NH; ++i) conv[i] = conv[i] / acc;

    const int ntdh = (last + nsub - 1) / nsub;
    for (int k = 0; k < ntdh; ++k) out[k] = 0.0;
    for (int j = 1; j <= last; ++j) {
        const int bin = (j + nsub - 1) / nsub;
        out[bin - 1] = out[bin - 1] + conv[j];
    }
    return ntdh;
}

}  // namespace mr
