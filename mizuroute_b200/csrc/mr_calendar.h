// Calendar of a simulation step (host side): simDatetime(1) = <sim_start> + step*dt in the standard (proleptic Gregorian)
// or the noleap calendar, and its day of year (datetime_data.f90:209-218) -- what the HYPE lake model reads.
#pragma once
#include <cmath>

namespace mr {

inline long long cal_days_from_civil(long long y, int m, int d, bool noleap) {      // days since 1970-01-01
    if (noleap) { static const int cum[12] = {0, 31, 59, 90, 120, 151, 181, 212, 243, 273, 304, 334}; return (y - 1970) * 365 + cum[m - 1] + (d - 1); }
    y -= m <= 2;
    const long long era = (y >= 0 ? y : y - 399) / 400;
    const unsigned yoe = (unsigned)(y - era * 400), doy = (153u * (unsigned)(m + (m > 2 ? -3 : 9)) + 2) / 5 + (unsigned)d - 1, doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
    return era * 146097 + (long long)doe - 719468;
}
inline void cal_civil_from_days(long long days, bool noleap, int &y, int &m, int &d) {
    if (noleap) {
        static const int ml[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
        const long long yy = days >= 0 ? days / 365 : -((-days + 364) / 365);
        int doy = (int)(days - yy * 365), k = 0;
        while (doy >= ml[k]) doy -= ml[k++];
        y = 1970 + (int)yy; m = k + 1; d = doy + 1;
        return;
    }
    const long long z = days + 719468, era = (z >= 0 ? z : z - 146096) / 146097;
    const unsigned doe = (unsigned)(z - era * 146097), yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
    const unsigned doy = doe - (365 * yoe + yoe / 4 - yoe / 100), mp = (5 * doy + 2) / 153;
    d = (int)(doy - (153 * mp + 2) / 5 + 1); m = (int)(mp < 10 ? mp + 3 : mp - 9); y = (int)(yoe + era * 400 + (m <= 2));
}
// month, day and day of year at the start of simulation step `step` (0-based)
inline void step_calendar(int y0, int m0, int d0, double sec0, bool noleap, double dt, long long step, int &month, int &day, int &doy) {
    const double t = (double)cal_days_from_civil(y0, m0, d0, noleap) * 86400.0 + sec0 + (double)step * dt;
    const long long days = (long long)std::floor((t + 1.e-6) / 86400.0);
    int y;
    cal_civil_from_days(days, noleap, y, month, day);
    doy = (int)(days - cal_days_from_civil(y, 1, 1, noleap)) + 1;
}

}  // namespace mr
