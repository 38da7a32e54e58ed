// Host side: the per-lake records of the parametric lake models (HypeParams, H06Lake of mr_dev.h) from the named per-reach
// parameter arrays a caller hands over with mr_set_lake_param.  Shared by mr_set_network and the host emulation in
// tests/emul, so the CPU tests exercise the very code that fills the device image.
#pragma once
#include <cmath>
#include <cstring>
#include <functional>
#include <string>
#include <vector>
#include "mr_dev.h"

namespace mr {

// name -> values in the caller's reach order, or nullptr when the parameter was not given (or has the wrong length)
using LakeParamLookup = std::function<const std::vector<double> *(const std::string &)>;

// HYP_* by lake slot (dataTypes.f90:202-213).  Returns the name of a missing parameter, "" when complete.
inline std::string build_hype_params(int nLake, const int *posOfSlot, const int *pos2rch, const LakeParamLookup &par, std::vector<HypeParams> &out) {
    static const char *names[HYP_COUNT] = {"HYP_E_emr", "HYP_E_lim", "HYP_E_min", "HYP_E_zero", "HYP_Qrate_emr", "HYP_Erate_emr",
                                           "HYP_Qrate_prim", "HYP_Qrate_amp", "HYP_Qrate_phs", "HYP_prim_F", "HYP_A_avg", "HYP_Qsim_mode"};
    out.assign(nLake > 0 ? nLake : 1, HypeParams());
    for (int k = 0; k < HYP_COUNT; ++k) {
        const std::vector<double> *v = par(names[k]);
        if (!v) return names[k];
        for (int s = 0; s < nLake; ++s) reinterpret_cast<double *>(&out[s])[k] = (*v)[pos2rch[posOfSlot[s]]];
    }
    return "";
}

// H06_* by lake slot (dataTypes.f90:215-254) for the lakes of type Hanasaki-2006, with the layout of their inflow memory
// (memDoubles = doubles needed for all [12][L31] blocks).  Returns the name of a missing parameter, "" when complete.
inline std::string build_h06_lakes(int nLake, const int *posOfSlot, const int *pos2rch, const int *lakeTypeByPos, double dt,
                                   const LakeParamLookup &par, std::vector<H06Lake> &out, long long &memDoubles) {
    static const char *mon[12] = {"Jan", "Feb", "Mar", "Apr", "May", "Jun", "Jul", "Aug", "Sep", "Oct", "Nov", "Dec"};
    static const char *scal[10] = {"H06_Smax", "H06_alpha", "H06_envfact", "H06_c1", "H06_c2", "H06_exponent", "H06_denominator",
                                   "H06_c_compare", "H06_frac_Sdead", "H06_E_rel_ini"};
    out.resize(nLake > 0 ? nLake : 1);
    memDoubles = 0;
    std::string missing;
    auto val = [&](const std::string &nm, int r) -> double { const std::vector<double> *v = par(nm); if (!v) { if (missing.empty()) missing = nm; return 0.0; } return (*v)[r]; };
    for (int s = 0; s < nLake; ++s) {
        H06Lake &L = out[s];
        std::memset(&L, 0, sizeof L);
        if (lakeTypeByPos[posOfSlot[s]] != MR_LAKE_HANASAKI06) continue;
        const int r = pos2rch[posOfSlot[s]];
        double *sc = &L.Smax;                           // the ten scalars are the first members, in this order
        for (int k = 0; k < 10; ++k) sc[k] = val(scal[k], r);
        for (int k = 0; k < 12; ++k) { L.I[k] = val(std::string("H06_I_") + mon[k], r); L.D[k] = val(std::string("H06_D_") + mon[k], r); }
        L.purpose = (int)val("H06_purpose", r);
        L.memF = val("H06_I_mem_F", r) != 0.0 ? 1 : 0;
        const double yrs = (double)(int)val("H06_I_mem_L", r);
        // row lengths of QPASTUP_IRF: floor(mem_L * {31, 30, 28.25 | 28} * secprday / dt), lake_route.f90:236-275
        L.L31 = (int)std::floor(yrs * 31 * 86400.0 / dt); L.L30 = (int)std::floor(yrs * 30 * 86400.0 / dt);
        L.LF = (int)std::floor(yrs * 28.25 * 86400.0 / dt); L.LFnoleap = (int)std::floor(yrs * 28 * 86400.0 / dt);
        L.memOff = memDoubles;
        if (L.memF) memDoubles += 12LL * L.L31;
    }
    return missing;
}

}  // namespace mr
