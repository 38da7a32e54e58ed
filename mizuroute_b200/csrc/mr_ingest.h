// Forcing ingest on the device: what get_basin_runoff does to the runoff between the forcing file and main_route
// (get_basin_runoff.f90:19-251) -- the time-weighted mean over the forcing records under a simulation step where dt_qsim and
// dt_ro differ (read_1D_forcing, read_runoff.f90:298-325; weights from timeMap_sim_forc, get_basin_runoff.f90:256-369, built by
// the caller), scale_forcing (:375-423) and sort_flux (process_remap.f90:271-314: forcing HRU -> river-network HRU, HRUs without
// forcing and negative values -> 0).  One value per (river-network HRU, step); compiles for the device (k_ingest) and for the
// host (tests/emul).
#pragma once
#include "mr_dev.h"

namespace mr {

// rec [nRec][nIn]: forcing records; src: column of this HRU in them (-1 = none); recIdx / recFrac [j0, j1): records under the
// step and their share of it (recFrac == nullptr: the step takes recIdx[j0] as it is); rescale: v -> A*v + B
MR_DEV double ingest_value(const double *rec, int nIn, int src, const int *recIdx, const double *recFrac, int j0, int j1,
                           int rescale, double A, double B, double fill) {
    if (src < 0) return 0.0;                                   // realMissing -> 0, sort_flux with remove_negatives
    double v;
    if (!recFrac) v = rec[(size_t)recIdx[j0] * nIn + src];
    else {
        double wsum = 0.0, wtot = 0.0;
        for (int j = j0; j < j1; ++j) {
            const double x = rec[(size_t)recIdx[j] * nIn + src];
            if (x != fill) { wsum = wsum + x * recFrac[j]; wtot = wtot + recFrac[j]; }
        }
        v = wtot == 0.0 ? fill : (wtot < 1.0 ? wsum / wtot : wsum);
    }
    if (rescale && v != fill && v != -9999.0) v = A * v + B;
    if (v == fill || v < 0.0) v = 0.0;
    return v;
}

}  // namespace mr
