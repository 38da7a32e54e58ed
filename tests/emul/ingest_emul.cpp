// Single-threaded HOST build of the forcing ingest (mizuroute_b200/csrc/mr_ingest.h: ingest_value, the body of k_ingest).
// TEST INFRASTRUCTURE ONLY (tests/test_ingest_emul.py); nothing under mizuroute_b200/ builds, loads or links this file.
#include "../../mizuroute_b200/csrc/mr_ingest.h"

extern "C" void ingest_emul_run(int nHRU, int nIn, int K, const double *rec, const int *srcOfHru, const int *recPtr, const int *recIdx,
                                const double *recFrac /* or NULL */, int rescale, double A, double B, double fill, double *out /* [K][nHRU] */) {
    for (int t = 0; t < K; ++t)
        for (int h = 0; h < nHRU; ++h)
            out[(size_t)t * nHRU + h] = mr::ingest_value(rec, nIn, srcOfHru[h], recIdx, recFrac, recPtr[t], recPtr[t + 1], rescale, A, B, fill);
}
