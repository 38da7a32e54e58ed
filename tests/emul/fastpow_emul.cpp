// Host build of mr_pow04_fast (mizuroute_b200/csrc/mr_dev.h), the x**0.4 of the KWT celerity -- test infrastructure only.
#include "../../mizuroute_b200/csrc/mr_dev.h"

extern "C" void fastpow_eval(int n, const double *x, double *y) {
    for (int i = 0; i < n; ++i) y[i] = mr::mr_pow04_fast(x[i]);
}
