// Host build of mizuroute_b200/csrc/mr_calendar.h (the calendar the library derives for every step of a batch) for
// tests/test_calendar.py.  TEST INFRASTRUCTURE ONLY.
#include "../../mizuroute_b200/csrc/mr_calendar.h"

extern "C" void calendar_emul(int y0, int m0, int d0, double sec0, int noleap, double dt, long long step, int *month, int *day, int *doy) {
    mr::step_calendar(y0, m0, d0, sec0, noleap != 0, dt, step, *month, *day, *doy);
}
