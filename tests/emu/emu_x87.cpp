// CPU check of methylasso_b200/csrc/x87.cuh against the host's long double (x87 80-bit on x86-64).
// Test infrastructure only.
#include <cmath>
#include <cstdint>
#include "../../methylasso_b200/csrc/x87.cuh"

extern "C" {

double emu_x87_mean(int n, const double* x) { return ml::x87_mean(n, [&](int i) { return x[i]; }); }

// R's real_mean (src/main/summary.c) with the host's long double
double ref_x87_mean(int n, const double* x) {
  volatile long double s = 0.0L;
  for (int i = 0; i < n; ++i) s = s + (long double)x[i];
  s = s / n;
  if (std::isfinite((double)s)) {
    volatile long double t = 0.0L;
    for (int i = 0; i < n; ++i) t = t + ((long double)x[i] - s);
    s = s + t / n;
  }
  return (double)s;
}

// count of mismatches over m arrays of n values each
long long emu_x87_compare(long long m, int n, const double* x, long long* first_bad) {
  long long bad = 0;
  *first_bad = -1;
  for (long long k = 0; k < m; ++k) {
    const double a = emu_x87_mean(n, x + k * n), b = ref_x87_mean(n, x + k * n);
    if (!(a == b) && !(a != a && b != b)) { if (bad == 0) *first_bad = k; ++bad; }
  }
  return bad;
}

int ref_long_double_digits() { return __LDBL_MANT_DIG__; }
}
