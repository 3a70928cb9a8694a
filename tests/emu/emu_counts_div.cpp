// tests/emu/emu_counts_div.cpp — TEST INFRASTRUCTURE.  The two-FMA quotient of counts_div.cuh on the CPU (the C library's fma
// is the IEEE operation the device's DFMA is): every pair checked against the division.
#include <cmath>
extern "C" long long emu_counts_div_mismatches(int max_num, int max_den) {
  long long bad = 0;
  for (int c = 1; c < max_den; ++c) {
    const double dc = c, r = 1.0 / dc;
    for (int n = 0; n <= max_num; ++n) {
      const double dn = n, q0 = dn * r;
      const double q = std::fma(std::fma(-q0, dc, dn), r, q0);
      bad += q != dn / dc;
    }
  }
  return bad;
}
