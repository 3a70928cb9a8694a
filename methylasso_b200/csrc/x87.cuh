// x87.cuh — R's mean() of a few doubles, bit for bit.
//
// R (summary.c real_mean; data.table's fastmean the same) accumulates in `long double`, the x87 80-bit format on
// x86-64: s = sum(x) / n, then s += sum(x - s) / n, every operation rounded to a 64-bit significand, and the result
// rounded once more to double.  That double rounding makes the result differ from the correctly rounded mean in
// about one case in 2^11, and R/call.R feeds these per-position means to a rank test where a last-bit difference
// can create or remove a tie (exact test vs normal approximation).  So the 64-bit roundings are reproduced here: a
// value is kept as an unevaluated pair hi + lo with hi = the nearest double and lo a multiple of the 64-bit ulp.
// Device logic shared with the CPU emulation (tests/emu/emu_x87.cpp checks it against the host's long double).
#pragma once
#include "hd.h"

namespace ml {

struct X87 { double hi, lo; };

ML_HD double x87_fma(double a, double b, double c) {
#ifdef __CUDA_ARCH__
  return __fma_rn(a, b, c);
#else
  return fma(a, b, c);
#endif
}

// hi + lo (|lo| well below |hi|), exact to far beyond 64 bits -> rounded to a 64-bit significand, ties to even
ML_HD X87 x87_round(double hi, double lo) {
  double s = hi + lo;
  lo = lo - (s - hi);                             // fast two-sum: s = RN(hi + lo), lo = the rest
  hi = s;
  if (hi == 0.0 || !is_finite(hi)) return X87{hi, 0.0};
  const uint64_t bits = f64_bits(hi);
  int e = (int)((bits >> 52) & 0x7ffull) - 1023;  // hi = 1.m * 2^e (subnormals do not occur for ratios in [0, 1])
  const bool pow2 = (bits & 0xfffffffffffffull) == 0;
  if (pow2 && lo != 0.0 && ((lo < 0.0) == (hi > 0.0))) --e;     // the value sits in the binade below hi
  // ulp of a 64-bit significand at this exponent: 2^(e - 63), built from its bit pattern
  const int qe = e - 63 + 1023;
  if (qe <= 0) return X87{hi, lo};                // out of the range of interest: leave the value as it is
  if (qe >= 2046) return X87{hi, lo};
  // q and 1 / q are powers of two: the "quotient" lo / q is a multiplication, exact like the division it replaces
  // (an fp64 division is ~120 cycles on the device and this function runs twice per element of every mean)
  const uint64_t qb = (uint64_t)qe << 52, ib = (uint64_t)(2046 - qe) << 52;
  double q, qi;
#ifdef __CUDA_ARCH__
  q = __longlong_as_double((long long)qb);
  qi = __longlong_as_double((long long)ib);
#else
  memcpy(&q, &qb, 8);
  memcpy(&qi, &ib, 8);
#endif
  const double t = rint(lo * qi);                 // |lo / q| <= 2^11: exact, rint rounds half to even, and
  return X87{hi, t * q};                          // hi / q is even, so this is ties-to-even of the significand
}

ML_HD X87 x87_add_d(const X87& a, double x) {
  const double s = a.hi + x;
  const double bb = s - a.hi;
  const double e = (a.hi - (s - bb)) + (x - bb);  // two-sum: exact
  return x87_round(s, e + a.lo);
}
ML_HD X87 x87_add(const X87& a, const X87& b) {
  const double s = a.hi + b.hi;
  const double bb = s - a.hi;
  const double e = (a.hi - (s - bb)) + (b.hi - bb);
  return x87_round(s, e + (a.lo + b.lo));
}
ML_HD X87 x87_d_sub(double x, const X87& a) {     // x - a
  const double s = x - a.hi;
  const double bb = s - x;
  const double e = (x - (s - bb)) + (-a.hi - bb);
  return x87_round(s, e - a.lo);
}
ML_HD X87 x87_div_n(const X87& a, int n) {
  const double d = (double)n;
  const double q1 = a.hi / d;
  const double r = x87_fma(-q1, d, a.hi);         // exact remainder of the leading quotient
  return x87_round(q1, (r + a.lo) / d);
}
ML_HD double x87_to_double(const X87& a) { return a.hi + a.lo; }

// mean(x[0 .. n)) as R computes it; `val(i)` returns x[i]
template <class F>
ML_HD double x87_mean(int n, F val) {
  X87 s{0.0, 0.0};
  for (int i = 0; i < n; ++i) s = x87_add_d(s, val(i));
  s = x87_div_n(s, n);
  if (is_finite(x87_to_double(s))) {
    X87 t{0.0, 0.0};
    for (int i = 0; i < n; ++i) t = x87_add(t, x87_d_sub(val(i), s));
    s = x87_add(s, x87_div_n(t, n));
  }
  return x87_to_double(s);
}

}  // namespace ml
